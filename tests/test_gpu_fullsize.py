"""GPU tier, BASELINE.json's full sizes: what the small oracle cases cannot reach.

* every convolution shape of ResNet-50 at N = 64, all three passes: the tcgen05 path (TF32 products) against
  this library's exact-fp32 CUDA-core kernels -- an independent implementation (plain loops over the definition) that the golden
  vectors of the reference's CUDA backend pin at small sizes (tests/test_gpu_parity_cuda.py runs the conv cases in both modes);
* size-independent properties at those sizes: linearity of the convolution, the adjoint identity <conv(x), dy> = <x, dgrad(dy)> =
  <w, wgrad(x, dy)>, batch-norm output statistics, pooling backward conserving the gradient mass;
* the LSTM of BASELINE.json configs[4] (T=256, B=64, I=H=1024) forward against the float64 oracle;
* 16-bit pooling and softmax (the float16 path the reference's VGG-16 config uses).
"""
import numpy as np
import pytest

from oracle import ops

pytestmark = pytest.mark.gpu

# C, H, K, R, stride, pad   (tools/bench_layers.py: the 20 distinct convolutions of ResNet-50 with pool1 -> 55 x 55)
R50_LAYERS = [(3, 224, 64, 7, 2, 3), (64, 55, 64, 1, 1, 0), (64, 55, 64, 3, 1, 1), (64, 55, 256, 1, 1, 0), (256, 55, 64, 1, 1, 0),
			  (256, 55, 128, 1, 2, 0), (256, 55, 512, 1, 2, 0), (128, 28, 128, 3, 1, 1), (128, 28, 512, 1, 1, 0), (512, 28, 128, 1, 1, 0),
			  (512, 28, 256, 1, 2, 0), (512, 28, 1024, 1, 2, 0), (256, 14, 256, 3, 1, 1), (256, 14, 1024, 1, 1, 0), (1024, 14, 256, 1, 1, 0),
			  (1024, 14, 512, 1, 2, 0), (1024, 14, 2048, 1, 2, 0), (512, 7, 512, 3, 1, 1), (512, 7, 2048, 1, 1, 0), (2048, 7, 512, 1, 1, 0)]

REL_TF32 = 1e-3      # BASELINE.json north_star: fp32 tensors within 1e-3 relative


def relerr(got, want):
	got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
	return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


def makeRng(seed):
	from puzzlelib_b200.backend import RandomNumberGenerator
	return RandomNumberGenerator(seed=seed)


def randn(bnd, rng, shape, scale=1.0):
	# generated on the device (these tensors reach 200 MB): a uniform fill is enough for a parity / property check
	ary = bnd.GPUArray.empty(shape, np.float32)
	bnd.fillUniform(ary, -scale, scale, rng)
	return ary


def absmax(bnd, ary):
	"""max |a| of a large device tensor: abs on the device, maximum over host chunks"""
	tmp = bnd.GPUArray.empty(ary.shape, np.float32)
	bnd.absKer(tmp, ary)
	flat = tmp.ravel()
	best, step = 0.0, 1 << 24
	for i in range(0, flat.size, step):
		best = max(best, float(flat[i:i + step].get().max()))
	return best


def absmaxDiff(bnd, a, b):
	diff = bnd.GPUArray.empty(a.shape, np.float32)
	bnd.addKer(np.float32)(diff, a, 1.0, b, -1.0)
	return absmax(bnd, diff)


def dot(bnd, a, b):
	return float(bnd.blas.dot(a.ravel(), b.ravel()))


@pytest.fixture
def exact(bnd):
	"""switches the float32 contractions to the CUDA-core kernels inside a `with exact():` block"""
	import contextlib

	@contextlib.contextmanager
	def mode():
		bnd.dnn.enableTensorOps(False)
		try:
			yield
		finally:
			bnd.dnn.enableTensorOps(True)
	return mode


@pytest.mark.parametrize("layer", R50_LAYERS, ids=lambda l: "%dx%d_%d_%dx%d_s%d" % (l[0], l[1], l[2], l[3], l[3], l[4]))
def test_resnet50_layer_at_n64_against_the_exact_fp32_kernels(bnd, exact, layer):
	C, H, K, R, s, p = layer
	N = 64
	P = (H + 2 * p - R) // s + 1
	rng = makeRng(C * 1000 + K)
	x = randn(bnd, rng, (N, C, H, H))
	w = randn(bnd, rng, (K, C, R, R), scale=1.0 / np.sqrt(C * R * R))
	dy = randn(bnd, rng, (N, K, P, P))

	y = bnd.dnn.convNd(x, w, None, s, p, 1, 1)
	dx = bnd.dnn.convNdBackwardData(dy, w, None, x, s, p, 1, None, 1, allocator=bnd.memoryPool)
	dw = bnd.GPUArray.zeros((K, C, R, R), np.float32)
	bnd.dnn.convNdBackwardParams(x, dy, w, s, p, 1, 1, False, False, dw, None, 1.0, 0.0)

	with exact():
		ye = bnd.dnn.convNd(x, w, None, s, p, 1, 1)
		dxe = bnd.dnn.convNdBackwardData(dy, w, None, x, s, p, 1, None, 1, allocator=bnd.memoryPool)
		dwe = bnd.GPUArray.zeros((K, C, R, R), np.float32)
		bnd.dnn.convNdBackwardParams(x, dy, w, s, p, 1, 1, False, False, dwe, None, 1.0, 0.0)

	assert y.shape == (N, K, P, P) and dx.shape == x.shape
	for name, got, want in (("fprop", y, ye), ("dgrad", dx, dxe), ("wgrad", dw, dwe)):
		# compared on the device: nothing of this size is rebuilt on the host
		err, ref = absmaxDiff(bnd, got, want), absmax(bnd, want)
		assert err < REL_TF32 * ref, (name, err, ref)

	# adjoint identities (exact in real arithmetic; TF32 products leave ~1e-3 relative to the magnitude sum)
	lhs, mid, rhs = dot(bnd, y, dy), dot(bnd, x, dx), dot(bnd, w, dw)
	scale = float(np.sqrt(dot(bnd, y, y) * dot(bnd, dy, dy)))
	assert abs(lhs - mid) < 2e-3 * scale and abs(lhs - rhs) < 2e-3 * scale, (lhs, mid, rhs, scale)


def test_convolution_is_linear_at_full_size(bnd):
	"""conv(a*x1 + b*x2) = a*conv(x1) + b*conv(x2) on the 256 -> 64 channel 55 x 55 layer at N = 64 (TF32: within 2e-3)"""
	rng = makeRng(5)
	x1, x2 = randn(bnd, rng, (64, 256, 55, 55)), randn(bnd, rng, (64, 256, 55, 55))
	w = randn(bnd, rng, (64, 256, 1, 1), scale=1.0 / 16)
	mix = bnd.GPUArray.empty(x1.shape, np.float32)
	bnd.linearKer(np.float32)(mix, x1, 0.5, 0.0)
	bnd.addKer(np.float32)(mix, mix, 1.0, x2, -2.0)
	y1, y2, ym = (bnd.dnn.convNd(t, w, None, 1, 0, 1, 1) for t in (x1, x2, mix))
	want = bnd.GPUArray.empty(y1.shape, np.float32)
	bnd.addKer(np.float32)(want, y1, 0.5, y2, -2.0)
	assert absmaxDiff(bnd, ym, want) < 2e-3 * absmax(bnd, want)


@pytest.mark.parametrize("shape", [(64, 64, 112, 112), (64, 256, 55, 55), (64, 512, 28, 28), (64, 2048, 7, 7)])
def test_batchnorm_properties_at_full_size(bnd, shape):
	"""per-channel mean 0 / variance 1 of the normalised output, sum(dx) = 0 and sum(dx * xhat) = 0 per channel (scale 1, bias 0)"""
	N, C, H, W = shape
	rng = makeRng(C)
	x = randn(bnd, rng, shape, scale=3.0)
	dy = randn(bnd, rng, shape)
	scale, bias = bnd.GPUArray.toGpu(np.ones((1, C, 1, 1), np.float32)), bnd.GPUArray.zeros((1, C, 1, 1), np.float32)
	mean, var = bnd.GPUArray.zeros((1, C, 1, 1), np.float32), bnd.GPUArray.toGpu(np.ones((1, C, 1, 1), np.float32))
	y, smean, sinv = bnd.dnn.batchNormNd(x, mean, var, scale, bias, 1e-5, 1.0, False)
	dx, dscale, dbias = bnd.dnn.batchNormNdBackward(dy, x, scale, smean, sinv, 1e-5)

	count = N * H * W

	def channel_sums(t):
		# [N, C, S] -> per-channel sums through the C-ABI's bias-gradient reduction
		from puzzlelib_b200.driver import lib, check, dtypeCode
		out = bnd.GPUArray.zeros((C, ), np.float32)
		check(lib.pz_bias_grad(dtypeCode(np.float32), t.ptr, out.ptr, N, C, H * W, 1.0, 0.0, None))
		return out.get().astype(np.float64)

	ysum = channel_sums(y)
	ysq = bnd.GPUArray.empty(shape, np.float32)
	bnd.mulKer(np.float32)(ysq, y, y)
	yvar = channel_sums(ysq) / count
	assert np.abs(ysum / count).max() < 1e-4
	assert np.abs(yvar - 1.0).max() < 1e-3
	assert np.abs(channel_sums(dx)).max() / count < 1e-5
	prod = bnd.GPUArray.empty(shape, np.float32)
	bnd.mulKer(np.float32)(prod, dx, y)
	assert np.abs(channel_sums(prod)).max() / count < 1e-4


def test_maxpool_backward_conserves_the_gradient_at_full_size(bnd):
	"""every output gradient lands on exactly one input position: sum(dx) = sum(dy) (3 x 3 stride-2 pool of conv1, N = 64)"""
	rng = makeRng(11)
	x = randn(bnd, rng, (64, 64, 112, 112))
	y = bnd.dnn.poolNd(x, 3, 2, 0, bnd.PoolMode.max.value)
	dy = randn(bnd, rng, y.shape)
	dx = bnd.dnn.poolNdBackward(dy, x, y, 3, 2, 0, bnd.PoolMode.max.value)
	ones_in, ones_out = bnd.GPUArray.empty(x.shape, np.float32).fill(1.0), bnd.GPUArray.empty(y.shape, np.float32).fill(1.0)
	total_in, total_out = dot(bnd, dx, ones_in), dot(bnd, dy, ones_out)
	assert abs(total_in - total_out) < 1e-3 * np.sqrt(dy.size)


def test_lstm_config5_forward_against_the_float64_oracle(bnd):
	"""BASELINE.json configs[4]: T=256, B=64, I=H=1024, float32 (TF32 recurrent GEMMs; the rounding compounds through 256 steps)"""
	T, B, I, H = 256, 64, 1024, 1024
	rng = np.random.RandomState(256)
	rnn, W, params = bnd.createRnn(I, H, np.float32, layers=1, mode=bnd.RNNMode.lstm)
	W.set((rng.randn(*W.shape) * (0.5 / np.sqrt(H))).astype(np.float32))
	host = {k: v.get().astype(np.float64) for k, v in params[0].items()}
	x = (rng.randn(T, B, I) * 0.5).astype(np.float32)
	out, _ = rnn.forward(bnd.GPUArray.toGpu(x), W, allocator=bnd.memoryPool)
	want, _ = ops.lstm_forward(x.astype(np.float64), host)
	got = out.get()
	assert got.shape == (T, B, H)
	assert relerr(got[:8], want[:8]) < 3e-3          # the first steps: plain TF32 rounding
	assert relerr(got, want) < 3e-2                  # all 256 steps: compounded through the recurrence


@pytest.mark.parametrize("mode", ["max", "avg"])
def test_pool2d_float16(bnd, mode):
	rng = np.random.RandomState(3)
	x = rng.randn(4, 6, 13, 11).astype(np.float16)
	omode = "max" if mode == "max" else "avgWithPad"
	want = ops.pool2d(x.astype(np.float64), 3, 2, 1, omode)
	dy = rng.randn(*want.shape).astype(np.float16)
	pm = (bnd.PoolMode.max if mode == "max" else bnd.PoolMode.avgWithPad).value
	gx, gdy = bnd.GPUArray.toGpu(x), bnd.GPUArray.toGpu(dy)
	y = bnd.dnn.poolNd(gx, 3, 2, 1, pm)
	assert y.dtype == np.float16 and relerr(y.get(), want) < 4e-3
	dx = bnd.dnn.poolNdBackward(gdy, gx, y, 3, 2, 1, pm)
	wantdx = ops.pool2d_bwd(x.astype(np.float64), y.get().astype(np.float64), dy.astype(np.float64), 3, 2, 1, omode)
	assert dx.dtype == np.float16 and relerr(dx.get(), wantdx) < 4e-3


@pytest.mark.parametrize("shape", [(16, 1000, 1, 1), (4, 21, 9, 7)])
def test_softmax_float16(bnd, shape):
	rng = np.random.RandomState(4)
	x = (rng.randn(*shape) * 3).astype(np.float16)
	dy = rng.randn(*shape).astype(np.float16)
	y = bnd.dnn.softmaxNd(bnd.GPUArray.toGpu(x))
	want = ops.softmax(x.astype(np.float64))
	assert y.dtype == np.float16 and np.abs(y.get().astype(np.float64) - want).max() < 2e-3
	dx = bnd.dnn.softmaxNdBackward(bnd.GPUArray.toGpu(dy), y)
	wantdx = ops.softmax_bwd(y.get().astype(np.float64), dy.astype(np.float64))
	assert np.abs(dx.get().astype(np.float64) - wantdx).max() < 4e-3
