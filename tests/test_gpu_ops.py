"""GPU tier: every operator of the hot path, called through the backend object (i.e. through the C-ABI of
libpzb200.so), against the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): bit-exact for max-pool argmax masks and index ops; fp32 tensors within 1e-3
relative (the tcgen05 path computes fp32 storage as TF32 products with fp32 accumulation, like cuDNN/cuBLAS with
TENSOR_OP_MATH on this GPU generation); elementwise / reduction kernels are held to the reference's own
fp32 bar of 1e-5 (Cuda/GPUBackend.py:218-220).
"""
import numpy as np
import pytest

from oracle import ops

pytestmark = pytest.mark.gpu

REL_TC = 1e-3      # tensor-core (TF32) contractions
ATOL32 = 1e-5      # fp32 CUDA-core kernels


def relerr(got, want):
	want = np.asarray(want, dtype=np.float64)
	return float(np.abs(np.asarray(got, dtype=np.float64) - want).max() / (np.abs(want).max() + 1e-30))


def G(bnd, ary):
	return bnd.GPUArray.toGpu(np.ascontiguousarray(ary))


# ================================================================================================ convolution
CONV_CASES = [
	# N, C, H, W, K, R, S, stride, pad, dil, groups, bias
	(2, 8, 12, 12, 16, 1, 1, 1, 0, 1, 1, False),
	(2, 8, 12, 12, 16, 3, 3, 1, 1, 1, 1, True),
	(2, 3, 20, 20, 8, 7, 7, 2, 3, 1, 1, False),       # ResNet conv1 geometry
	(2, 16, 11, 11, 32, 1, 1, 2, 0, 1, 1, False),     # strided 1x1 (first conv of a ResNet stage)
	(3, 6, 9, 10, 4, 2, 3, 1, 0, 1, 2, True),         # groups, non-square
	(2, 4, 13, 13, 6, 3, 3, 2, 1, 2, 1, False),       # dilation + stride
	(1, 1, 28, 28, 16, 3, 3, 1, 0, 1, 1, True),       # LeNet conv1
	(2, 16, 13, 13, 32, 4, 4, 1, 0, 1, 1, True),      # LeNet conv2
	(4, 64, 14, 14, 64, 3, 3, 1, 1, 1, 1, False),
	(2, 256, 7, 7, 512, 1, 1, 1, 0, 1, 1, False),
	(1, 130, 5, 5, 200, 3, 3, 1, 1, 1, 1, True),      # ragged: nothing a multiple of the tile sizes
	(2, 12, 9, 9, 12, 3, 3, 1, 1, 1, 4, False),       # 4 groups
	(1, 5, 1, 40, 7, 1, 5, 1, (0, 2), 1, 1, True),    # 1-d conv geometry (H = 1)
	(2, 3, 30, 30, 64, 7, 7, 2, 3, 1, 1, False),      # conv1 with 64 filters: dgrad runs the col2im scatter epilogue
	(2, 3, 17, 19, 64, 3, 3, 1, 1, 1, 1, False),      # VGG conv1_1 geometry (col2im dgrad, stride 1)
	(3, 1, 12, 12, 32, 3, 3, 1, 0, 2, 1, False),      # one input channel, dilation (col2im dgrad)
	(2, 48, 15, 15, 40, 3, 3, 2, 1, 1, 1, True),      # channel counts that are not multiples of 32, strided 3x3
	(2, 96, 9, 9, 64, 5, 5, 1, 2, 1, 1, False),       # 25 taps, channel-ordered k
	(2, 64, 10, 10, 96, 3, 3, 1, 1, 1, 2, True),      # groups with channel-ordered k
	(1, 32, 12, 224, 40, 3, 3, 1, 1, 1, 1, True),     # a 224-wide map: too wide for the halo kernel, per-tap gather
	(1, 32, 9, 126, 40, 3, 3, 1, 1, 1, 1, True),      # the widest map the stride-1 halo kernel takes (halo of 386 rows)
	(2, 40, 30, 31, 48, 3, 3, 1, 0, 1, 1, False),     # halo kernel without padding (output smaller than the input)
	(2, 32, 11, 13, 32, 3, 2, 1, (1, 0), 1, 1, False),    # non-square filter, asymmetric padding per axis
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_fwd_bwd(bnd, case):
	N, C, H, W, K, R, S, stride, pad, dil, groups, bias = case
	rng = np.random.RandomState(hash(case) % (2 ** 31))
	x = rng.randn(N, C, H, W).astype(np.float32)
	w = rng.randn(K, C // groups, R, S).astype(np.float32)
	b = rng.randn(K).astype(np.float32) if bias else None

	y = ops.conv2d(x, w, b, stride, pad, dil, groups)
	dy = rng.randn(*y.shape).astype(np.float32)

	dx_, dw_ = G(bnd, x), G(bnd, w)
	out = bnd.dnn.convNd(dx_, dw_, G(bnd, b) if bias else None, stride, pad, dil, groups, allocator=bnd.memoryPool)
	assert out.shape == y.shape
	assert relerr(out.get(), y) < REL_TC

	dgrad = bnd.dnn.convNdBackwardData(G(bnd, dy), dw_, None, dx_, stride, pad, dil, None, groups, allocator=bnd.memoryPool)
	assert relerr(dgrad.get(), ops.conv2d_bwd_data(dy, w, x.shape, None, stride, pad, dil, 0, groups)) < REL_TC

	w0 = rng.randn(*w.shape).astype(np.float32)
	b0 = rng.randn(K).astype(np.float32)
	wgrad, bgrad = G(bnd, w0), G(bnd, b0)
	res = bnd.dnn.convNdBackwardParams(dx_, G(bnd, dy), dw_, stride, pad, dil, groups, bias, False, wgrad, bgrad if bias else None,
									   0.5, 0.25, allocator=bnd.memoryPool)
	want = ops.conv2d_bwd_params(x, dy, w.shape, stride, pad, dil, groups, True, False, w0, b0, 0.5, 0.25)
	assert relerr(wgrad.get(), want[0]) < REL_TC
	if bias:
		assert res[1] is bgrad
		assert relerr(bgrad.get(), want[1]) < 1e-5


def test_conv2d_wgrad_overwrite_and_accumulate(bnd):
	# Module.backward default momentum = 0 overwrites, Sequential.backward default momentum = 1 accumulates
	rng = np.random.RandomState(11)
	x, dy = rng.randn(8, 32, 28, 28).astype(np.float32), rng.randn(8, 48, 28, 28).astype(np.float32)
	w = rng.randn(48, 32, 3, 3).astype(np.float32)
	want = ops.conv2d_bwd_params(x, dy, w.shape, 1, 1)
	wgrad = G(bnd, np.full(w.shape, np.nan, np.float32))
	bnd.dnn.convNdBackwardParams(G(bnd, x), G(bnd, dy), G(bnd, w), 1, 1, 1, 1, False, False, wgrad, None, 1.0, 0.0)
	assert relerr(wgrad.get(), want) < REL_TC          # NaNs in the old buffer must not leak at momentum = 0
	bnd.dnn.convNdBackwardParams(G(bnd, x), G(bnd, dy), G(bnd, w), 1, 1, 1, 1, False, False, wgrad, None, 1.0, 1.0)
	assert relerr(wgrad.get(), 2 * want) < REL_TC


@pytest.mark.parametrize("case", [(2, 6, 5, 5, 4, 3, 2, 1, 1, 1), (2, 4, 6, 7, 6, 4, 2, 1, 0, 2), (1, 8, 4, 4, 8, 2, 2, 0, 0, 1)])
def test_deconv2d(bnd, case):
	# Deconv forward = conv backward-data + bias, incl. postpad (Backend/Dnn.py:211-231; CuDnn.c:269-284)
	N, inmaps, H, W, outmaps, size, stride, pad, postpad, groups = case
	rng = np.random.RandomState(7)
	x = rng.randn(N, inmaps, H, W).astype(np.float32)
	w = rng.randn(inmaps, outmaps // groups, size, size).astype(np.float32)
	b = rng.randn(outmaps).astype(np.float32)

	y = ops.conv2d_bwd_data(x, w, None, b, stride, pad, 1, postpad, groups)
	out = bnd.dnn.convNdBackwardData(G(bnd, x), G(bnd, w), G(bnd, b), None, stride, pad, 1, postpad, groups, allocator=bnd.memoryPool)
	assert out.shape == y.shape
	assert relerr(out.get(), y) < REL_TC

	g = rng.randn(*y.shape).astype(np.float32)
	dx = bnd.dnn.convNd(G(bnd, g), G(bnd, w), None, stride, pad, 1, groups)
	assert relerr(dx.get(), ops.conv2d(g, w, None, stride, pad, 1, groups)) < REL_TC

	wgrad, bgrad = bnd.dnn.convNdBackwardParams(G(bnd, g), G(bnd, x), G(bnd, w), stride, pad, 1, groups, True, True)
	want = ops.conv2d_bwd_params(g, x, w.shape, stride, pad, 1, groups, True, True)
	assert relerr(wgrad.get(), want[0]) < REL_TC and relerr(bgrad.get(), want[1]) < 1e-5


# 16-bit storage (the reference has fp16; bf16 is added with the same rules): operands are read exactly by the
# tensor core (kind::f16), fp32 accumulation, outputs rounded once to the storage type.  Tolerance = a few storage ulps.
CONV16_CASES = [
	# N, C, H, W, K, R, S, stride, pad, dil, groups, bias
	(2, 64, 14, 14, 96, 3, 3, 1, 1, 1, 1, True),       # channel-ordered k, 64-element k-blocks
	(2, 48, 15, 15, 80, 3, 3, 2, 1, 1, 1, False),      # strided 3x3 (parity classes), ragged channel counts
	(3, 128, 7, 7, 64, 1, 1, 1, 0, 1, 1, False),       # 1x1
	(2, 64, 12, 12, 128, 1, 1, 2, 0, 1, 1, False),     # strided 1x1
	(2, 3, 20, 20, 64, 3, 3, 1, 1, 1, 1, True),        # 3 input channels: general gather for fprop, >= 48 filters for dgrad
	(2, 96, 9, 9, 64, 3, 3, 1, 1, 1, 2, True),         # groups
]


def _dtypes16():
	from puzzlelib_b200.driver import bfloat16
	return [np.dtype(np.float16)] + ([bfloat16] if bfloat16 is not None else [])


@pytest.mark.parametrize("case", CONV16_CASES)
@pytest.mark.parametrize("dtname", ["float16", "bfloat16"])
def test_conv2d_16bit(bnd, case, dtname):
	from puzzlelib_b200.driver import bfloat16
	dt = np.dtype(np.float16) if dtname == "float16" else bfloat16
	if dt is None:
		pytest.skip("ml_dtypes.bfloat16 is not available")
	tol = 4e-3 if dtname == "float16" else 2e-2
	N, C, H, W, K, R, S, stride, pad, dil, groups, bias = case
	rng = np.random.RandomState(hash(case) % (2 ** 31))
	x = rng.randn(N, C, H, W).astype(dt)
	w = (rng.randn(K, C // groups, R, S) * 0.2).astype(dt)
	b = rng.randn(K).astype(dt) if bias else None
	xf, wf = x.astype(np.float64), w.astype(np.float64)

	y = ops.conv2d(xf, wf, b.astype(np.float64) if bias else None, stride, pad, dil, groups)
	dy = rng.randn(*y.shape).astype(dt)
	dyf = dy.astype(np.float64)

	dx_, dw_ = G(bnd, x), G(bnd, w)
	out = bnd.dnn.convNd(dx_, dw_, G(bnd, b) if bias else None, stride, pad, dil, groups, allocator=bnd.memoryPool)
	assert out.dtype == dt and out.shape == y.shape
	assert relerr(out.get().astype(np.float64), y) < tol

	dgrad = bnd.dnn.convNdBackwardData(G(bnd, dy), dw_, None, dx_, stride, pad, dil, None, groups, allocator=bnd.memoryPool)
	assert dgrad.dtype == dt
	assert relerr(dgrad.get().astype(np.float64), ops.conv2d_bwd_data(dyf, wf, x.shape, None, stride, pad, dil, 0, groups)) < tol

	w0, b0 = rng.randn(*w.shape).astype(dt), rng.randn(K).astype(dt)
	wgrad, bgrad = G(bnd, w0), G(bnd, b0)
	bnd.dnn.convNdBackwardParams(dx_, G(bnd, dy), dw_, stride, pad, dil, groups, bias, False, wgrad, bgrad if bias else None, 0.5, 0.25,
								 allocator=bnd.memoryPool)
	want = ops.conv2d_bwd_params(xf, dyf, w.shape, stride, pad, dil, groups, True, False, w0.astype(np.float64), b0.astype(np.float64),
								 0.5, 0.25)
	assert relerr(wgrad.get().astype(np.float64), want[0]) < tol
	if bias:
		assert relerr(bgrad.get().astype(np.float64), want[1]) < tol


@pytest.mark.parametrize("case", [(64, 1000, 2048, 0, 0), (128, 96, 300, 0, 1), (70, 130, 64, 1, 0), (33, 17, 4, 0, 0), (16, 4096, 8192, 0, 0)])
@pytest.mark.parametrize("dtname", ["float16", "bfloat16"])
def test_gemm_16bit(bnd, case, dtname):
	from puzzlelib_b200.driver import bfloat16
	dt = np.dtype(np.float16) if dtname == "float16" else bfloat16
	if dt is None:
		pytest.skip("ml_dtypes.bfloat16 is not available")
	tol = 4e-3 if dtname == "float16" else 2e-2
	M, N, K, ta, tb = case
	rng = np.random.RandomState(M * 31 + N)
	A = (rng.randn(*((K, M) if ta else (M, K))) * 0.3).astype(dt)
	B = (rng.randn(*((N, K) if tb else (K, N))) * 0.3).astype(dt)
	C0 = rng.randn(M, N).astype(dt)
	Af, Bf = A.astype(np.float64), B.astype(np.float64)

	out = bnd.blas.gemm(G(bnd, A), G(bnd, B), None, bool(ta), bool(tb), allocator=bnd.memoryPool)
	assert out.dtype == dt
	assert relerr(out.get().astype(np.float64), ops.gemm(Af, Bf, None, ta, tb)) < tol

	acc = G(bnd, C0)
	bnd.blas.gemm(G(bnd, A), G(bnd, B), acc, bool(ta), bool(tb), 0.5, 0.75)
	assert relerr(acc.get().astype(np.float64), ops.gemm(Af, Bf, C0.astype(np.float64), ta, tb, 0.5, 0.75)) < tol


def test_conv2d_argument_errors(bnd):
	x = bnd.GPUArray.zeros((2, 4, 8, 8), np.float32)
	w = bnd.GPUArray.zeros((6, 3, 3, 3), np.float32)
	with pytest.raises(ValueError, match="input maps"):
		bnd.dnn.convNd(x, w)
	w = bnd.GPUArray.zeros((6, 4, 3, 3), np.float32)
	with pytest.raises(ValueError, match="output gpuarray"):
		bnd.dnn.convNd(x, w, out=bnd.GPUArray.zeros((2, 6, 5, 5), np.float32))
	with pytest.raises(ValueError):
		bnd.dnn.convNd(x, w, stride=(1, 2, 3))
	with pytest.raises(ValueError, match="map size"):
		bnd.dnn.convNd(bnd.GPUArray.zeros((2, 4, 2, 2), np.float32), w)


# ================================================================================================ GEMM
GEMM_CASES = [(128, 128, 32, 0, 0), (64, 1000, 2048, 0, 0), (64, 2048, 1000, 0, 1), (2048, 1000, 64, 1, 0), (100, 200, 77, 0, 0),
			  (300, 70, 129, 1, 0), (33, 17, 5, 0, 1), (1, 10, 1024, 0, 0), (64, 1024, 800, 0, 0), (128, 4096, 4096, 0, 0)]


@pytest.mark.parametrize("case", GEMM_CASES)
def test_gemm(bnd, case):
	M, N, K, ta, tb = case
	rng = np.random.RandomState(M * 31 + N)
	A = rng.randn(*((K, M) if ta else (M, K))).astype(np.float32)
	B = rng.randn(*((N, K) if tb else (K, N))).astype(np.float32)
	C0 = rng.randn(M, N).astype(np.float32)

	out = bnd.blas.gemm(G(bnd, A), G(bnd, B), None, bool(ta), bool(tb), allocator=bnd.memoryPool)
	assert relerr(out.get(), ops.gemm(A, B, None, ta, tb)) < REL_TC

	acc = G(bnd, C0)
	res = bnd.blas.gemm(G(bnd, A), G(bnd, B), acc, bool(ta), bool(tb), 0.5, 0.75)
	assert res is acc
	assert relerr(acc.get(), ops.gemm(A, B, C0, ta, tb, 0.5, 0.75)) < REL_TC


@pytest.mark.parametrize("formats", [(0, 0, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0)])
def test_gemm_batched(bnd, formats):
	# reference tests: Cuda/Wrappers/CuBlas.py:60-189 (gbp / bgp operand and output layouts, one transposed operand at a time)
	fa, fb, fo = formats
	rng = np.random.RandomState(17)
	groups, M, K, N = 3, 20, 33, 12

	def layout(mats, fmt):          # list of per-group matrices -> gbp or bgp tensor
		t = np.stack(mats)
		return np.ascontiguousarray(t if fmt == 0 else t.transpose(1, 0, 2)).astype(np.float32)

	def per_group(t, fmt):
		return [t[g] if fmt == 0 else t[:, g, :] for g in range(groups)]

	As = [rng.randn(M, K) for _ in range(groups)]
	for ta, tb in ((False, False), (True, False), (False, True)):
		Bs = [rng.randn(*((N, K) if tb else ((M, N) if ta else (K, N)))) for _ in range(groups)]
		out = bnd.blas.gemmBatched(G(bnd, layout(As, fa)), G(bnd, layout(Bs, fb)), fa, fb, fo, ta, tb, allocator=bnd.memoryPool)
		want = [(a.T if ta else a) @ (b.T if tb else b) for a, b in zip(As, Bs)]
		got = per_group(out.get(), fo)
		for g in range(groups):
			assert relerr(got[g], want[g]) < REL_TC


def test_gemm_bias_epilogue_and_errors(bnd):
	rng = np.random.RandomState(3)
	A, B, b = rng.randn(64, 800).astype(np.float32), rng.randn(800, 1024).astype(np.float32), rng.randn(1024).astype(np.float32)
	out = bnd.blas.gemmBias(G(bnd, A), G(bnd, B), G(bnd, b))
	assert relerr(out.get(), ops.gemm(A, B) + b) < REL_TC
	with pytest.raises(ValueError):
		bnd.blas.gemm(G(bnd, A), G(bnd, A))
	with pytest.raises(ValueError):
		bnd.blas.gemm(G(bnd, A), G(bnd, B), transpA=True, transpB=True)
	with pytest.raises(ValueError):            # a bias that does not have one entry per output column (Linear(transpose=True)'s quirk)
		bnd.blas.gemmBias(G(bnd, A), G(bnd, B), G(bnd, b[:800]))
	with pytest.raises(ValueError):
		bnd.blas.gemmBias(G(bnd, A), G(bnd, B), G(bnd, b.astype(np.float16)))


# ================================================================================================ batch norm
@pytest.mark.parametrize("shape", [(4, 5, 2, 3), (16, 64, 14, 14), (8, 3, 33, 35), (2, 7, 1, 1), (32, 256, 7, 7), (3, 2, 112, 112),
								   (64, 8, 55, 55), (5, 3, 7, 7), (64, 12, 28, 28), (1, 4, 9, 9), (13, 300, 3, 5), (64, 2, 112, 112)])
def test_batchnorm_train_bwd_infer(bnd, shape):
	rng = np.random.RandomState(sum(shape))
	N, C = shape[:2]
	x = (rng.randn(*shape) * 2 + 0.5).astype(np.float32)
	scale, bias = rng.randn(C).astype(np.float32), rng.randn(C).astype(np.float32)
	mean0, var0 = rng.randn(C).astype(np.float32), (1 + rng.randn(C) ** 2).astype(np.float32)
	factor = 0.3

	y, mu, inv, newmean, newvar = ops.batchnorm_train(x, scale, bias, mean0, var0, 1e-5, factor)
	mean, var = G(bnd, mean0), G(bnd, var0)
	out, savemean, saveinvvar = bnd.dnn.batchNormNd(G(bnd, x), mean, var, G(bnd, scale), G(bnd, bias), 1e-5, factor, False,
													allocator=bnd.memoryPool)
	assert np.allclose(out.get(), y, atol=2e-5, rtol=1e-5)
	assert np.allclose(savemean.get(), mu, atol=ATOL32) and np.allclose(saveinvvar.get(), inv, rtol=1e-5, atol=ATOL32)
	assert np.allclose(mean.get(), newmean, atol=ATOL32)
	assert np.allclose(var.get(), newvar, rtol=1e-5, atol=ATOL32)      # unbiased running variance (cuDNN), see oracle

	dy = rng.randn(*shape).astype(np.float32)
	dx, dscale, dbias = ops.batchnorm_bwd(x, dy, scale, mu, inv)
	ingrad, scalegrad, bgrad = bnd.dnn.batchNormNdBackward(G(bnd, dy), G(bnd, x), G(bnd, scale), savemean, saveinvvar, 1e-5,
														   allocator=bnd.memoryPool)
	# dx is a difference of O(|dy| * |scale| * invstd) terms: hold it to 1e-5 of that scale (with 2 samples per channel it
	# cancels to ~0 and a bound relative to max|dx| would be meaningless)
	bar = 1e-5 * max(1.0, float(np.abs(dy).max() * np.abs(scale).max() * inv.max()))
	assert np.abs(ingrad.get() - dx).max() < bar
	assert relerr(scalegrad.get(), dscale) < 1e-5 and relerr(bgrad.get(), dbias) < 1e-5

	data = G(bnd, x)
	res = bnd.dnn.batchNormNd(data, G(bnd, mean0), G(bnd, var0), G(bnd, scale), G(bnd, bias), 1e-5, 0, True, out=data)
	assert res is data                                                   # in-place inference (BatchNormND inplace flag)
	assert np.allclose(data.get(), ops.batchnorm_infer(x, scale, bias, mean0, var0), atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("dtname", ["float16", "bfloat16"])
@pytest.mark.parametrize("shape", [(8, 6, 7, 7), (16, 32, 14, 14), (4, 5, 55, 55), (64, 4, 28, 28)])
def test_batchnorm_16bit(bnd, shape, dtname):
	"""16-bit activations, fp32 parameters and statistics (CuDnnNorm.c:118-121); odd planes are not 16-byte aligned"""
	from puzzlelib_b200.driver import bfloat16
	dt = np.dtype(np.float16) if dtname == "float16" else bfloat16
	tol = 4e-3 if dtname == "float16" else 3e-2
	rng = np.random.RandomState(sum(shape))
	C = shape[1]
	x = (rng.randn(*shape) + 0.5).astype(dt)
	dy = rng.randn(*shape).astype(dt)
	scale, bias = (1 + 0.2 * rng.randn(C)).astype(np.float32), rng.randn(C).astype(np.float32)
	zero, one = np.zeros(C, np.float32), np.ones(C, np.float32)
	x32, dy32 = x.astype(np.float32), dy.astype(np.float32)

	y, mu, inv, newmean, newvar = ops.batchnorm_train(x32, scale, bias, zero, one, 1e-5, 1.0)
	mean, var = G(bnd, zero), G(bnd, one)
	out, savemean, saveinvvar = bnd.dnn.batchNormNd(G(bnd, x), mean, var, G(bnd, scale), G(bnd, bias), 1e-5, 1.0, False)
	assert out.dtype == dt and savemean.dtype == np.float32
	assert relerr(out.get().astype(np.float32), y) < tol
	assert np.allclose(savemean.get(), mu, atol=1e-5) and np.allclose(saveinvvar.get(), inv, rtol=1e-5)
	assert np.allclose(mean.get(), newmean, atol=1e-5) and np.allclose(var.get(), newvar, rtol=1e-5)

	dx, dscale, dbias = ops.batchnorm_bwd(x32, dy32, scale, mu, inv)
	ingrad, scalegrad, bgrad = bnd.dnn.batchNormNdBackward(G(bnd, dy), G(bnd, x), G(bnd, scale), savemean, saveinvvar, 1e-5)
	assert ingrad.dtype == dt and scalegrad.dtype == np.float32
	bar = tol * max(1.0, float(np.abs(dy32).max() * np.abs(scale).max() * inv.max()))
	assert np.abs(ingrad.get().astype(np.float32) - dx).max() < bar
	assert relerr(scalegrad.get(), dscale) < 1e-4 and relerr(bgrad.get(), dbias) < 1e-4


def test_batchnorm_is_deterministic_and_works_in_place(bnd):
	"""The cluster kernels reduce statistics through fixed trees (no atomics): two runs give the same bits.  Train-mode in-place
	normalisation (dnn.batchNormNd(..., out=data)) gives the bits of the out-of-place call.  (Planes too large for the
	shared-memory stash -- above ~100 KB per CTA in a cluster of 8, e.g. the backward pass at 64 x C x 55 x 55 -- take the
	two-kernel path, which accumulates with fp32 atomics and is reproducible to rounding only.)"""
	rng = np.random.RandomState(11)
	for shape in ((16, 16, 55, 55), (32, 64, 7, 7), (6, 10, 28, 28), (64, 128, 28, 28)):
		C = shape[1]
		x, dy = rng.randn(*shape).astype(np.float32), rng.randn(*shape).astype(np.float32)
		scale, bias = rng.randn(C).astype(np.float32), rng.randn(C).astype(np.float32)
		runs = []
		for _ in range(2):
			mean, var = G(bnd, np.zeros(C, np.float32)), G(bnd, np.ones(C, np.float32))
			out, sm, siv = bnd.dnn.batchNormNd(G(bnd, x), mean, var, G(bnd, scale), G(bnd, bias), 1e-5, 0.5, False)
			dx, ds, db = bnd.dnn.batchNormNdBackward(G(bnd, dy), G(bnd, x), G(bnd, scale), sm, siv, 1e-5)
			runs.append([a.get() for a in (out, sm, siv, mean, var, dx, ds, db)])
		for a, b in zip(*runs):
			assert np.array_equal(a, b)

		data = G(bnd, x)
		mean, var = G(bnd, np.zeros(C, np.float32)), G(bnd, np.ones(C, np.float32))
		res, _, _ = bnd.dnn.batchNormNd(data, mean, var, G(bnd, scale), G(bnd, bias), 1e-5, 0.5, False, out=data)
		assert res is data and np.array_equal(data.get(), runs[0][0])


def test_batchnorm_large_mean_is_stable(bnd):
	# variance of data with mean >> std: a naive E[x^2] - E[x]^2 in fp32 loses all digits here
	rng = np.random.RandomState(9)
	x = (1000.0 + rng.randn(32, 4, 20, 20)).astype(np.float32)
	one, zero = np.ones(4, np.float32), np.zeros(4, np.float32)
	_, mu, inv, _, _ = ops.batchnorm_train(x, one, zero, zero, one)
	_, savemean, saveinvvar = bnd.dnn.batchNormNd(G(bnd, x), G(bnd, zero), G(bnd, one), G(bnd, one), G(bnd, zero))
	assert np.allclose(savemean.get(), mu, rtol=1e-6)
	assert np.allclose(saveinvvar.get(), inv, rtol=2e-3)


def test_instancenorm(bnd):
	rng = np.random.RandomState(5)
	x = rng.randn(3, 4, 6, 5).astype(np.float32)
	scale, bias = rng.randn(4).astype(np.float32), rng.randn(4).astype(np.float32)
	out, savemean, saveinvvar, extscale = bnd.instanceNorm2d(G(bnd, x), G(bnd, scale), G(bnd, bias), 1e-5)
	mu, var = x.mean(axis=(2, 3), keepdims=True), x.var(axis=(2, 3), keepdims=True)
	want = (x - mu) / np.sqrt(var + 1e-5) * scale.reshape(1, 4, 1, 1) + bias.reshape(1, 4, 1, 1)
	assert np.allclose(out.get(), want, atol=2e-5)
	g = rng.randn(*x.shape).astype(np.float32)
	ingrad, scalegrad, biasgrad = bnd.instanceNorm2dBackward(G(bnd, g), G(bnd, x), extscale, savemean, saveinvvar, 1e-5, True)
	xhat = (x - mu) / np.sqrt(var + 1e-5)
	assert np.allclose(biasgrad.get(), g.sum(axis=(0, 2, 3)), atol=1e-4)
	assert np.allclose(scalegrad.get(), (g * xhat).sum(axis=(0, 2, 3)), atol=1e-4)


# ================================================================================================ pooling
POOL_CASES = [(2, 3, 8, 8, 2, 2, 0), (2, 2, 9, 9, 3, 2, 0), (3, 2, 6, 6, 3, 2, 1), (2, 4, 112, 112, 3, 2, 0), (2, 5, 7, 7, 7, 1, 0),
			  (1, 2, 10, 13, (2, 3), (2, 1), (1, 1)), (2, 3, 5, 5, 3, 1, 1)]


@pytest.mark.parametrize("case", POOL_CASES)
@pytest.mark.parametrize("mode", ["max", "avgWithPad", "avgNoPad"])
def test_pool2d(bnd, case, mode):
	N, C, H, W, size, stride, pad = case
	rng = np.random.RandomState(H * 7 + W)
	x = rng.randn(N, C, H, W).astype(np.float32)
	code = {"max": bnd.PoolMode.max, "avgWithPad": bnd.PoolMode.avgWithPad, "avgNoPad": bnd.PoolMode.avgNoPad}[mode].value

	y = ops.pool2d(x, size, stride, pad, mode)
	out = bnd.dnn.poolNd(G(bnd, x), size, stride, pad, code, allocator=bnd.memoryPool)
	assert out.shape == y.shape
	if mode == "max":
		assert np.array_equal(out.get(), y.astype(np.float32))
	else:
		assert np.allclose(out.get(), y, atol=ATOL32)

	dy = rng.randn(*y.shape).astype(np.float32)
	dx = bnd.dnn.poolNdBackward(G(bnd, dy), G(bnd, x), out, size, stride, pad, code, allocator=bnd.memoryPool)
	assert np.allclose(dx.get(), ops.pool2d_bwd(x, out.get(), dy, size, stride, pad, mode), atol=ATOL32)


def test_pool2d_max_backward_ties_go_to_the_first_maximum(bnd):
	# post-ReLU windows full of zeros: route like the mask kernel does (first maximum), deterministically
	x = np.zeros((1, 2, 6, 6), np.float32)
	x[0, 1, 2, 2] = x[0, 1, 2, 3] = 3.0
	dy = np.arange(1, 19, dtype=np.float32).reshape(1, 2, 3, 3)
	out = bnd.dnn.poolNd(G(bnd, x), 2, 2, 0, 0)
	dx = bnd.dnn.poolNdBackward(G(bnd, dy), G(bnd, x), out, 2, 2, 0, 0).get()
	assert np.array_equal(dx, ops.pool2d_bwd(x, out.get(), dy, 2, 2, 0, "max").astype(np.float32))
	assert dx[0, 0, 0, 0] == 1.0 and dx[0, 0, 0, 1] == 0.0 and dx[0, 1, 2, 2] == 14.0 and dx[0, 1, 2, 3] == 0.0


@pytest.mark.parametrize("case", POOL_CASES)
def test_maxpool2d_mask_is_bit_exact(bnd, case):
	N, C, H, W, size, stride, pad = case
	size, stride, pad = ops._pair(size), ops._pair(stride), ops._pair(pad)
	rng = np.random.RandomState(H + W)
	x = rng.randn(N, C, H, W).astype(np.float32)
	x[rng.rand(*x.shape) < 0.4] = 0.0                 # many exact ties, as after a ReLU
	x = np.maximum(x, 0.0)

	y, mask = ops.maxpool2d_mask(x, size, stride, pad)
	out, gmask = bnd.poolmod.maxpool2d(G(bnd, x), size, stride, pad, allocator=bnd.memoryPool)
	assert gmask.dtype == np.int32
	assert (mask == gmask.get()).all()                # the reference's own assertion, Cuda/Kernels/Pool.py:263
	assert np.array_equal(out.get(), y)

	dy = rng.randn(*y.shape).astype(np.float32)
	dx = bnd.poolmod.maxpool2dBackward(G(bnd, dy), x.shape, gmask, size, stride, pad, allocator=bnd.memoryPool)
	assert np.array_equal(dx.get(), ops.maxpool2d_mask_bwd(dy, x.shape, mask, size, stride, pad))

	if stride[0] >= size[0] and stride[1] >= size[1] and pad == (0, 0):              # unpool needs non-overlapping windows to be a function
		up = bnd.poolmod.maxunpool2d(out, x.shape, gmask)
		assert np.array_equal(up.get(), ops.maxunpool2d(y, x.shape, mask))
		back = bnd.poolmod.maxunpool2dBackward(up, y.shape, gmask)
		assert np.array_equal(back.get(), y)


# ================================================================================================ softmax
@pytest.mark.parametrize("shape", [(5, 8, 2, 3), (64, 1000, 1, 1), (3, 1500, 1, 1), (7, 10, 1, 1), (2, 21, 17, 9), (1, 1, 1, 1)])
def test_softmax(bnd, shape):
	rng = np.random.RandomState(shape[1])
	x = (rng.randn(*shape) * 3).astype(np.float32)
	y = ops.softmax(x)
	out = bnd.dnn.softmaxNd(G(bnd, x), bnd.SoftMaxMode.spatial.value, allocator=bnd.memoryPool)
	assert np.allclose(out.get(), y, atol=1e-6, rtol=1e-5)
	dy = rng.randn(*shape).astype(np.float32)
	dx = bnd.dnn.softmaxNdBackward(G(bnd, dy), out, allocator=bnd.memoryPool)
	assert np.allclose(dx.get(), ops.softmax_bwd(out.get(), dy), atol=1e-6, rtol=1e-5)

	out = bnd.dnn.softmaxNd(G(bnd, x), bnd.SoftMaxMode.perActivation.value)
	assert np.allclose(out.get(), ops.softmax(x, "perActivation"), atol=1e-6, rtol=1e-5)


def test_softmax_is_max_subtracted(bnd):
	x = np.array([[1000.0, 1001.0, 999.0], [-1e4, 0.0, -1e4]], np.float32).reshape(2, 3, 1, 1)
	out = bnd.dnn.softmaxNd(G(bnd, x)).get()
	assert np.isfinite(out).all() and np.allclose(out, ops.softmax(x), atol=1e-6)


# ================================================================================================ elementwise
ACTS = [("sigmoid", ()), ("tanh", ()), ("relu", ()), ("leakyRelu", (0.01, )), ("elu", (1.0, )), ("softPlus", ()), ("clip", (0.0, 6.0)),
		("gelu", ())]


@pytest.mark.parametrize("kind,args", ACTS)
@pytest.mark.parametrize("n", [1, 7, 1000, 4099, 1 << 20])
def test_activations(bnd, kind, args, n):
	rng = np.random.RandomState(n % 1000)
	x = (rng.randn(n) * 3).astype(np.float32)
	x[::5] = 0.0
	g = rng.randn(n).astype(np.float32)

	out = bnd.GPUArray.empty((n, ), np.float32)
	getattr(bnd, "%sKer" % kind)(np.float32)(out, G(bnd, x), *args)
	y = ops.activation(kind, x, *args)
	# the reference compiles these kernels with -use_fast_math (fast expf / logf / division): 1e-5 absolute is ITS bar
	tol = {"tanh": 1e-3, "gelu": 1e-4}.get(kind, 2e-5)      # fast-math tanhf is the MUFU.TANH approximation (2^-11)
	assert np.allclose(out.get(), y, atol=tol, rtol=tol)

	ref = x if kind == "gelu" else out.get()
	ingrad = bnd.GPUArray.empty((n, ), np.float32)
	getattr(bnd, "%sDerKer" % kind)(np.float32)(ingrad, G(bnd, g), G(bnd, ref), *args)
	assert np.allclose(ingrad.get(), ops.activation_bwd(kind, g, ref, *args), atol=tol, rtol=tol)


def test_relu_is_exact_and_works_in_place_and_on_unaligned_views(bnd):
	rng = np.random.RandomState(1)
	x = rng.randn(3, 1001).astype(np.float32)
	d = G(bnd, x)
	bnd.reluKer(np.float32)(d, d)
	assert np.array_equal(d.get(), x * (x > 0))
	flat = G(bnd, x).ravel()
	view = flat[3:2002]                              # 12-byte offset: exercises the head / tail peeling
	out = bnd.GPUArray.zeros((2000, ), np.float32)[1:]
	bnd.reluKer(np.float32)(out, view)
	xs = x.ravel()[3:2002]
	assert np.array_equal(out.get(), xs * (xs > 0))


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_16bit_elementwise(bnd, dtype):
	from puzzlelib_b200.driver import bfloat16
	dt = np.dtype(np.float16) if dtype == "float16" else bfloat16
	rng = np.random.RandomState(2)
	x32 = rng.randn(5000).astype(np.float32)
	x = x32.astype(dt)
	d = G(bnd, x)
	out = bnd.GPUArray.empty(x.shape, dt)
	bnd.reluKer(dt)(out, d)
	xf = x.astype(np.float32)
	assert np.array_equal(out.get().astype(np.float32), xf * (xf > 0))
	back = G(bnd, x32).astype(dt)
	assert np.array_equal(back.get().view(np.uint16), x.view(np.uint16))      # round-to-nearest-even like numpy
	assert np.array_equal(back.astype(np.float32).get(), xf)


def test_blas1_kernels(bnd):
	rng = np.random.RandomState(4)
	n = 100003
	x, y = rng.randn(n).astype(np.float32), rng.randn(n).astype(np.float32)
	dy = G(bnd, y)
	bnd.toVectorAddVectorKer(np.float32)(dy, G(bnd, x), 0.3)
	assert np.allclose(dy.get(), y + np.float32(0.3) * x, atol=1e-6)
	out = bnd.GPUArray.empty((n, ), np.float32)
	bnd.addKer(np.float32)(out, G(bnd, x), 0.25, G(bnd, y), -1.5)
	assert np.allclose(out.get(), 0.25 * x - 1.5 * y, atol=1e-6)
	bnd.linearKer(np.float32)(out, G(bnd, x), 2.0, -1.0)
	assert np.allclose(out.get(), 2 * x - 1, atol=1e-6)
	bnd.mulKer(np.float32)(out, G(bnd, x), G(bnd, y))
	assert np.array_equal(out.get(), x * y)
	bnd.add2Ker(np.float32)(out, G(bnd, x), G(bnd, y))
	assert np.array_equal(out.get(), (np.float32(0) + x) + y)                # same bits as fill(0) + 2 axpy (Add.py:15-23)
	p, m = rng.randn(n).astype(np.float32), rng.randn(n).astype(np.float32)
	dp, dm = G(bnd, p), G(bnd, m)
	bnd.classicMomSGDKer(np.float32)(dp, G(bnd, x), dm, 0.1, 0.9)
	wp, wm = ops.sgd_momentum(p, x, m, 0.1, 0.9)
	assert np.allclose(dp.get(), wp, atol=1e-6) and np.allclose(dm.get(), wm, atol=1e-6)


def test_matvec_helpers(bnd):
	rng = np.random.RandomState(6)
	mat, vec = rng.randn(37, 130).astype(np.float32), rng.randn(130).astype(np.float32)
	out = bnd.matmod.addVecToMat(G(bnd, vec), G(bnd, mat), axis=1)
	assert np.array_equal(out.get(), mat + vec)
	col = rng.randn(37).astype(np.float32)
	assert np.array_equal(bnd.matmod.addVecToMat(G(bnd, col), G(bnd, mat), axis=0).get(), mat + col[:, None])
	small = rng.randn(13).astype(np.float32)
	assert np.array_equal(bnd.matmod.addVecToMat(G(bnd, small), G(bnd, mat), axis=1).get(), mat + np.tile(small, 10))

	assert np.allclose(bnd.matmod.matsum(G(bnd, mat), 0).get(), mat.sum(0), atol=1e-4)
	assert np.allclose(bnd.matmod.matsum(G(bnd, mat), 1).get(), mat.sum(1), atol=1e-4)
	acc = rng.randn(130).astype(np.float32)
	dacc = G(bnd, acc)
	bnd.matmod.matsum(G(bnd, mat), 0, dacc, 0.5, 2.0)
	assert np.allclose(dacc.get(), 2 * acc + 0.5 * mat.sum(0), atol=1e-4)
	t3 = rng.randn(4, 9, 11).astype(np.float32)
	assert np.allclose(bnd.matmod.matsum(G(bnd, t3), 1).get(), t3.sum(1), atol=1e-4)

	logits = rng.randn(64, 1000).astype(np.float32)
	logits[3, 17] = logits[3, 900] = 50.0             # tie: first occurrence wins
	assert np.array_equal(bnd.matmod.argmax(G(bnd, logits), 1).get(), ops.argmax(logits, 1))
	assert np.array_equal(bnd.matmod.argmax(G(bnd, t3), 1).get(), ops.argmax(t3, 1))
	assert np.array_equal(bnd.matmod.argmin(G(bnd, t3), 0).get(), np.argmin(t3, 0).astype(np.int32))


# ================================================================================================ GPUArray / memory
def test_gpuarray_roundtrip_views_fill_and_arith(bnd):
	rng = np.random.RandomState(8)
	h = rng.randn(4, 5, 6).astype(np.float32)
	a = G(bnd, h)
	assert np.array_equal(a.get(), h)
	assert np.array_equal(a[1:3].get(), h[1:3]) and np.array_equal(a[:, 2].get(), h[:, 2])
	assert np.array_equal(a[:, 1:4, 2:5].get(), h[:, 1:4, 2:5])       # 2 discontiguous axes -> pitched copies
	a[:, 2] = np.zeros((4, 6), np.float32)
	h[:, 2] = 0
	assert np.array_equal(a.get(), h)
	assert np.array_equal(a[:, 1:3].copy().get(), h[:, 1:3])
	assert np.array_equal(a.reshape(20, 6).get(), h.reshape(20, 6))

	for dtype, val in ((np.float32, 1.5), (np.float16, -2.0), (np.int32, 7), (np.int8, -3), (np.float64, 3.25), (np.int64, 1 << 40)):
		f = bnd.GPUArray.empty((1003, ), dtype)
		f.fill(val)
		assert (f.get() == np.array(val, dtype)).all()
	b = G(bnd, h)
	assert np.array_equal((a + b).get(), h + h) and np.array_equal((a * b).get(), h * h)
	a += b
	assert np.array_equal(a.get(), h + h)
	assert float(b.max().get()) == h.max() and float(b.min().get()) == h.min()
	with pytest.raises(ValueError):
		a + G(bnd, h[:2])


def test_memory_pool_reuses_blocks(bnd):
	pool = bnd.Driver.MemoryPool()
	a = bnd.GPUArray.empty((1000, ), np.float32, allocator=pool)
	ptr = a.ptr
	del a
	stats = pool.getStats()
	assert stats["heldBlocks"] == 1 and stats["activeBlocks"] == 0 and stats["heldBytes"] == pool.allocSize(4000)
	b = bnd.GPUArray.empty((1001, ), np.float32, allocator=pool)      # same size class -> same block, no cudaMalloc
	assert b.ptr == ptr
	del b
	pool.freeHeld()
	assert pool.getStats()["heldBlocks"] == 0


def test_concatenate_split_tile_sharedarray(bnd):
	rng = np.random.RandomState(10)
	a, b = rng.randn(3, 4, 5).astype(np.float32), rng.randn(3, 2, 5).astype(np.float32)
	cat = bnd.concatenate([G(bnd, a), G(bnd, b)], axis=1)
	assert np.array_equal(cat.get(), np.concatenate([a, b], axis=1))
	parts = bnd.split(cat, [4, 2], axis=1)
	assert np.array_equal(parts[0].get(), a) and np.array_equal(parts[1].get(), b)
	assert np.array_equal(bnd.tile(G(bnd, a), 3, axis=0).get(), np.tile(a, (3, 1, 1)))

	shared = bnd.SharedArray(np.float32, allocator=bnd.memoryPool)
	shared.register((3, 5), np.float32, "w")
	shared.register((7, ), np.float32, "b")
	shared.build()
	assert shared.ary.size == 16 + 8 and shared["b"].ptr - shared["w"].ptr == 64      # 16-byte aligned blocks
	shared["w"].set(np.ones((3, 5), np.float32))
	shared["b"].fill(2.0)
	flat = shared.ary.get()
	assert flat[:15].sum() == 15 and (flat[16:23] == 2).all()


# ================================================================================================ recurrent layers
def _rnn_host_params(params):
	return {k: v.get().astype(np.float64) for k, v in params.items()}


@pytest.mark.parametrize("case", [(4, 2, 4, 2, 1), (7, 5, 33, 20, 1), (6, 16, 64, 96, 2), (12, 64, 128, 128, 1)])
def test_lstm_forward_backward(bnd, case):
	# reference test: Cuda/Wrappers/CuDnnRnn.py:178-300 (lstmTest); GEMMs run as TF32 on the tensor cores
	T, B, insz, H, layers = case
	rng = np.random.RandomState(T * 100 + H)
	rnn, W, params = bnd.createRnn(insz, H, np.float32, layers=layers, mode=bnd.RNNMode.lstm)
	assert W.shape == (sum(4 * H * ((insz if l == 0 else H) + H) + 8 * H for l in range(layers)), )
	W.set((rng.randn(*W.shape) * 0.2).astype(np.float32))
	host = [_rnn_host_params(p) for p in params]

	x = rng.randn(T, B, insz).astype(np.float32)
	dy = rng.randn(T, B, H).astype(np.float32)
	out, reserve = rnn.forward(G(bnd, x), W, allocator=bnd.memoryPool)

	acts, caches = [x.astype(np.float64)], []
	for l in range(layers):
		o, cache = ops.lstm_forward(acts[-1], host[l])
		acts.append(o)
		caches.append(cache)
	tol = REL_TC * (1 + T / 4.0)          # tf32 rounding compounds through the recurrence
	assert relerr(out.get(), acts[-1]) < tol

	ingrad, dhx, dcx = rnn.backwardData(G(bnd, dy), out, W, reserve, allocator=bnd.memoryPool)
	dw = rnn.backwardParams(G(bnd, x), out, reserve, allocator=bnd.memoryPool)
	dwparams = bnd.acquireRnnParams(rnn, dw)

	g = dy.astype(np.float64)
	for l in range(layers - 1, -1, -1):
		g, dp = ops.lstm_backward(acts[l], host[l], caches[l], g)
		for name, want in dp.items():
			assert relerr(dwparams[l][name].get(), want) < 2 * tol, name
	assert relerr(ingrad.get(), g) < 2 * tol


@pytest.mark.parametrize("mode", ["relu", "tanh"])
def test_plain_rnn_forward_backward(bnd, mode):
	# reference tests: Cuda/Wrappers/CuDnnRnn.py:20-176 (reluTest / tanhTest, uni-directional part)
	T, B, insz, H = 5, 6, 24, 40
	rng = np.random.RandomState(9)
	rnn, W, params = bnd.createRnn(insz, H, np.float32, mode=getattr(bnd.RNNMode, mode))
	W.set((rng.randn(*W.shape) * 0.2).astype(np.float32))
	host = _rnn_host_params(params[0])
	x, dy = rng.randn(T, B, insz).astype(np.float32), rng.randn(T, B, H).astype(np.float32)

	out, reserve = rnn.forward(G(bnd, x), W, allocator=bnd.memoryPool)
	want = ops.rnn_forward(x, host, mode)
	tol = REL_TC * (1 + T / 4.0)          # tf32 rounding compounds through the recurrence
	assert relerr(out.get(), want) < tol

	ingrad, _, _ = rnn.backwardData(G(bnd, dy), out, W, reserve, allocator=bnd.memoryPool)
	dw = rnn.backwardParams(G(bnd, x), out, reserve, allocator=bnd.memoryPool)
	dx, dp = ops.rnn_backward(x, host, want, dy, mode)
	assert relerr(ingrad.get(), dx) < 2 * tol
	dwparams = bnd.acquireRnnParams(rnn, dw)[0]
	for name, w in dp.items():
		assert relerr(dwparams[name].get(), w) < 2 * tol, name


def test_rnn_module_last_step_and_sequences(bnd):
	# Modules/RNN.py:122-161: getSequences=False returns the last step and scatters the gradient into a zero sequence
	import refshim
	M = refshim.modules()
	rng = np.random.RandomState(3)
	np.random.seed(3)
	T, B, insz, H = 6, 4, 16, 32
	mod = M.RNN(insz, H, mode="lstm", getSequences=False)
	x = rng.randn(T, B, insz).astype(np.float32)
	out = mod(G(bnd, x))
	assert out.shape == (B, H)
	host = _rnn_host_params(mod.params[0])
	want, cache = ops.lstm_forward(x, host)
	assert relerr(out.get(), want[-1]) < REL_TC

	g = rng.randn(B, H).astype(np.float32)
	mod.zeroGradParams() if hasattr(mod, "zeroGradParams") else None
	mod.backward(G(bnd, g))
	full = np.zeros((T, B, H))
	full[-1] = g
	dx, dp = ops.lstm_backward(x, host, cache, full)
	assert relerr(mod.grad.get(), dx) < 2 * REL_TC
	dwparams = bnd.acquireRnnParams(mod.descRnn, mod.vars["W"].grad)[0]
	assert relerr(dwparams["ri"].get(), dp["ri"]) < 2 * REL_TC



@pytest.mark.parametrize("case", [(3, 3, 4, 2), (6, 8, 40, 24), (10, 32, 64, 64)])
def test_gru_forward_backward(bnd, case):
	# reference test: Cuda/Wrappers/CuDnnRnn.py:303-419 (gruTest)
	T, B, insz, H = case
	rng = np.random.RandomState(T * 7 + H)
	rnn, W, params = bnd.createRnn(insz, H, np.float32, mode=bnd.RNNMode.gru)
	assert W.shape == (3 * H * (insz + H) + 6 * H, )
	W.set((rng.randn(*W.shape) * 0.3).astype(np.float32))
	host = _rnn_host_params(params[0])
	x, dy = rng.randn(T, B, insz).astype(np.float32), rng.randn(T, B, H).astype(np.float32)
	h0 = rng.randn(1, B, H).astype(np.float32)

	out, reserve = rnn.forward(G(bnd, x), W, hidden=G(bnd, h0), allocator=bnd.memoryPool)
	want, cache = ops.gru_forward(x, host, h0[0])
	tol = REL_TC * (1 + T / 4.0)
	assert relerr(out.get(), want) < tol
	ingrad, dhx, _ = rnn.backwardData(G(bnd, dy), out, W, reserve, allocator=bnd.memoryPool)
	dw = rnn.backwardParams(G(bnd, x), out, reserve, allocator=bnd.memoryPool)
	dx, dp = ops.gru_backward(x, host, cache, dy)
	assert relerr(ingrad.get(), dx) < 2 * tol
	dwparams = bnd.acquireRnnParams(rnn, dw)[0]
	for name, w in dp.items():
		assert relerr(dwparams[name].get(), w) < 2 * tol, name


@pytest.mark.parametrize("mode", ["lstm", "gru", "tanh"])
def test_bidirectional_rnn(bnd, mode):
	# reference test: Cuda/Wrappers/CuDnnRnn.py:95-176 (bidirectional tanhTest): two parameter sets, outputs concatenated
	T, B, insz, H = 5, 4, 12, 16
	rng = np.random.RandomState(31)
	rnn, W, params = bnd.createRnn(insz, H, np.float32, mode=getattr(bnd.RNNMode, mode), direction=bnd.DirectionMode.bi)
	assert len(params) == 2
	W.set((rng.randn(*W.shape) * 0.3).astype(np.float32))
	host = [_rnn_host_params(params[d]) for d in range(2)]
	x, dy = rng.randn(T, B, insz).astype(np.float32), rng.randn(T, B, 2 * H).astype(np.float32)

	out, reserve = rnn.forward(G(bnd, x), W, allocator=bnd.memoryPool)
	assert out.shape == (T, B, 2 * H)
	ingrad, _, _ = rnn.backwardData(G(bnd, dy), out, W, reserve, allocator=bnd.memoryPool)
	dw = rnn.backwardParams(G(bnd, x), out, reserve, allocator=bnd.memoryPool)
	dwparams = bnd.acquireRnnParams(rnn, dw)

	wantdx = np.zeros(x.shape)
	tol = REL_TC * (1 + T / 4.0)
	for d, reverse in ((0, False), (1, True)):
		dyd = dy[:, :, d * H:(d + 1) * H]
		if mode == "lstm":
			o, cache = ops.lstm_forward(x, host[d], reverse=reverse)
			dxd, dp = ops.lstm_backward(x, host[d], cache, dyd)
		elif mode == "gru":
			o, cache = ops.gru_forward(x, host[d], reverse=reverse)
			dxd, dp = ops.gru_backward(x, host[d], cache, dyd, reverse=reverse)
		else:
			xs, dys = (x[::-1], dyd[::-1]) if reverse else (x, dyd)
			o = ops.rnn_forward(xs, host[d], "tanh")
			dxd, dp = ops.rnn_backward(xs, host[d], o, dys, "tanh")
			if reverse:
				o, dxd = o[::-1], dxd[::-1]
		assert relerr(out.get()[:, :, d * H:(d + 1) * H], o) < tol
		wantdx += dxd
		for name, w in dp.items():
			assert relerr(dwparams[d][name].get(), w) < 2 * tol, (d, name)
	assert relerr(ingrad.get(), wantdx) < 2 * tol


def test_bidirectional_rnn_module_last_steps(bnd):
	# Modules/RNN.py:131-161: a bidirectional layer without sequences returns [forward last step, backward first step]
	import refshim
	M = refshim.modules()
	np.random.seed(9)
	rng = np.random.RandomState(9)
	T, B, insz, H = 4, 3, 8, 8
	mod = M.RNN(insz, H, mode="gru", direction="bi", getSequences=False)
	x = rng.randn(T, B, insz).astype(np.float32)
	fwd, bwd = mod(G(bnd, x))
	assert fwd.shape == (B, H) and bwd.shape == (B, H)
	host = [_rnn_host_params(mod.params[i]) for i in range(2)]
	assert relerr(fwd.get(), ops.gru_forward(x, host[0])[0][-1]) < 3e-3
	assert relerr(bwd.get(), ops.gru_forward(x, host[1], reverse=True)[0][0]) < 3e-3
	mod.backward([G(bnd, rng.randn(B, H).astype(np.float32)), G(bnd, rng.randn(B, H).astype(np.float32))])
	assert mod.grad.shape == (T, B, insz)


# ================================================================================================ 3-d convolution / pooling
CONV3D_CASES = [
	# N, C, D, H, W, K, (T, R, S), stride, pad, dilation, groups, bias
	(2, 4, 6, 9, 9, 8, (3, 3, 3), 1, 1, 1, 1, True),                  # reference test geometry (Cuda/Wrappers/CuDnn.py:106-144)
	(2, 32, 5, 12, 10, 48, (2, 3, 3), (1, 2, 1), (0, 1, 1), 1, 1, False),
	(1, 6, 7, 8, 8, 4, (3, 1, 2), (2, 1, 1), (1, 0, 0), (2, 1, 1), 2, True),     # depth stride + depth dilation + groups
]


@pytest.mark.parametrize("case", CONV3D_CASES)
def test_conv3d_fwd_bwd(bnd, case):
	N, C, D, H, W, K, fsize, stride, pad, dil, groups, bias = case
	rng = np.random.RandomState(abs(hash(case)) % (2 ** 31))
	x = rng.randn(N, C, D, H, W).astype(np.float32)
	w = rng.randn(K, C // groups, *fsize).astype(np.float32)
	b = rng.randn(K).astype(np.float32) if bias else None
	y = ops.conv3d(x, w, b, stride, pad, dil, groups)
	dy = rng.randn(*y.shape).astype(np.float32)

	gx, gw = G(bnd, x), G(bnd, w)
	out = bnd.dnn.convNd(gx, gw, G(bnd, b) if bias else None, stride, pad, dil, groups, allocator=bnd.memoryPool)
	assert out.shape == y.shape
	assert relerr(out.get(), y) < REL_TC

	dgrad = bnd.dnn.convNdBackwardData(G(bnd, dy), gw, None, gx, stride, pad, dil, None, groups, allocator=bnd.memoryPool)
	assert relerr(dgrad.get(), ops.conv3d_bwd_data(dy, w, x.shape, stride, pad, dil, groups)) < REL_TC

	w0 = rng.randn(*w.shape).astype(np.float32)
	wgrad, bgrad = G(bnd, w0), G(bnd, np.zeros(K, np.float32))
	bnd.dnn.convNdBackwardParams(gx, G(bnd, dy), gw, stride, pad, dil, groups, True, False, wgrad, bgrad, 0.5, 0.25,
								 allocator=bnd.memoryPool)
	dw, db = ops.conv3d_bwd_params(x, dy, w.shape, stride, pad, dil, groups)
	assert relerr(wgrad.get(), 0.5 * dw + 0.25 * w0) < REL_TC
	assert relerr(bgrad.get(), 0.5 * db) < 1e-5


@pytest.mark.parametrize("case", [((2, 3, 6, 8, 8), 2, 2, 0), ((2, 5, 7, 9, 10), (3, 2, 3), (2, 1, 2), (1, 0, 1)), ((1, 4, 5, 6, 6), 3, 1, 1)])
@pytest.mark.parametrize("mode", ["max", "avgWithPad", "avgNoPad"])
def test_pool3d(bnd, case, mode):
	# reference test: Cuda/Wrappers/CuDnn.py:414-451 (3-d max pooling with padding, forward + backward)
	shape, size, stride, pad = case
	rng = np.random.RandomState(sum(shape))
	x = rng.randn(*shape).astype(np.float32)
	code = {"max": bnd.PoolMode.max, "avgWithPad": bnd.PoolMode.avgWithPad, "avgNoPad": bnd.PoolMode.avgNoPad}[mode].value
	want, _ = ops.pool3d(x, size, stride, pad, mode)
	gx = G(bnd, x)
	out = bnd.dnn.poolNd(gx, size, stride, pad, code, allocator=bnd.memoryPool)
	assert out.shape == want.shape
	assert relerr(out.get(), want) < 1e-6
	dy = rng.randn(*want.shape).astype(np.float32)
	dx = bnd.dnn.poolNdBackward(G(bnd, dy), gx, out, size, stride, pad, code, allocator=bnd.memoryPool)
	assert relerr(dx.get(), ops.pool3d_bwd(x, dy, size, stride, pad, mode)) < 1e-6


def test_conv3d_pool3d_modules(bnd):
	import refshim
	M = refshim.modules()
	np.random.seed(4)
	rng = np.random.RandomState(4)
	net = M.Sequential()
	net.append(M.Conv3D(3, 8, 3, pad=1, initscheme="he")).append(M.MaxPool3D(2, 2)).append(M.AvgPool3D(2, 1, includePad=False))
	x = rng.randn(2, 3, 6, 8, 8).astype(np.float32)
	out = net(M.gpuarray.to_gpu(x))
	conv = net.graph[0]
	y = ops.conv3d(x, conv.W.get(), conv.b.get().ravel(), 1, 1, 1)
	p1, _ = ops.pool3d(y, 2, 2, 0, "max")
	p2, _ = ops.pool3d(p1, 2, 1, 0, "avgNoPad")
	assert out.shape == p2.shape and relerr(out.get(), p2) < REL_TC
	g = rng.randn(*p2.shape).astype(np.float32)
	net.backward(M.gpuarray.to_gpu(g))
	d1 = ops.pool3d_bwd(p1, g, 2, 1, 0, "avgNoPad")
	d0 = ops.pool3d_bwd(y, d1, 2, 2, 0, "max")
	assert relerr(net.grad.get(), ops.conv3d_bwd_data(d0, conv.W.get(), x.shape, 1, 1, 1)) < REL_TC


# ------------------------------------------------------------------------------------------ training closure
@pytest.mark.parametrize("shape", [(64, 10), (7, 1000), (5, 6, 3, 4), (1, 2)])
@pytest.mark.parametrize("weighted", [False, True])
def test_cross_entropy(bnd, shape, weighted):
	# reference test: Cuda/Kernels/Costs.py:253-301 (crossEntropyTest / wceTest)
	rng = np.random.RandomState(sum(shape))
	scores = rng.randn(*shape).astype(np.float32)
	labels = rng.randint(0, shape[1], (shape[0], ) + shape[2:]).astype(np.int32)
	weights = (rng.rand(shape[1]) + 0.5).astype(np.float32) if weighted else None
	err, grad = bnd.costmod.crossEntropy(G(bnd, scores), G(bnd, labels), None if weights is None else G(bnd, weights),
										 allocator=bnd.memoryPool)
	assert grad.shape == shape and err.shape == ()
	want_err, want_grad = ops.cross_entropy(scores, labels, weights)
	assert np.abs(grad.get() - want_grad).max() < 1e-6
	assert abs(float(err.get()) - want_err) < 1e-4 * max(1.0, abs(want_err))


def test_cross_entropy_cost_object_and_accuracy(bnd):
	# Cost/CrossEntropy.py:27-55: error per sample, accumulated mean error, validation = fraction of wrong arg-max labels
	from PuzzleLib.Cost.CrossEntropy import CrossEntropy
	rng = np.random.RandomState(12)
	cost = CrossEntropy(maxlabels=10)
	total = 0.0
	for step in range(2):
		scores = rng.randn(32, 10).astype(np.float32)
		labels = rng.randint(0, 10, (32, )).astype(np.int32)
		error, grad = cost(G(bnd, scores), G(bnd, labels))
		want_err, want_grad = ops.cross_entropy(scores, labels)
		total += want_err
		assert abs(error - want_err / 32) < 1e-5
		assert np.abs(grad.get() - want_grad).max() < 1e-6
	assert abs(cost.getMeanError() - total / 64) < 1e-5
	val = cost.validate(G(bnd, scores), G(bnd, labels))
	assert abs(val - float((scores.argmax(axis=1) != labels).mean())) < 1e-7
	maps = rng.randn(4, 5, 3, 3).astype(np.float32)
	maplabels = rng.randint(0, 5, (4, 3, 3)).astype(np.int32)
	assert abs(CrossEntropy().validate(G(bnd, maps), G(bnd, maplabels)) - float((maps.argmax(axis=1) != maplabels).mean())) < 1e-7


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_nesterov_and_adam_kernels(bnd, dtype):
	# reference tests: Optimizers/NesterovSGD.py:39-58, Optimizers/Adam.py:57-80 (calcTest)
	rng = np.random.RandomState(3)
	shape = (11, 13)
	atol = 1e-5 if dtype == np.float32 else 2e-2
	w, dw, mom = (rng.randn(*shape).astype(dtype) for _ in range(3))
	gw, gm = G(bnd, w), G(bnd, mom)
	bnd.nesterovMomSGDKer(np.dtype(dtype))(gw, G(bnd, dw), gm, 0.01, 0.9)
	pw, pm = ops.nesterov_update(w, dw, mom, 0.01, 0.9)
	assert np.abs(gw.get() - pw).max() < atol and np.abs(gm.get() - pm).max() < atol

	ms = (1.0 + rng.randn(*shape) ** 2).astype(np.float32)
	mg = rng.randn(*shape).astype(np.float32)
	gw, gmg, gms = G(bnd, w), G(bnd, mg), G(bnd, ms)
	bnd.adamKer(np.dtype(dtype))(gw, G(bnd, dw), gmg, gms, 0.0316, 0.1, 0.001, 1e-8)
	pw, pa, ps = ops.adam_update(w, dw, mg, ms, 0.0316, 0.1, 0.001, 1e-8)
	assert np.abs(gw.get() - pw).max() < atol
	assert np.abs(gmg.get() - pa).max() < 1e-5 and np.abs(gms.get() - ps).max() < 1e-5


def test_adam_and_nesterov_optimizers_train_a_linear_layer(bnd):
	# Optimizers/Optimizer.py trainSimpleTest: the cost of a small regression goes down under every optimizer
	import refshim
	M = refshim.modules()
	from PuzzleLib.Cost.CrossEntropy import CrossEntropy
	from PuzzleLib.Optimizers.Adam import Adam; from PuzzleLib.Optimizers.NesterovSGD import NesterovSGD
	for make in (lambda: Adam(alpha=1e-2), lambda: NesterovSGD(learnRate=1e-1, momRate=0.9)):
		np.random.seed(5)
		rng = np.random.RandomState(5)
		net = M.Sequential()
		net.append(M.Linear(20, 4))
		data = rng.randn(64, 20).astype(np.float32)
		labels = (data[:, :4].argmax(axis=1)).astype(np.int32)
		opt = make()
		opt.setupOn(net)
		cost = CrossEntropy()
		errors = []
		for step in range(40):
			opt.zeroGradParams()
			error, grad = cost(net(G(bnd, data)), G(bnd, labels))
			net.backward(grad)
			opt.update()
			errors.append(float(error))
		assert errors[-1] < 0.5 * errors[0], errors[::8]


# ------------------------------------------------------------------------------------------ memory reorganisation
@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_transpose_moveaxis_swapaxes(bnd, dtype):
	# reference tests: Cuda/Kernels/Memory.py:221-260 (every permutation / axis pair of these shapes)
	import itertools
	rng = np.random.RandomState(0)
	for shape in [(10, ), (10, 3), (10, 3, 5, 4, 2)]:
		host = rng.randn(*shape).astype(dtype)
		data = G(bnd, host)
		for axes in itertools.permutations(range(len(shape))):
			assert np.array_equal(bnd.dnn.transpose(data, axes=axes, allocator=bnd.memoryPool).get(), np.transpose(host, axes))
		for a, b in itertools.product(range(len(shape)), repeat=2):
			assert np.array_equal(bnd.dnn.moveaxis(data, a, b, allocator=bnd.memoryPool).get(), np.moveaxis(host, a, b))
			assert np.array_equal(bnd.dnn.swapaxes(data, a, b, allocator=bnd.memoryPool).get(), np.swapaxes(host, a, b))
	assert np.array_equal(bnd.dnn.transpose(G(bnd, host)).get(), host.T)
	with pytest.raises(ValueError):
		bnd.dnn.transpose(data, axes=(0, 0, 1, 2, 3))
	big = rng.randn(64, 49, 512).astype(np.float32)
	assert np.array_equal(bnd.dnn.transpose(G(bnd, big), axes=(0, 2, 1)).get(), big.transpose(0, 2, 1))


def test_depth_concat_and_split(bnd):
	# reference test: Cuda/Kernels/Memory.py:263-296 (maps centred in the largest plane)
	rng = np.random.RandomState(1)
	hosts = [rng.randn(3, 4, 3, 3).astype(np.float32), rng.randn(3, 2, 6, 6).astype(np.float32), rng.randn(3, 5, 4, 4).astype(np.float32)]
	arys = [G(bnd, h) for h in hosts]
	out = bnd.dnn.depthConcat(arys, allocator=bnd.memoryPool)
	want = np.zeros((3, 11, 6, 6), np.float32)
	want[:, :4, 1:4, 1:4], want[:, 4:6], want[:, 6:, 1:5, 1:5] = hosts
	assert np.array_equal(out.get(), want)
	hostgrad = rng.randn(*want.shape).astype(np.float32)
	grads = bnd.dnn.depthSplit(G(bnd, hostgrad), arys, allocator=bnd.memoryPool)
	for got, ref in zip(grads, [hostgrad[:, :4, 1:4, 1:4], hostgrad[:, 4:6], hostgrad[:, 6:, 1:5, 1:5]]):
		assert np.array_equal(got.get(), ref)


def test_memory_modules(bnd):
	# Modules/Transpose.py, MoveAxis.py, SwapAxes.py, DepthConcat.py: forward reorganises, backward undoes it
	import refshim
	M = refshim.modules()
	rng = np.random.RandomState(2)
	x = rng.randn(4, 3, 5, 2).astype(np.float32)
	for mod, fwd in ((M.Transpose(axes=(2, 0, 3, 1)), lambda a: a.transpose(2, 0, 3, 1)), (M.MoveAxis(1, 3), lambda a: np.moveaxis(a, 1, 3)),
					 (M.SwapAxes(0, 2), lambda a: np.swapaxes(a, 0, 2))):
		out = mod(G(bnd, x))
		assert np.array_equal(out.get(), fwd(x))
		mod.backward(out)
		assert np.array_equal(mod.grad.get(), x)
	cat = M.DepthConcat()
	a, b = rng.randn(2, 3, 4, 4).astype(np.float32), rng.randn(2, 1, 2, 2).astype(np.float32)
	out = cat([G(bnd, a), G(bnd, b)])
	assert out.shape == (2, 4, 4, 4) and np.array_equal(out.get()[:, 3:, 1:3, 1:3], b)
	cat.backward(out)
	assert np.array_equal(cat.grad[0].get(), a) and np.array_equal(cat.grad[1].get(), b)


def test_rng_fills_are_reproducible_and_well_distributed(bnd):
	# reference: Cuda/Source/Libs/CuRand.c fillInteger / fillUniform / fillNormal (moments and range checks; the stream itself is
	# Philox here, XORWOW there)
	from puzzlelib_b200.backend import RandomNumberGenerator
	n = 1 << 20
	a, b = RandomNumberGenerator(seed=77), RandomNumberGenerator(seed=77)
	x, y = bnd.GPUArray.empty((n, ), np.uint32), bnd.GPUArray.empty((n + 3, ), np.uint32)
	a.fillInteger(x)
	b.fillInteger(y)
	hx = x.get()
	assert np.array_equal(hx, y.get()[:n])                       # a function of (seed, offset, index) only
	a.fillInteger(x)
	assert not np.array_equal(hx, x.get())                        # the offset advances
	bits = np.unpackbits(hx.view(np.uint8))
	assert abs(bits.mean() - 0.5) < 2e-3 and len(np.unique(hx)) > 0.999 * n
	u = bnd.GPUArray.empty((n, ), np.float32)
	a.fillUniform(u, -2.0, 3.0)
	hu = u.get()
	assert hu.min() > -2.0 and hu.max() <= 3.0 and abs(hu.mean() - 0.5) < 1e-2 and abs(hu.var() - 25.0 / 12.0) < 2e-2
	a.fillNormal(u, 1.0, 2.0)
	hn = u.get()
	assert abs(hn.mean() - 1.0) < 1e-2 and abs(hn.std() - 2.0) < 1e-2 and abs(((hn - 1.0) ** 3).mean()) < 0.1
	assert abs((np.abs(hn - 1.0) < 2.0).mean() - 0.6827) < 5e-3
	with pytest.raises(ValueError):
		a.fillUniform(x)


@pytest.mark.parametrize("dtype", [np.float32, np.float16])
def test_dropout_kernel_and_module(bnd, dtype):
	# reference: Cuda/Kernels/ElementWise.py:495-580, Modules/Dropout.py:36-75
	import refshim
	M = refshim.modules()
	rng = np.random.RandomState(8)
	shape = (8, 6, 5, 7)
	x = rng.randn(*shape).astype(dtype)
	parttype = np.uint32 if dtype == np.float32 else np.uint16
	words = rng.randint(0, np.iinfo(parttype).max, size=x.size, dtype=np.int64).astype(parttype)
	v, p = int(0.7 * np.iinfo(parttype).max), 0.7
	out = bnd.GPUArray.empty(shape, dtype)
	bnd.dropoutKer(np.dtype(dtype))(out, G(bnd, x), G(bnd, words), v, np.float32(p))
	want = (x.astype(np.float32) * (words.reshape(shape) < v) / np.float32(p)).astype(dtype)
	assert np.array_equal(out.get(), want)
	mapwords = words[:shape[0] * shape[1]]
	bnd.dropout2dKer(np.dtype(dtype))(out, G(bnd, x), G(bnd, mapwords), v, np.float32(p), shape[2] * shape[3])
	want2 = (x.astype(np.float32) * (mapwords.reshape(shape[0], shape[1], 1, 1) < v) / np.float32(p)).astype(dtype)
	assert np.array_equal(out.get(), want2)

	mod = M.Dropout(p=0.3)
	mod.calcMode(dtype)
	big = np.ones((64, 1024), dtype)
	y = mod(G(bnd, big)).get().astype(np.float32)
	kept = y != 0
	assert abs(kept.mean() - 0.7) < 0.02 and np.allclose(y[kept], 1.0 / 0.7, rtol=2e-3)
	mod.backward(G(bnd, big))
	assert np.array_equal(mod.grad.get().astype(np.float32) != 0, kept)      # the same mask on the way back
	mod.evalMode()
	assert np.array_equal(mod(G(bnd, big)).get(), big)


@pytest.mark.parametrize("dtype,atol", [(np.float32, 1e-5), (np.float16, 2e-2)])
def test_lrn_across_and_within_maps(bnd, dtype, atol):
	# reference tests: Cuda/Wrappers/CuDnnNorm.py:185-268 (mapLRN2dTest, crossMapLRN2dTest) + larger / even-window cases
	rng = np.random.RandomState(14)
	for shape, N in [((2, 2, 9, 10), 5), ((2, 10, 2, 3), 5), ((3, 17, 6, 7), 4), ((1, 3, 5, 5), 7)]:
		x, g = rng.randn(*shape).astype(dtype), rng.randn(*shape).astype(dtype)
		alpha, beta, K = 1.0, 0.5, 2.0
		y = bnd.dnn.crossMapLRN(G(bnd, x), N=N, alpha=alpha, beta=beta, K=K, allocator=bnd.memoryPool)
		dx = bnd.dnn.crossMapLRNBackward(G(bnd, x), y, G(bnd, g), N=N, alpha=alpha, beta=beta, K=K, allocator=bnd.memoryPool)
		wy, wdx = ops.lrn(x, N, alpha, beta, K, True, grad=g)
		assert np.abs(y.get() - wy).max() < atol and np.abs(dx.get() - wdx).max() < 4 * atol
		y = bnd.dnn.mapLRN(G(bnd, x), N=N, alpha=alpha, beta=beta, K=K, allocator=bnd.memoryPool)
		dx = bnd.dnn.mapLRNBackward(G(bnd, x), G(bnd, g), N=N, alpha=alpha, beta=beta, K=K, allocator=bnd.memoryPool)
		wy, wdx = ops.lrn(x, N, alpha, beta, K, False, grad=g)
		assert np.abs(y.get() - wy).max() < atol and np.abs(dx.get() - wdx).max() < 4 * atol


@pytest.mark.parametrize("dtype,rtol", [(np.float32, 1e-5), (np.float16, 2e-3)])
def test_blas_dot_and_norms(bnd, dtype, rtol):
	# reference tests: Cuda/Wrappers/CuBlas.py (dot / l1norm / l2norm against numpy)
	rng = np.random.RandomState(21)
	for n in (1, 1000, 1 << 20):
		x, y = rng.randn(n).astype(dtype), rng.randn(n).astype(dtype)
		x64, y64 = x.astype(np.float64), y.astype(np.float64)
		scale = np.abs(x64 * y64).sum() + 1e-30
		assert abs(bnd.blas.dot(G(bnd, x), G(bnd, y)) - float(x64 @ y64)) < rtol * scale
		assert abs(bnd.blas.l1norm(G(bnd, x)) - np.abs(x64).sum()) < rtol * np.abs(x64).sum()
		assert abs(bnd.blas.l2norm(G(bnd, x)) - np.sqrt((x64 * x64).sum())) < rtol * np.sqrt((x64 * x64).sum())
	with pytest.raises(ValueError):
		bnd.blas.dot(G(bnd, x), G(bnd, y[:5]))


def test_training_kernels_match_the_reference_cpu_backend_golden_vectors(bnd):
	# tests/golden/ref_cpu_train.npz holds outputs of the reference's own gcc-JIT CPU kernels / optimizer objects
	import os
	import refshim
	M = refshim.modules()
	from PuzzleLib.Optimizers.Adam import Adam; from PuzzleLib.Optimizers.NesterovSGD import NesterovSGD; from PuzzleLib.Optimizers.MomentumSGD import MomentumSGD
	g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cpu_train.npz"))
	f32 = np.dtype(np.float32)
	w, dw, mom, mg, ms = (g[k] for k in ("w", "dw", "mom", "mg", "ms"))
	for name, ker in (("classic", bnd.classicMomSGDKer), ("nesterov", bnd.nesterovMomSGDKer)):
		gw, gm = G(bnd, w), G(bnd, mom)
		ker(f32)(gw, G(bnd, dw), gm, 0.01, 0.9)
		assert np.allclose(gw.get(), g["%s_w" % name], rtol=0, atol=1e-6) and np.allclose(gm.get(), g["%s_mom" % name], rtol=0, atol=1e-6)
	gw, gmg, gms = G(bnd, w), G(bnd, mg), G(bnd, ms)
	lr, fix1, fix2, eps = (float(v) for v in g["adam_args"])
	bnd.adamKer(f32)(gw, G(bnd, dw), gmg, gms, lr, fix1, fix2, eps)
	assert np.allclose(gw.get(), g["adam_w"], rtol=0, atol=1e-6) and np.allclose(gmg.get(), g["adam_mg"], rtol=0, atol=1e-6)
	assert np.allclose(gms.get(), g["adam_ms"], rtol=0, atol=1e-6)

	x = g["drop_x"]
	out = bnd.GPUArray.empty(x.shape, np.float32)
	bnd.dropoutKer(f32)(out, G(bnd, x), G(bnd, g["drop_words"]), int(g["drop_v"][0]), g["drop_p"][0])
	assert np.allclose(out.get(), g["drop_y"], rtol=3e-7, atol=0) and np.array_equal(out.get() == 0, g["drop_y"] == 0)
	bnd.dropout2dKer(f32)(out, G(bnd, x), G(bnd, g["drop2d_words"]), int(g["drop_v"][0]), g["drop_p"][0], x.shape[2] * x.shape[3])
	assert np.allclose(out.get(), g["drop2d_y"], rtol=3e-7, atol=0) and np.array_equal(out.get() == 0, g["drop2d_y"] == 0)

	class Holder:
		def __init__(self, var):
			self.var = var

		def getVarTable(self):
			return {self.var: ["w"]}

		def getVar(self, name):
			return self.var

	for name, make in (("adam", lambda: Adam(alpha=1e-2)), ("nesterov", lambda: NesterovSGD(learnRate=1e-1, momRate=0.9)),
					   ("momentum", lambda: MomentumSGD(learnRate=1e-1, momRate=0.9))):
		var = M.Variable(G(bnd, w))
		opt = make()
		opt.setupOn(Holder(var))
		for gr in g["opt_grads"]:
			var.grad.set(gr)
			opt.update()
		assert np.allclose(var.data.get(), g["opt_%s_w" % name], rtol=0, atol=2e-5), name


@pytest.mark.parametrize("dtype,tol", [(np.float32, 1e-5), (np.float16, 2e-2)])
def test_matvec_grouped(bnd, dtype, tol):
	# reference: Cuda/Kernels/MatVec.py:311-343 (matmod.matvec behind Blas.mulTensorOnVecGroup); exact fp32 accumulation
	rng = np.random.RandomState(17)
	for shape in [(5, 7), (3, 33, 65), (2, 4, 1, 9), (6, 128, 40)]:
		mat = rng.randn(*shape).astype(dtype)
		m64 = mat.astype(np.float64)
		for axis in (0, 1):
			vec = rng.randn(*(shape[:-2] + ((shape[-1], ) if axis == 1 else (shape[-2], )))).astype(dtype)
			v64 = vec.astype(np.float64)
			want = np.einsum("...hw,...w->...h", m64, v64) if axis == 1 else np.einsum("...hw,...h->...w", m64, v64)
			got = bnd.matmod.matvec(G(bnd, mat), G(bnd, vec), axis=axis, allocator=bnd.memoryPool)
			assert got.shape == want.shape and np.abs(got.get() - want).max() < tol * max(1.0, np.abs(want).max())
			base = rng.randn(*want.shape).astype(dtype)
			out = G(bnd, base)
			bnd.matmod.matvec(G(bnd, mat), G(bnd, vec), axis=axis, out=out, alpha=0.5, beta=2.0)
			ref = 2.0 * base.astype(np.float64) + 0.5 * want
			assert np.abs(out.get() - ref).max() < 2 * tol * max(1.0, np.abs(ref).max())
	with pytest.raises(ValueError):
		bnd.matmod.matvec(G(bnd, mat), G(bnd, vec[:, :3]), axis=0)


# ================================================================================================ deferred zero fill
@pytest.mark.parametrize("dtname", ["float32", "float16"])
def test_deferred_fill_fuses_with_accumulation_and_keeps_the_bits(bnd, dtname):
	"""`y.fill(0); y += a1*x1; y += a2*x2` (Modules/Add.py:15-23, Replicate.py:18-29) is issued as ONE launch; the result must
	be bit-identical to the three launches, and nothing may observe y (or move x) before the pending launch is flushed."""
	from puzzlelib_b200 import driver
	dt = np.dtype(dtname)
	rng = np.random.RandomState(5)
	n = (1 << 19) + 7                                   # above the deferral threshold for both types, odd tail
	x1, x2, x3 = (rng.randn(n).astype(dt) for _ in range(3))
	g1, g2, g3 = G(bnd, x1), G(bnd, x2), G(bnd, x3)
	axpy = bnd.toVectorAddVectorKer(dt)

	def reference(arrays, alphas):
		y = bnd.GPUArray.empty((n, ), dt)
		check = driver.lib.pz_memset8(y.ptr, 0, y.nbytes, None)
		assert check == 0
		for a, alpha in zip(arrays, alphas):
			assert driver.lib.pz_axpy(driver.dtypeCode(dt), y.ptr, a.ptr, alpha, n, None) == 0
		return y.get()

	for arrays, alphas in (((g1, g2), (1.0, 1.0)), ((g1, g2), (0.5, -1.25)), ((g1, g2, g3), (1.0, 1.0, 2.0)), ((g1, ), (3.0, ))):
		want = reference(arrays, alphas)
		launches = driver.launchCount()
		y = bnd.GPUArray.empty((n, ), dt)
		y.fill(0)
		assert driver.deferred is not None                              # held back
		for a, alpha in zip(arrays, alphas):
			axpy(y.ravel(), a.ravel(), alpha)                            # views of the same memory, like the reference modules
		got = y.get()
		assert driver.deferred is None
		assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), (alphas, dtname)
		assert driver.launchCount() - launches == max(1, len(arrays) - 1)

	# hazards: the source changes after the accumulation was absorbed -> the pending launch must have run first
	y = bnd.GPUArray.empty((n, ), dt)
	y.fill(0)
	src = G(bnd, x1)
	axpy(y, src, 1.0)
	src.fill(7)
	assert np.array_equal(y.get(), x1)
	# a pending fill is visible to every other consumer
	y.fill(0)
	assert not y.get().any()
	z = bnd.GPUArray.empty((n, ), dt)
	z.fill(0)
	w = z + g1                                                           # not the fused kernel: plain read of a pending array
	assert np.array_equal(w.get(), x1)
	# a small or non-zero fill is not deferred
	small = bnd.GPUArray.empty((16, ), dt)
	small.fill(0)
	assert driver.deferred is None
	y.fill(1)
	assert driver.deferred is None and np.array_equal(y.get(), np.ones(n, dt))


def test_lstm_dropout_between_layers(bnd):
	"""cudnnSetRNNDescriptor's dropout: in training the output of every layer but the last passes through a dropout mask (keep rule
	and scaling of Modules/Dropout.py:33-77); inference ignores it.  Checked against the oracle with the masks the forward pass stored."""
	T, B, insz, H, p = 6, 8, 24, 32, 0.4
	rng = np.random.RandomState(77)
	rnn, W, params = bnd.createRnn(insz, H, np.float32, layers=2, mode=bnd.RNNMode.lstm, dropout=p, seed=1234)
	W.set((rng.randn(*W.shape) * 0.2).astype(np.float32))
	host = [_rnn_host_params(q) for q in params]
	x = rng.randn(T, B, insz).astype(np.float32)
	dy = rng.randn(T, B, H).astype(np.float32)

	# inference: no dropout
	plain = rnn.forward(G(bnd, x), W, test=True, allocator=bnd.memoryPool)
	l1, _ = ops.lstm_forward(x.astype(np.float64), host[0])
	l2, _ = ops.lstm_forward(l1, host[1])
	tol = REL_TC * (1 + T / 4.0)
	assert relerr(plain.get(), l2) < tol

	out, reserve = rnn.forward(G(bnd, x), W, allocator=bnd.memoryPool)
	assert len(reserve.rands) == 1 and reserve.rands[0].shape == (T, B, H)
	keep = reserve.rands[0].get().astype(np.uint64) < int((1.0 - p) * np.iinfo(np.uint32).max)
	assert 0.45 < keep.mean() < 0.75                                   # about 1 - p of the activations survive
	scale = keep / (1.0 - p)

	o1, c1 = ops.lstm_forward(x.astype(np.float64), host[0])
	o2, c2 = ops.lstm_forward(o1 * scale, host[1])
	assert relerr(out.get(), o2) < tol
	assert relerr(out.get(), l2) > 10 * tol                            # and it is not the inference result

	ingrad, _, _ = rnn.backwardData(G(bnd, dy), out, W, reserve, allocator=bnd.memoryPool)
	dw = rnn.backwardParams(G(bnd, x), out, reserve, allocator=bnd.memoryPool)
	dwparams = bnd.acquireRnnParams(rnn, dw)
	g2, dp2 = ops.lstm_backward(o1 * scale, host[1], c2, dy.astype(np.float64))
	g1, dp1 = ops.lstm_backward(x.astype(np.float64), host[0], c1, g2 * scale)
	assert relerr(ingrad.get(), g1) < 2 * tol
	for layer, dp in ((0, dp1), (1, dp2)):
		for name, want in dp.items():
			assert relerr(dwparams[layer][name].get(), want) < 2 * tol, (layer, name)

	# a second descriptor with the same seed draws the same masks
	rnn2, W2, _ = bnd.createRnn(insz, H, np.float32, layers=2, mode=bnd.RNNMode.lstm, dropout=p, seed=1234)
	W2.set(W.get())
	out2, reserve2 = rnn2.forward(G(bnd, x), W2, allocator=bnd.memoryPool)
	assert np.array_equal(reserve2.rands[0].get(), reserve.rands[0].get()) and np.array_equal(out2.get(), out.get())


@pytest.mark.parametrize("dtname", ["float32", "float16"])
def test_relu_of_a_pending_sum_is_one_launch_with_the_same_bits(bnd, dtname):
	"""Add -> Activation(relu) of every ResNet block (Modules/Add.py:15-23, Activation.py:69-71): `y.fill(0); y += x1; y += x2;
	relu(out, y)` is issued as ONE launch that stores both y and out; bits identical to the four launches"""
	from puzzlelib_b200 import driver
	dt = np.dtype(dtname)
	rng = np.random.RandomState(8)
	n = (1 << 19) + 5
	x1, x2 = (rng.randn(n).astype(dt) for _ in range(2))
	g1, g2 = G(bnd, x1), G(bnd, x2)
	axpy, relu = bnd.toVectorAddVectorKer(dt), bnd.reluKer(dt)

	# the four separate launches, issued directly through the C-ABI (no deferral)
	code = driver.dtypeCode(dt)
	yr, outr = bnd.GPUArray.empty((n, ), dt), bnd.GPUArray.empty((n, ), dt)
	assert driver.lib.pz_memset8(yr.ptr, 0, yr.nbytes, None) == 0
	assert driver.lib.pz_axpy(code, yr.ptr, g1.ptr, 0.5, n, None) == 0
	assert driver.lib.pz_axpy(code, yr.ptr, g2.ptr, 1.5, n, None) == 0
	relu(outr, yr)
	want_y, want_out = yr.get(), outr.get()
	assert (want_out >= 0).all() and (want_out > 0).any() and (want_out == 0).any()

	launches = driver.launchCount()
	y, out = bnd.GPUArray.empty((n, ), dt), bnd.GPUArray.empty((n, ), dt)
	y.fill(0)
	axpy(y, g1, 0.5)
	axpy(y, g2, 1.5)
	assert driver.deferred is not None and driver.launchCount() == launches      # the complete sum is still pending
	relu(out, y)
	assert driver.deferred is None and driver.launchCount() - launches == 1
	assert np.array_equal(y.get().view(np.uint8), want_y.view(np.uint8))
	assert np.array_equal(out.get().view(np.uint8), want_out.view(np.uint8))

	# in place, or on another tensor: the sum is flushed first and the plain kernels run
	y.fill(0)
	axpy(y, g1, 1.0)
	axpy(y, g2, 1.0)
	relu(y, y)
	s = (x1.astype(np.float32) + x2.astype(np.float32)).astype(dt)
	assert np.array_equal(y.get(), np.where(s > 0, s, dt.type(0)))
	y.fill(0)
	axpy(y, g1, 1.0)
	axpy(y, g2, 1.0)
	relu(out, g1)
	assert np.array_equal(out.get(), np.where(x1 > 0, x1, dt.type(0))) and np.array_equal(y.get(), s)


@pytest.mark.parametrize("dtname", ["float32", "float16"])
def test_relu_derivative_of_a_pending_sum_is_one_launch_with_the_same_bits(bnd, dtname):
	"""Replicate.updateGrad (sum of two gradients, Modules/Replicate.py:24-29) followed by the previous block's Activation.updateGrad
	(Modules/Activation.py:74-76): one launch stores the sum and sum * (ref > 0)"""
	from puzzlelib_b200 import driver
	dt = np.dtype(dtname)
	rng = np.random.RandomState(9)
	n = (1 << 19) + 3
	x1, x2, r = (rng.randn(n).astype(dt) for _ in range(3))
	r = np.where(r > 0, r, dt.type(0))                           # a ReLU output
	g1, g2, gr = G(bnd, x1), G(bnd, x2), G(bnd, r)
	axpy, reluDer = bnd.toVectorAddVectorKer(dt), bnd.reluDerKer(dt)

	code = driver.dtypeCode(dt)
	yr, inr = bnd.GPUArray.empty((n, ), dt), bnd.GPUArray.empty((n, ), dt)
	assert driver.lib.pz_memset8(yr.ptr, 0, yr.nbytes, None) == 0
	assert driver.lib.pz_axpy(code, yr.ptr, g1.ptr, 1.0, n, None) == 0
	assert driver.lib.pz_axpy(code, yr.ptr, g2.ptr, 1.0, n, None) == 0
	reluDer(inr, yr, gr)
	want_y, want_in = yr.get(), inr.get()

	launches = driver.launchCount()
	y, ingrad = bnd.GPUArray.empty((n, ), dt), bnd.GPUArray.empty((n, ), dt)
	y.fill(0)
	axpy(y, g1, 1.0)
	axpy(y, g2, 1.0)
	reluDer(ingrad, y, gr)
	assert driver.deferred is None and driver.launchCount() - launches == 1
	assert np.array_equal(y.get().view(np.uint8), want_y.view(np.uint8))
	assert np.array_equal(ingrad.get().view(np.uint8), want_in.view(np.uint8))

	# in place (inplace=True activations hand the same array as ingrad and outgrad): flushed, then the plain kernel
	y.fill(0)
	axpy(y, g1, 1.0)
	axpy(y, g2, 1.0)
	reluDer(y, y, gr)
	assert np.array_equal(y.get().view(np.uint8), want_in.view(np.uint8))


@pytest.mark.parametrize("shape", [(16, 32, 55, 55), (8, 16, 112, 112), (64, 48, 14, 14)])
def test_relu_after_batchnorm_is_folded_into_the_pending_launch(bnd, shape):
	"""conv -> bn -> relu of every ResNet block: the train-mode batch-norm launch is held back for one call and a ReLU of its output
	rides on the same pass (cluster kernel) or follows it inside the same C call (other paths).  Every tensor the reference
	materialises -- y, relu(y), saved and running statistics -- is bit-identical to the two separate launches."""
	from puzzlelib_b200 import driver
	N, C, H, W = shape
	rng = np.random.RandomState(C)
	x = (rng.randn(*shape) * 2 + 0.5).astype(np.float32)
	scale, bias = rng.randn(1, C, 1, 1).astype(np.float32), rng.randn(1, C, 1, 1).astype(np.float32)
	relu = bnd.reluKer(np.float32)

	def run(fused):
		gx = G(bnd, x)
		mean, var = bnd.GPUArray.zeros((1, C, 1, 1), np.float32), G(bnd, np.ones((1, C, 1, 1), np.float32))
		gs, gb = G(bnd, scale), G(bnd, bias)
		driver.flushDeferred()
		launches = driver.launchCount()
		y, sm, siv = bnd.dnn.batchNormNd(gx, mean, var, gs, gb, 1e-5, 0.25, False)
		if not fused:
			driver.flushDeferred()                          # the plain launch, then the plain ReLU kernel
		else:
			assert driver.deferred is not None and driver.launchCount() == launches
		z = bnd.GPUArray.empty(shape, np.float32)
		relu(z, y)
		assert driver.deferred is None
		return [a.get() for a in (y, z, sm, siv, mean, var)], driver.launchCount() - launches

	plain, nplain = run(False)
	fused, nfused = run(True)
	for name, a, b in zip(("y", "relu", "savemean", "saveinvvar", "mean", "var"), plain, fused):
		assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), name
	assert (fused[1] >= 0).all() and (fused[1] == np.where(fused[0] > 0, fused[0], 0)).all()
	assert nfused <= nplain

	# anything else that touches the output first gets the plain launch
	gx = G(bnd, x)
	mean, var = bnd.GPUArray.zeros((1, C, 1, 1), np.float32), G(bnd, np.ones((1, C, 1, 1), np.float32))
	y, sm, siv = bnd.dnn.batchNormNd(gx, mean, var, G(bnd, scale), G(bnd, bias), 1e-5, 0.25, False)
	assert np.array_equal(y.get().view(np.uint8), plain[0].view(np.uint8)) and driver.deferred is None


def test_batchnorm_in_place_on_the_two_kernel_path(bnd):
	"""planes too large for the shared-memory stash take the statistics + apply kernels; with out == data the apply kernel must not
	rebuild the mean from an element another CTA has already normalised (the pivot travels through a side buffer)"""
	shape = (64, 4, 112, 112)
	rng = np.random.RandomState(12)
	x = (rng.randn(*shape) * 1.5 + 2.0).astype(np.float32)
	scale, bias = rng.randn(1, 4, 1, 1).astype(np.float32), rng.randn(1, 4, 1, 1).astype(np.float32)

	def run(inplace):
		gx = G(bnd, x)
		mean, var = bnd.GPUArray.zeros((1, 4, 1, 1), np.float32), G(bnd, np.ones((1, 4, 1, 1), np.float32))
		y, sm, siv = bnd.dnn.batchNormNd(gx, mean, var, G(bnd, scale), G(bnd, bias), 1e-5, 1.0, False, out=gx if inplace else None)
		return y.get(), sm.get(), siv.get(), mean.get()

	# (this path accumulates with atomics: runs agree to rounding, not bit for bit)
	want, mu, inv, _, _ = ops.batchnorm_train(x, scale.ravel(), bias.ravel(), np.zeros(4), np.ones(4), 1e-5, 1.0)
	for got in (run(False), run(True)):
		assert relerr(got[0], want) < 2e-5
		assert np.allclose(got[1].ravel(), mu, atol=1e-5) and np.allclose(got[2].ravel(), inv, rtol=1e-5)
		assert np.allclose(got[3].ravel(), mu, atol=1e-5)                # factor 1: the running mean is the batch mean


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(64, 64, 28, 28), (32, 16, 55, 55), (8, 24, 112, 112), (64, 256, 7, 7)])
@pytest.mark.parametrize("terms", ["both", "scale", "bias-first"])
def test_batchnorm_backward_with_absorbed_parameter_gradient_accumulation(bnd, shape, terms):
	"""BatchNormND.accGradParams (Modules/BatchNormND.py:74-83) follows the backward pass with two addVectorToVector launches over C
	floats; the backend holds the pass back and lets them ride on it (pz_bn_bwd_acc).  Same bits as the separate launches -- on the
	cluster kernels and on the two-kernel path (the 112 x 112 planes), whatever the order and number of accumulations."""
	from puzzlelib_b200 import driver
	rng = np.random.RandomState(sum(shape) + len(terms))
	C = shape[1]
	x, dy = rng.randn(*shape).astype(np.float32), rng.randn(*shape).astype(np.float32)
	scale = rng.randn(C).astype(np.float32)
	mu, inv = x.mean(axis=(0, 2, 3)).astype(np.float32), (1.0 / np.sqrt(x.var(axis=(0, 2, 3)) + 1e-5)).astype(np.float32)
	acc0 = rng.randn(2, C).astype(np.float32)
	alpha, beta = 0.75, 0.9
	add = bnd.addKer(np.float32)

	def run(fused):
		gx, gdy, gscale, gmu, ginv = (G(bnd, a) for a in (x, dy, scale, mu, inv))
		sacc, bacc = G(bnd, acc0[0]), G(bnd, acc0[1])
		before = driver.launchCount()
		dx, sg, bg = bnd.dnn.batchNormNdBackward(gdy, gx, gscale, gmu, ginv, 1e-5, allocator=bnd.memoryPool)
		if not fused:
			driver.flushDeferred()
		if terms == "both":
			add(sacc, sg, alpha, sacc, beta)
			add(bacc, bg, alpha, bacc, beta)
		elif terms == "scale":
			add(sacc, sg, alpha, sacc, beta)
		else:
			add(bacc, bg, 1.0, bacc, 1.0)
			add(sacc, sg, alpha, sacc, 0.0)
		driver.flushDeferred()
		launches = driver.launchCount() - before
		return [a.get() for a in (dx, sg, bg, sacc, bacc)], launches

	fused, nf = run(True)
	plain, npl = run(False)
	for a, b in zip(fused, plain):
		assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
	assert nf == npl - (1 if terms == "scale" else 2)            # the accumulations launched nothing of their own
	want = alpha * plain[1] + (beta if terms != "bias-first" else 0.0) * acc0[0]
	assert np.allclose(fused[3], want, rtol=1e-5, atol=1e-5)
