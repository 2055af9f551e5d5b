"""GPU tier: the module mirror and whole networks.

Whole-net parity is TEACHER-FORCED per layer: after one forward+backward pass of the GPU net, every leaf module's
output / input-gradient / parameter-gradient is recomputed by the CPU oracle from that module's own GPU inputs, so each
operator is held to the per-op bar (1e-3 relative for tensor-core contractions, 1e-5 for the rest) on the real
activations of the real net instead of on an error that compounds over 50 layers.
"""
import numpy as np
import pytest

from oracle import ops, refnet

pytestmark = pytest.mark.gpu


def relerr(got, want):
	want = np.asarray(want, dtype=np.float64)
	return float(np.abs(np.asarray(got, dtype=np.float64) - want).max() / (np.abs(want).max() + 1e-30))


@pytest.fixture()
def M(bnd):
	import refshim
	return refshim.modules()


def record_backward(modules):
	"""Patch Module.backward so that every leaf remembers the gradient it was handed."""
	orig = modules.Module.backward

	def backward(self, grad, *args, **kwargs):
		self._gradIn = grad
		self._bwdKwargs = kwargs
		return orig(self, grad, *args, **kwargs)

	modules.Module.backward = backward
	return orig


def check_leaf(M, mod, failures, tol_tc=1e-3, tol=1e-5, floor=0.0):
	"""`floor`: the rounding of 16-bit storage (every bar is at least this)"""
	name = "%s(%s)" % (type(mod).__name__, mod.name)
	g = getattr(mod, "_gradIn", None)
	tol_tc, tol = max(tol_tc, floor), max(tol, floor)

	def expect(tag, got, want, bar):
		err = relerr(got, want)
		bar = max(bar, floor)
		if not err < bar:
			failures.append("%s %s relerr %.3e > %.1e" % (name, tag, err, bar))

	if isinstance(mod, M.Conv2D):
		x, w = mod.inData.get(), mod.W.get()
		b = mod.b.get().ravel() if mod.b is not None else None
		expect("fwd", mod.data.get(), ops.conv2d(x, w, b, mod.stride, mod.pad, mod.dilation, mod.groups), tol_tc)
		if g is not None:
			dy = g.get()
			expect("dgrad", mod.grad.get(), ops.conv2d_bwd_data(dy, w, x.shape, None, mod.stride, mod.pad, mod.dilation, 0, mod.groups), tol_tc)
			want = ops.conv2d_bwd_params(x, dy, w.shape, mod.stride, mod.pad, mod.dilation, mod.groups, withbias=b is not None)
			expect("wgrad", mod.vars["W"].grad.get(), want[0] if b is not None else want, tol_tc)
			if b is not None:
				expect("bgrad", mod.vars["b"].grad.get().ravel(), want[1], tol)
	elif isinstance(mod, M.BatchNorm2D):
		x = mod.inData.get()
		y, mu, inv, _, _ = ops.batchnorm_train(x, mod.scale.get().ravel(), mod.bias.get().ravel(), np.zeros(mod.maps), np.ones(mod.maps),
											   mod.epsilon, 1.0)
		expect("fwd", mod.data.get(), y, 2e-5)
		# the mean is accurate relative to the channel's spread (fp32 accumulation), not relative to max |mean| (which is ~0)
		if not float(np.abs((mod.savemean.get().ravel() - mu) * inv).max()) < max(2e-5, floor):
			failures.append("%s savemean off by more than 2e-5 std" % name)
		expect("saveinvvar", mod.saveinvvar.get().ravel(), inv, tol)
		if g is not None:
			dx, dscale, dbias = ops.batchnorm_bwd(x, g.get(), mod.scale.get().ravel(), mu, inv)
			expect("dx", mod.grad.get(), dx, 2e-5)
			expect("dscale", mod.vars["scale"].grad.get().ravel(), dscale, 2e-5)
			expect("dbias", mod.vars["bias"].grad.get().ravel(), dbias, 2e-5)
	elif isinstance(mod, M.Activation):
		kind = mod.activation.value
		expect("fwd", mod.data.get(), ops.activation(kind, mod.inData.get(), *mod.actArgs), tol)
		if g is not None:
			expect("bwd", mod.grad.get(), ops.activation_bwd(kind, g.get(), mod.data.get(), *mod.actArgs), tol)
	elif isinstance(mod, M.MaxPool2D) and not mod.useMask:
		x = mod.inData.get()
		expect("fwd", mod.data.get(), ops.pool2d(x, mod.size, mod.stride, mod.pad, "max"), 1e-7)
		if g is not None:
			expect("bwd", mod.grad.get(), ops.pool2d_bwd(x, mod.data.get(), g.get(), mod.size, mod.stride, mod.pad, "max"), tol)
	elif isinstance(mod, M.MaxPool2D):
		x = mod.inData.get()
		y, mask = ops.maxpool2d_mask(x, mod.size, mod.stride, mod.pad)
		if not (np.array_equal(mod.data.get(), y) and np.array_equal(mod.mask.get(), mask)):
			failures.append("%s mask / output not bit-exact" % name)
		if g is not None and not np.array_equal(mod.grad.get(), ops.maxpool2d_mask_bwd(g.get(), x.shape, mask, mod.size, mod.stride, mod.pad)):
			failures.append("%s mask backward not bit-exact" % name)
	elif isinstance(mod, M.AvgPool2D):
		x = mod.inData.get()
		expect("fwd", mod.data.get(), ops.pool2d(x, mod.size, mod.stride, mod.pad, "avgWithPad"), tol)
		if g is not None:
			expect("bwd", mod.grad.get(), ops.pool2d_bwd(x, mod.data.get(), g.get(), mod.size, mod.stride, mod.pad, "avgWithPad"), tol)
	elif isinstance(mod, M.Linear):
		x, w = mod.inData.get(), mod.W.get()
		y = ops.gemm(x, w, transpB=mod.transpose) + (mod.b.get() if mod.useBias else 0.0)
		expect("fwd", mod.data.get(), y, tol_tc)
		if g is not None:
			dy = g.get()
			expect("dgrad", mod.grad.get(), ops.gemm(dy, w, transpB=not mod.transpose), tol_tc)
			expect("wgrad", mod.vars["W"].grad.get(), ops.gemm(x, dy, transpA=True) if not mod.transpose else ops.gemm(dy, x, transpA=True), tol_tc)
			if mod.useBias:
				expect("bgrad", mod.vars["b"].grad.get(), dy.sum(0), 2e-5)
	elif isinstance(mod, M.SoftMax):
		x = mod.inData.get()
		expect("fwd", mod.data.get(), ops.softmax(x.reshape(x.shape + (1, 1))).reshape(x.shape), tol)
		if g is not None:
			shape = x.shape + (1, 1)
			expect("bwd", mod.grad.get(), ops.softmax_bwd(mod.data.get().reshape(shape), g.get().reshape(shape)).reshape(x.shape), tol)
	elif isinstance(mod, M.Add):
		expect("fwd", mod.data.get(), sum(d.get().astype(np.float64) for d in mod.inData), 1e-6)
	elif isinstance(mod, M.Replicate):
		if g is not None:
			expect("bwd", mod.grad.get(), sum(d.get().astype(np.float64) for d in g), 1e-6)
	elif isinstance(mod, (M.Flatten, M.Identity)):
		pass
	else:
		failures.append("%s: no checker" % name)


def run_and_check(M, net, x, gy, floor=0.0):
	orig = record_backward(M)
	try:
		net.zeroGradParams()
		out = net(M.gpuarray.to_gpu(x))
		net.backward(M.gpuarray.to_gpu(gy))
	finally:
		M.Module.backward = orig

	failures = []
	import refshim
	leaves = list(refshim.leaves(net))
	for mod in leaves:
		check_leaf(M, mod, failures, floor=floor)
	return out, leaves, failures


# ================================================================================================ module-level tests
def test_linear_module_like_the_reference_test(M):
	# reference: Modules/Linear.py:114-140 (calcTest)
	rng = np.random.RandomState(0)
	insize, outsize = 5, 1
	x = rng.randn(5, insize).astype(np.float32)
	linear = M.Linear(insize, outsize, initscheme="he")
	linear(M.gpuarray.to_gpu(x))
	W, b = linear.W.get(), linear.b.get()
	assert relerr(linear.data.get(), x @ W + b) < 1e-3
	g = rng.randn(5, outsize).astype(np.float32)
	linear.backward(M.gpuarray.to_gpu(g))
	assert relerr(linear.grad.get(), g @ W.T) < 1e-3
	assert relerr(linear.vars["W"].grad.get(), x.T @ g) < 1e-3
	assert np.allclose(linear.vars["b"].grad.get(), g.sum(0), atol=1e-5)
	with pytest.raises(M.ModuleError):
		linear(M.gpuarray.to_gpu(rng.randn(5, insize + 1).astype(np.float32)))
	with pytest.raises(M.ModuleError):
		linear(M.gpuarray.to_gpu(x.astype(np.float16)))


def test_batchnorm_module_factor_schedule_and_eval(M):
	# running stats: factor = max(initFactor / numOfProps, minFactor) (reference: Modules/BatchNormND.py:47-63)
	rng = np.random.RandomState(1)
	bn = M.BatchNorm2D(5)
	mean, var = np.zeros(5), np.ones(5)
	for step in range(1, 4):
		x = rng.randn(16, 5, 4, 2).astype(np.float32) + step
		bn(M.gpuarray.to_gpu(x))
		f = max(1.0 / step, 0.1)
		mean = (1 - f) * mean + f * x.mean(axis=(0, 2, 3))
		var = (1 - f) * var + f * x.var(axis=(0, 2, 3), ddof=1)
	assert np.allclose(bn.mean.get().ravel(), mean, atol=1e-5) and np.allclose(bn.var.get().ravel(), var, rtol=1e-4)
	assert bn.savemean.shape == (1, 5, 1, 1)

	bn.evalMode()
	x = rng.randn(4, 5, 4, 2).astype(np.float32)
	y = bn(M.gpuarray.to_gpu(x)).get()
	want = ops.batchnorm_infer(x, bn.scale.get().ravel(), bn.bias.get().ravel(), bn.mean.get().ravel(), bn.var.get().ravel())
	assert np.allclose(y, want, atol=2e-5)

	inplace = M.BatchNorm2D(5, inplace=True)
	with pytest.raises(M.ModuleError, match="inplace"):
		inplace(M.gpuarray.to_gpu(x))


def test_activation_inplace_and_add_replicate_aliasing(M):
	rng = np.random.RandomState(2)
	x = rng.randn(3, 4, 5, 5).astype(np.float32)
	act = M.Activation(M.relu, inplace=True)
	d = M.gpuarray.to_gpu(x)
	assert act(d) is d and np.array_equal(d.get(), x * (x > 0))

	rep = M.Replicate(2)
	outs = rep(d)
	assert outs[0] is d and outs[1] is d                     # Replicate aliases (reference Q9)
	add = M.Add()
	y = add([d, M.gpuarray.to_gpu(x)])
	assert np.array_equal(y.get(), (np.float32(0) + d.get()) + x)
	add.backward(y)
	assert add.grad[0] is y and add.grad[1] is y             # Add.updateGrad aliases the same grad object
	rep.backward([y, y])
	assert np.array_equal(rep.grad.get(), y.get() + y.get())


def test_maxpool_module_mask_switch(M):
	rng = np.random.RandomState(3)
	x = np.maximum(rng.randn(2, 3, 9, 9), 0).astype(np.float32)
	a, b = M.MaxPool2D(3, 2), M.MaxPool2D(3, 2, useMask=True)
	ya, yb = a(M.gpuarray.to_gpu(x)), b(M.gpuarray.to_gpu(x))
	assert np.array_equal(ya.get(), yb.get()) and b.mask.dtype == np.int32
	g = rng.randn(*ya.shape).astype(np.float32)
	a.backward(M.gpuarray.to_gpu(g))
	b.backward(M.gpuarray.to_gpu(g))
	assert np.allclose(a.grad.get(), b.grad.get(), atol=1e-6)      # both route ties to the first maximum

	unpool = M.MaxUnpool2D(M.MaxPool2D(2, 2))
	pooled = unpool.maxpool2d(M.gpuarray.to_gpu(x[:, :, :8, :8].copy()))
	up = unpool(pooled)
	assert up.shape == (2, 3, 8, 8) and np.array_equal(unpool.maxpool2d(up).get(), pooled.get())


def test_deconv_and_conv1d_modules(M):
	rng = np.random.RandomState(4)
	deconv = M.Deconv2D(4, 6, 3, stride=2, pad=1, postpad=1, initscheme="he")
	x = rng.randn(2, 4, 5, 5).astype(np.float32)
	y = deconv(M.gpuarray.to_gpu(x))
	assert y.shape == deconv.dataShapeFrom(x.shape) == (2, 6, 10, 10)
	want = ops.conv2d_bwd_data(x, deconv.W.get(), None, deconv.b.get().ravel(), 2, 1, 1, 1)
	assert relerr(y.get(), want) < 1e-3
	g = rng.randn(*y.shape).astype(np.float32)
	deconv.backward(M.gpuarray.to_gpu(g))
	assert relerr(deconv.grad.get(), ops.conv2d(g, deconv.W.get(), None, 2, 1)) < 1e-3
	wg, bg = ops.conv2d_bwd_params(g, x, deconv.W.shape, 2, 1, withbias=True, deconv=True)
	assert relerr(deconv.vars["W"].grad.get(), wg) < 1e-3 and relerr(deconv.vars["b"].grad.get().ravel(), bg) < 1e-5

	conv = M.Conv1D(3, 5, 4, stride=2, pad=1, initscheme="he")
	x = rng.randn(2, 3, 17).astype(np.float32)
	y = conv(M.gpuarray.to_gpu(x))
	assert y.shape == conv.dataShapeFrom(x.shape)
	want = ops.conv2d(x[:, :, None, :], conv.W.get(), conv.b.get().ravel(), (1, 2), (0, 1))[:, :, 0]
	assert relerr(y.get(), want) < 1e-3
	conv.backward(M.gpuarray.to_gpu(rng.randn(*y.shape).astype(np.float32)))
	assert conv.grad.shape == x.shape


def test_module_gradcheck_small_net(M):
	# central differences through the GPU net itself, the reference's TestLib/GradientCheck.py:25-52 method
	rng = np.random.RandomState(5)
	np.random.seed(5)
	net = M.Sequential()
	net.append(M.Conv2D(2, 4, 3, pad=1, initscheme="he")).append(M.Activation(M.tanh)).append(M.AvgPool2D(2, 2))
	net.append(M.Flatten()).append(M.Linear(4 * 3 * 3, 5, initscheme="he")).append(M.SoftMax())
	x = rng.randn(3, 2, 6, 6).astype(np.float32)
	gy = rng.randn(3, 5).astype(np.float32)
	d = M.gpuarray.to_gpu(x)

	net.zeroGradParams()
	net(d)
	net.backward(M.gpuarray.to_gpu(gy))
	conv = net[0]
	W = conv.W.get()
	analytic = conv.vars["W"].grad.get()

	eps = 1e-2                       # TF32 products: use a step well above the 2^-11 rounding of the contraction
	for idx in [(0, 0, 1, 1), (2, 1, 0, 2), (3, 0, 2, 0)]:
		vals = []
		for sign in (1, -1):
			Wp = W.copy()
			Wp[idx] += sign * eps
			conv.W.set(Wp)
			vals.append(float((net(d).get().astype(np.float64) * gy).sum()))
		conv.W.set(W)
		numeric = (vals[0] - vals[1]) / (2 * eps)
		assert abs(numeric - analytic[idx]) < 2e-2 * max(1.0, abs(numeric)) + 5e-3


# ================================================================================================ whole nets
def test_lenet_teacher_forced_parity(M):
	from PuzzleLib.Models.Nets.LeNet import loadLeNet
	np.random.seed(1234)
	net = loadLeNet(None, initscheme=None)
	rng = np.random.RandomState(1234)
	x = rng.randn(64, 1, 28, 28).astype(np.float32)
	gy = (rng.randn(64, 10) * 1e-1).astype(np.float32)
	out, leaves, failures = run_and_check(M, net, x, gy)
	assert out.shape == (64, 10) and len(leaves) == 10
	assert failures == []

	# and end to end against the fp64 oracle net with the same weights (TF32 error compounds over 4 contractions only)
	ref = refnet.lenet()
	gconvs = [m for m in leaves if isinstance(m, (M.Conv2D, M.Linear))]
	rconvs = [l for l in ref.leaves() if isinstance(l, (refnet.Conv, refnet.Linear))]
	for gm, rl in zip(gconvs, rconvs):
		rl.W, rl.b = gm.W.get(), gm.b.get().ravel()
	y = ref.forward(x, np.float64)
	ref.backward(gy, np.float64, 1.0, 0.0)
	# max-pool / ReLU decisions flip on near-ties under TF32 rounding, so compare in the L2 sense end to end
	def l2err(got, want):
		return float(np.linalg.norm(got.astype(np.float64) - want) / np.linalg.norm(want))

	assert l2err(out.get(), y) < 3e-3
	assert l2err(net.grad.get(), ref.dx) < 5e-2
	assert l2err(gconvs[0].vars["W"].grad.get(), rconvs[0].dW) < 5e-2


def test_resnet50_teacher_forced_parity(M):
	from PuzzleLib.Models.Nets.ResNet import loadResNet
	np.random.seed(1234)
	net = loadResNet(None, "50", initscheme="he")
	assert net.numOfParams() == 25557032
	rng = np.random.RandomState(1234)
	x = rng.randn(2, 3, 224, 224).astype(np.float32)
	gy = (rng.randn(2, 1000) * 1e-3).astype(np.float32)
	out, leaves, failures = run_and_check(M, net, x, gy)

	assert out.shape == (2, 1000) and np.allclose(out.get().sum(axis=1), 1.0, atol=1e-4)
	assert sum(isinstance(m, M.Conv2D) for m in leaves) == 53 and sum(isinstance(m, M.BatchNorm2D) for m in leaves) == 53
	assert net["pool1"].data.shape == (2, 64, 55, 55)          # pool1 is 3x3 s2 p0 -> 55, not 56
	assert net.grad.shape == x.shape                           # conv1 dgrad IS computed
	assert failures == [], "\n".join(failures[:20])


@pytest.mark.parametrize("dtname,floor", [("float16", 3e-3), ("bfloat16", 2.5e-2)])
def test_resnet50_teacher_forced_parity_16bit(M, dtname, floor):
	"""BASELINE.json configs[3] / the reference's float16 path (Containers/Container.py:216-223 calcMode): every layer of ResNet-50
	in 16-bit storage against the float64 oracle fed with the SAME 16-bit inputs; the bar is the storage rounding (2^-11, 2^-8)"""
	from PuzzleLib.Models.Nets.ResNet import loadResNet
	from puzzlelib_b200 import seam, driver
	if dtname == "bfloat16" and driver.bfloat16 is None:
		pytest.skip("ml_dtypes.bfloat16 is not available")
	dt = np.dtype(np.float16) if dtname == "float16" else driver.bfloat16
	np.random.seed(4321)
	net = loadResNet(None, "50", initscheme="he")
	if dtname == "float16":
		net.calcMode(np.float16)
	else:
		seam.calcMode(net, dt)
	rng = np.random.RandomState(4321)
	x = rng.randn(2, 3, 224, 224).astype(dt)
	gy = (rng.randn(2, 1000) * 8.0).astype(dt)               # large enough that the gradients stay out of float16's subnormal range
	out, leaves, failures = run_and_check(M, net, x, gy, floor=floor)
	assert out.dtype == dt and out.shape == (2, 1000)
	assert sum(isinstance(m, M.Conv2D) for m in leaves) == 53
	assert failures == [], "\n".join(failures[:20])


def test_vgg16_forward_shapes_small_batch(M):
	from PuzzleLib.Models.Nets.VGG import loadVGG
	np.random.seed(7)
	net = loadVGG(None, "16", initscheme="he")
	rng = np.random.RandomState(7)
	x = rng.randn(1, 3, 224, 224).astype(np.float32)
	gy = (rng.randn(1, 1000) * 1e-3).astype(np.float32)
	out, leaves, failures = run_and_check(M, net, x, gy)
	assert out.shape == (1, 1000)
	assert failures == [], "\n".join(failures[:20])


# ================================================================================================ optimizer / global state
def test_momentum_sgd_global_state_matches_oracle(M):
	from PuzzleLib.Optimizers.MomentumSGD import MomentumSGD
	from PuzzleLib.Models.Nets.LeNet import loadLeNet
	np.random.seed(3)
	net = loadLeNet(None, initscheme=None)
	rng = np.random.RandomState(3)

	opt = MomentumSGD(learnRate=0.05, momRate=0.9)
	opt.setupOn(net, useGlobalState=True)
	flat = opt.globalVar[np.dtype(np.float32)]
	assert flat.data.size >= net.numOfParams()
	conv = net[0]
	assert flat.data.ptr <= conv.W.ptr < flat.data.ptr + flat.data.nbytes        # module vars are views of the flat buffer

	p0 = flat.data.get()
	mom = np.zeros_like(p0)
	x = rng.randn(8, 1, 28, 28).astype(np.float32)
	gy = rng.randn(8, 10).astype(np.float32)
	for _ in range(3):
		opt.zeroGradParams()
		net(M.gpuarray.to_gpu(x))
		net.backward(M.gpuarray.to_gpu(gy))
		g = flat.grad.get()
		assert np.abs(g).max() > 0
		opt.update()
		p0, mom = ops.sgd_momentum(p0, g, mom, 0.05, 0.9, dtype=np.float32)
		assert np.allclose(flat.data.get(), p0, atol=1e-6)
		off = (conv.W.ptr - flat.data.ptr) // 4
		assert np.array_equal(conv.W.get().ravel(), flat.data.get()[off:off + conv.W.size])


def test_step_graph_replay_matches_eager_steps(M):
	# driver.StepGraph: a step driven through the unchanged module API, captured once and replayed, must do
	# exactly what the same number of eager steps does -- same kernels, same buffers, deterministic kernels only on this net.
	from puzzlelib_b200 import driver
	from PuzzleLib.Optimizers.MomentumSGD import MomentumSGD

	def build():
		np.random.seed(21)
		net = M.Sequential()
		# minFactor = 1: the running-average factor max(1/n, minFactor) is the same scalar at every step, so the eager runs below
		# do the same number of steps as the graph (which otherwise keeps warming up until the factor settles)
		net.append(M.Conv2D(8, 16, 3, pad=1, initscheme="he")).append(M.BatchNorm2D(16, minFactor=1.0)).append(M.Activation(M.relu))
		net.append(M.MaxPool2D(2, 2)).append(M.Conv2D(16, 32, 1, useBias=False, initscheme="he")).append(M.AvgPool2D(4, 4))
		net.append(M.Flatten()).append(M.Linear(32 * 2 * 2, 10, initscheme="he")).append(M.SoftMax())
		opt = MomentumSGD(learnRate=1e-2, momRate=0.9)
		opt.setupOn(net, useGlobalState=True)
		return net, opt

	rng = np.random.RandomState(0)
	x = M.gpuarray.to_gpu(rng.randn(4, 8, 16, 16).astype(np.float32))
	gy = M.gpuarray.to_gpu((rng.randn(4, 10) * 0.1).astype(np.float32))

	def make_step(net, opt):
		def step():
			opt.zeroGradParams()
			net(x)
			net.backward(gy)
			opt.update()
			net.reset()
		return step

	warm, replays = 2, 3

	def eager():
		net, opt = build()
		step = make_step(net, opt)
		for _ in range(warm + replays):      # StepGraph runs the step `warmup` times eagerly; the capture pass only records
			step()
		driver.Device.synchronize()
		return net

	net1, net3 = eager(), eager()
	# split-K / statistics accumulate with fp32 red.add, so two eager runs agree to rounding only: that is the bar for the graph
	noise = max(relerr(net3.graph[0].W.get(), net1.graph[0].W.get()), relerr(net3.graph[7].W.get(), net1.graph[7].W.get()))
	tol = max(1e-5, 20 * noise)

	net2, opt2 = build()
	graph = driver.StepGraph(make_step(net2, opt2), warmup=warm)
	for _ in range(replays):
		graph.launch()
	graph.synchronize()

	w1 = net1.graph[0].W.get()
	w2 = net2.graph[0].W.get()
	assert np.abs(w1).max() > 0 and relerr(w2, w1) < tol
	assert relerr(net2.graph[7].W.get(), net1.graph[7].W.get()) < tol
	assert relerr(net2.graph[1].mean.get(), net1.graph[1].mean.get()) < tol
	graph.destroy()


def test_step_graph_warms_until_scalars_settle_and_redraws_dropout(M):
	"""a captured step freezes the scalars its kernels were handed.  StepGraph keeps warming the step up until they
	repeat (the batch-norm factor max(1/n, minFactor) does after 1/minFactor steps), refuses a step whose scalars never settle
	(Adam's bias-corrected rate), and dropout masks differ from replay to replay (the generator offset lives on the device)."""
	from puzzlelib_b200 import driver
	from PuzzleLib.Optimizers.Adam import Adam
	from PuzzleLib.Optimizers.MomentumSGD import MomentumSGD

	rng = np.random.RandomState(1)
	x = M.gpuarray.to_gpu(rng.randn(8, 4, 6, 6).astype(np.float32))
	gy = M.gpuarray.to_gpu(rng.randn(8, 4, 6, 6).astype(np.float32))

	np.random.seed(2)
	net = M.Sequential()
	net.append(M.Conv2D(4, 4, 3, pad=1, initscheme="he")).append(M.BatchNorm2D(4, minFactor=0.25)).append(M.Dropout(p=0.5))
	opt = MomentumSGD(learnRate=1e-2, momRate=0.9)
	opt.setupOn(net, useGlobalState=True)

	masks = []

	def step():
		opt.zeroGradParams()
		net(x)
		net.backward(gy)
		opt.update()

	graph = driver.StepGraph(step, warmup=2)
	assert graph.warmupRuns >= 5                   # factor 1, 1/2, 1/3, 1/4 = minFactor, and one repeat
	for _ in range(3):
		graph.launch()
		graph.synchronize()
		masks.append(net.graph[2].data.get() == 0)
	assert not np.array_equal(masks[0], masks[1]) and not np.array_equal(masks[1], masks[2])
	# host reads wait for replays in flight
	graph.launch()
	w = net.graph[0].W.get()
	assert np.isfinite(w).all()
	graph.destroy()

	adam = Adam(alpha=1e-3)
	lin = M.Linear(6, 3, initscheme="he")
	adam.setupOn(lin, useGlobalState=True)
	xs, gs = M.gpuarray.to_gpu(rng.randn(5, 6).astype(np.float32)), M.gpuarray.to_gpu(rng.randn(5, 3).astype(np.float32))

	def adamStep():
		adam.zeroGradParams()
		lin(xs)
		lin.backward(gs)
		adam.update()

	with pytest.raises(RuntimeError, match="Adam"):
		driver.StepGraph(adamStep, warmup=2, maxWarmup=6)


def test_checkpoint_round_trip_through_the_reference_save_and_load(M):
	"""the reference's own Module / Container / Optimizer save + load above the seam.  The file object is an
	in-memory stand-in with h5py's call surface (tests/h5standin.py; h5py is not in this image) -- the backend's part of a
	checkpoint is GPUArray.get / set, parameters shared through varlinks, batch-norm running statistics as attrs."""
	import h5standin
	from PuzzleLib.Containers.Sequential import Sequential
	from PuzzleLib.Modules import Conv2D, BatchNorm2D, Activation, MaxPool2D, Flatten, Linear
	from PuzzleLib.Modules.Activation import relu
	from PuzzleLib.Optimizers.MomentumSGD import MomentumSGD
	from PuzzleLib.Backend import gpuarray

	def build(seed):
		np.random.seed(seed)
		net = Sequential(name="ckpt")
		net.append(Conv2D(3, 8, 3, pad=1, name="conv"))
		net.append(BatchNorm2D(8, name="bn"))
		net.append(Activation(relu, name="relu"))
		net.append(MaxPool2D(name="pool"))
		net.append(Flatten(name="flat"))
		net.append(Linear(8 * 4 * 4, 5, name="fc"))
		return net

	rng = np.random.RandomState(0)
	x = gpuarray.to_gpu(rng.randn(6, 3, 8, 8).astype(np.float32))
	g = gpuarray.to_gpu(rng.randn(6, 5).astype(np.float32))

	src = build(1)
	opt = MomentumSGD(learnRate=0.1, momRate=0.9)
	opt.setupOn(src, useGlobalState=True)
	for _ in range(3):                       # a few steps: running statistics and momentum state become non-trivial
		src.zeroGradParams()
		src(x)
		src.backward(g)
		opt.update()

	hdf, ohdf = h5standin.File(), h5standin.File()
	src.save(hdf=hdf)
	opt.save(ohdf)
	assert hdf.closed == 1 and {"params", "links", "attrs"} <= set(hdf.keys())

	dst = build(2)
	assert not np.allclose(dst["conv"].vars["W"].data.get(), src["conv"].vars["W"].data.get())
	dst.load(hdf)
	for mod in ("conv", "bn", "fc"):
		for vname, var in src[mod].vars.items():
			assert np.array_equal(var.data.get(), dst[mod].vars[vname].data.get()), (mod, vname)
	for aname, attr in src["bn"].attrs.items():
		assert np.array_equal(attr.get(), dst["bn"].attrs[aname].get()), aname

	src.evalMode()
	dst.evalMode()
	assert np.array_equal(src(x).get(), dst(x).get())

	opt2 = MomentumSGD(learnRate=0.5, momRate=0.1)
	opt2.setupOn(dst, useGlobalState=True)
	opt2.load(ohdf)
	assert opt2.learnRate == opt.learnRate and opt2.momRate == opt.momRate and opt2.t == opt.t == 3
	for sname, state in opt.states.items():
		for ename, entity in state.items():
			assert np.array_equal(entity.get(), opt2.states[sname][ename].get()), (sname, ename)
