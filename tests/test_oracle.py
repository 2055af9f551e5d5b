"""CPU tier: the oracle against (a) golden vectors produced by the reference itself (tools/gen_golden.py, reference
numpy CPU backend) and (b) finite-difference / adjoint identities for the backward ops the reference cannot run on
CPU (the method of the reference's TestLib/GradientCheck.py:25-52)."""
import os

import numpy as np
import pytest

from oracle import ops, refnet

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_cpu_ops.npz"))


@pytest.mark.parametrize("i", range(5))
def test_conv_fwd_matches_reference_cpu(i):
	size, stride, pad, dil, hasbias = (int(v) for v in GOLD["conv%d_cfg" % i])
	b = GOLD["conv%d_b" % i].ravel() if hasbias else None
	y = ops.conv2d(GOLD["conv%d_x" % i], GOLD["conv%d_w" % i], b, stride, pad, dil)
	assert y.shape == GOLD["conv%d_y" % i].shape
	assert np.allclose(y, GOLD["conv%d_y" % i], atol=1e-4, rtol=1e-5)


@pytest.mark.parametrize("i", range(5))
def test_pool_fwd_matches_reference_cpu(i):
	size, stride, pad, kind = (int(v) for v in GOLD["pool%d_cfg" % i])
	y = ops.pool2d(GOLD["pool%d_x" % i], size, stride, pad, "max" if kind == 0 else "avgWithPad")
	assert np.allclose(y, GOLD["pool%d_y" % i], atol=1e-6)
	if kind == 0:
		ym, mask = ops.maxpool2d_mask(GOLD["pool%d_x" % i], size, stride, pad)
		assert np.array_equal(ym, GOLD["pool%d_y" % i])
		x = GOLD["pool%d_x" % i]
		flat = x.reshape(x.shape[0], x.shape[1], -1)
		assert np.array_equal(np.take_along_axis(flat, mask.reshape(*mask.shape[:2], -1).astype(np.int64), 2).reshape(ym.shape), ym)


@pytest.mark.parametrize("i", range(2))
def test_batchnorm_infer_matches_reference_cpu(i):
	g = lambda key: GOLD["bn%d_%s" % (i, key)]
	y = ops.batchnorm_infer(g("x"), g("scale"), g("bias"), g("mean"), g("var"))
	assert np.allclose(y, g("y"), atol=1e-5)


@pytest.mark.parametrize("i", range(2))
def test_linear_matches_reference_cpu(i):
	g = lambda key: GOLD["lin%d_%s" % (i, key)]
	transpose = bool(g("transpose")[0])
	y = ops.gemm(g("x"), g("w"), transpB=transpose) + g("b")
	assert np.allclose(y, g("y"), atol=1e-5)
	assert np.allclose(ops.gemm(g("g"), g("w"), transpB=not transpose), g("dx"), atol=1e-5)
	dw = ops.gemm(g("x"), g("g"), transpA=True) if not transpose else ops.gemm(g("g"), g("x"), transpA=True)
	assert np.allclose(dw, g("dw"), atol=1e-5)
	if not transpose:
		assert np.allclose(ops.matsum(g("g"), 0), g("db"), atol=1e-5)


@pytest.mark.parametrize("kind", ["sigmoid", "tanh", "relu", "leakyRelu", "elu", "softPlus", "clip"])
def test_activation_matches_reference_cpu(kind):
	y = ops.activation(kind, GOLD["act_x"])
	assert np.allclose(y, GOLD["act_%s_y" % kind], atol=1e-6)
	dx = ops.activation_bwd(kind, GOLD["act_g"], GOLD["act_%s_y" % kind])
	assert np.allclose(dx, GOLD["act_%s_dx" % kind], atol=1e-6)


# ------------------------------------------------------------------------------------------ backward ops: gradient checks
def numgrad(f, x, eps=1e-6):
	g = np.zeros_like(x)
	it = np.nditer(x, flags=["multi_index"])
	while not it.finished:
		idx = it.multi_index
		old = x[idx]
		x[idx] = old + eps
		fp = f()
		x[idx] = old - eps
		fm = f()
		x[idx] = old
		g[idx] = (fp - fm) / (2 * eps)
		it.iternext()
	return g


@pytest.mark.parametrize("cfg", [(1, 1, 0, 1), (2, 1, 1, 1), (1, 2, 2, 1), (2, 0, 1, 2), (1, 1, 1, 3)])
def test_conv_backward_gradcheck(cfg):
	stride, pad, dil, groups = cfg
	rng = np.random.RandomState(0)
	x = rng.randn(2, 2 * groups, 6, 5)
	w = rng.randn(3 * groups, 2, 3, 2)
	y = ops.conv2d(x, w, None, stride, pad, dil, groups)
	gy = rng.randn(*y.shape)

	loss = lambda: float((ops.conv2d(x, w, None, stride, pad, dil, groups) * gy).sum())
	dx = ops.conv2d_bwd_data(gy, w, x.shape, stride=stride, pad=pad, dilation=dil, groups=groups)
	dw, db = ops.conv2d_bwd_params(x, gy, w.shape, stride, pad, dil, groups, withbias=True)
	assert np.allclose(dx, numgrad(loss, x), atol=1e-5)
	assert np.allclose(dw, numgrad(loss, w), atol=1e-5)
	assert np.allclose(db, gy.sum(axis=(0, 2, 3)))

	old = rng.randn(*w.shape)
	dw2 = ops.conv2d_bwd_params(x, gy, w.shape, stride, pad, dil, groups, wgrad=old, scale=0.5, momentum=0.25)
	assert np.allclose(dw2, 0.5 * dw + 0.25 * old)


def test_deconv_is_conv_transpose():
	# <deconv(x), y> == <x, conv(y)>: Deconv forward is conv backward-data (Backend/Dnn.py:211-215)
	rng = np.random.RandomState(1)
	w = rng.randn(4, 3, 3, 3)
	x = rng.randn(2, 4, 5, 5)
	out = ops.conv2d_bwd_data(x, w, None, stride=2, pad=1, postpad=1)
	assert out.shape == (2, 3, 10, 10)
	y = rng.randn(*out.shape)
	assert np.isclose((out * y).sum(), (x * ops.conv2d(y, w, None, 2, 1)).sum())


def test_batchnorm_backward_gradcheck():
	rng = np.random.RandomState(2)
	x, scale, bias = rng.randn(3, 4, 2, 3), rng.randn(4), rng.randn(4)
	gy = rng.randn(*x.shape)
	zero, one = np.zeros(4), np.ones(4)
	y, mu, inv, newmean, newvar = ops.batchnorm_train(x, scale, bias, zero, one, 1e-5, 1.0)
	assert np.allclose(newmean, x.mean(axis=(0, 2, 3)))
	assert np.allclose(newvar, x.var(axis=(0, 2, 3), ddof=1))

	# the reference test's own closed form (Cuda/Wrappers/CuDnnNorm.py:50-64)
	norm = 3 * 2 * 3
	hm, hi, hs = mu.reshape(1, 4, 1, 1), inv.reshape(1, 4, 1, 1), scale.reshape(1, 4, 1, 1)
	bg = gy.sum(axis=(0, 2, 3), keepdims=True)
	vg = -0.5 * (gy * (x - hm)).sum(axis=(0, 2, 3), keepdims=True) * hs * hi ** 3
	want = gy * hs * hi + (2 * vg * (x - hm) + (-hi * bg * hs)) / norm

	dx, dscale, dbias = ops.batchnorm_bwd(x, gy, scale, mu, inv)
	assert np.allclose(dx, want)
	loss = lambda: float((ops.batchnorm_train(x, scale, bias, zero, one, 1e-5, 1.0)[0] * gy).sum())
	assert np.allclose(dx, numgrad(loss, x), atol=1e-5)
	assert np.allclose(dscale, numgrad(loss, scale), atol=1e-5)
	assert np.allclose(dbias, numgrad(loss, bias), atol=1e-5)


@pytest.mark.parametrize("mode", ["max", "avgWithPad", "avgNoPad"])
def test_pool_backward_gradcheck(mode):
	rng = np.random.RandomState(3)
	x = rng.randn(2, 2, 7, 6)
	y = ops.pool2d(x, 3, 2, 1, mode)
	gy = rng.randn(*y.shape)
	loss = lambda: float((ops.pool2d(x, 3, 2, 1, mode) * gy).sum())
	assert np.allclose(ops.pool2d_bwd(x, y, gy, 3, 2, 1, mode), numgrad(loss, x), atol=1e-5)


def test_maxpool_mask_semantics():
	# ties: first strict maximum of the row-major scan wins (Cuda/Kernels/Pool.py:34-42); backward gathers by mask
	x = np.zeros((1, 1, 4, 4), np.float32)
	y, mask = ops.maxpool2d_mask(x, 2, 2, 0)
	assert np.array_equal(mask[0, 0], np.array([[0, 2], [8, 10]], np.int32))
	x[0, 0, 1, 1] = 1.0
	y, mask = ops.maxpool2d_mask(x, 3, 1, 1)
	assert mask[0, 0, 0, 0] == 5 and y[0, 0, 0, 0] == 1.0 and mask[0, 0, 3, 3] == 10
	gy = np.arange(16, dtype=np.float32).reshape(1, 1, 4, 4)
	dx = ops.maxpool2d_mask_bwd(gy, x.shape, mask, 3, 1, 1)
	assert dx.sum() == gy.sum() and dx[0, 0, 1, 1] == gy[0, 0, :3, :3].sum()
	up = ops.maxunpool2d(y, x.shape, mask)
	assert up[0, 0, 1, 1] == 1.0
	assert np.array_equal(ops.maxunpool2d_bwd(up, y.shape, mask), y)


def test_softmax_backward_gradcheck():
	rng = np.random.RandomState(4)
	for mode in ("spatial", "perActivation"):
		x = rng.randn(3, 5, 2, 2)
		y = ops.softmax(x, mode)
		axes = 1 if mode == "spatial" else (1, 2, 3)
		assert np.allclose(y.sum(axis=axes), 1.0)
		gy = rng.randn(*x.shape)
		loss = lambda: float((ops.softmax(x, mode) * gy).sum())
		assert np.allclose(ops.softmax_bwd(y, gy, mode), numgrad(loss, x), atol=1e-6)


def test_gelu_derivative_is_the_reference_formula_not_the_true_one():
	x = np.linspace(-2, 2, 9)
	d = ops.activation_bwd("gelu", np.ones_like(x), x)
	true = 0.5 * (1 + ops._erf(x / np.sqrt(2))) + x / np.sqrt(2 * np.pi) * np.exp(-0.5 * x * x)
	assert not np.allclose(d, true)          # 1/sqrt(pi), sic (ElementWise.py:469-475)
	assert np.allclose(d[4], true[4])


def test_refnet_lenet_gradcheck_and_accumulate():
	rng = np.random.RandomState(5)
	net = refnet.init_he(refnet.lenet(), seed=7)
	x = rng.randn(2, 1, 28, 28)
	gy = rng.randn(2, 10)
	net.forward(x, np.float64)
	net.backward(gy, np.float64, scale=1.0, momentum=0.0)
	first = net.layers[0]
	W = first.W.astype(np.float64)
	first.W = W
	idxs = [(0, 0, 0, 0), (3, 0, 1, 2), (15, 0, 2, 2)]
	for idx in idxs:
		old = W[idx]
		W[idx] = old + 1e-5
		fp = float((net.forward(x, np.float64) * gy).sum())
		W[idx] = old - 1e-5
		fm = float((net.forward(x, np.float64) * gy).sum())
		W[idx] = old
		assert np.isclose(first.dW[idx], (fp - fm) / 2e-5, rtol=1e-4, atol=1e-7)
	# Sequential.backward semantics: momentum = 1 accumulates into the existing gradient
	before = first.dW.copy()
	net.forward(x, np.float64)
	net.backward(gy, np.float64, scale=1.0, momentum=1.0)
	assert np.allclose(first.dW, 2 * before)


def test_refnet_resnet50_shapes():
	net = refnet.resnet50()
	convs = [l for l in net.leaves() if isinstance(l, refnet.Conv)]
	bns = [l for l in net.leaves() if isinstance(l, refnet.BatchNorm)]
	assert len(convs) == 53 and len(bns) == 53
	nparams = sum(l.W.size for l in convs) + sum(2 * l.scale.size for l in bns) + 2048 * 1000 + 1000
	assert nparams == 25557032      # 25.56 M parameters


def test_lstm_oracle_gradients_match_finite_differences():
	# the LSTM restatement follows the reference's host loop (Cuda/Wrappers/CuDnnRnn.py:178-300); its backward is pinned here
	# against central differences of its own forward (fp64)
	rng = np.random.RandomState(5)
	T, B, insz, H = 3, 2, 4, 3
	params = {}
	for g in "ifco":
		params["w" + g] = rng.randn(H, insz) * 0.5
		params["r" + g] = rng.randn(H, H) * 0.5
		params["bw" + g] = rng.randn(H) * 0.1
		params["br" + g] = rng.randn(H) * 0.1
	x = rng.randn(T, B, insz)
	dy = rng.randn(T, B, H)

	out, cache = ops.lstm_forward(x, params)
	dx, dp = ops.lstm_backward(x, params, cache, dy)

	def loss(xv, pv):
		return float((ops.lstm_forward(xv, pv)[0] * dy).sum())

	eps = 1e-6
	for idx in [(0, 0, 0), (1, 1, 2), (2, 0, 3)]:
		xp, xm = x.copy(), x.copy()
		xp[idx] += eps
		xm[idx] -= eps
		assert abs((loss(xp, params) - loss(xm, params)) / (2 * eps) - dx[idx]) < 1e-6
	for name, idx in [("wi", (0, 1)), ("rf", (2, 0)), ("wc", (1, 3)), ("ro", (1, 1)), ("bwf", (2, )), ("brc", (0, ))]:
		pp, pm = {k: v.copy() for k, v in params.items()}, {k: v.copy() for k, v in params.items()}
		pp[name][idx] += eps
		pm[name][idx] -= eps
		assert abs((loss(x, pp) - loss(x, pm)) / (2 * eps) - dp[name][idx]) < 1e-6


def test_rnn_oracle_gradients_match_finite_differences():
	rng = np.random.RandomState(6)
	T, B, insz, H = 4, 2, 3, 3
	params = {"wi": rng.randn(H, insz) * 0.5, "ri": rng.randn(H, H) * 0.5, "bwi": rng.randn(H) * 0.1, "bri": rng.randn(H) * 0.1}
	x, dy = rng.randn(T, B, insz), rng.randn(T, B, H)
	out = ops.rnn_forward(x, params, "tanh")
	dx, dp = ops.rnn_backward(x, params, out, dy, "tanh")
	eps = 1e-6

	def loss(xv, pv):
		return float((ops.rnn_forward(xv, pv, "tanh") * dy).sum())

	xp, xm = x.copy(), x.copy()
	xp[1, 0, 2] += eps
	xm[1, 0, 2] -= eps
	assert abs((loss(xp, params) - loss(xm, params)) / (2 * eps) - dx[1, 0, 2]) < 1e-6
	pp, pm = {k: v.copy() for k, v in params.items()}, {k: v.copy() for k, v in params.items()}
	pp["ri"][1, 2] += eps
	pm["ri"][1, 2] -= eps
	assert abs((loss(x, pp) - loss(x, pm)) / (2 * eps) - dp["ri"][1, 2]) < 1e-6


def test_gru_oracle_gradients_match_finite_differences():
	rng = np.random.RandomState(8)
	T, B, insz, H = 3, 2, 4, 3
	params = {}
	for g in "rih":
		params["w" + g] = rng.randn(H, insz) * 0.5
		params["r" + g] = rng.randn(H, H) * 0.5
		params["bw" + g] = rng.randn(H) * 0.1
		params["br" + g] = rng.randn(H) * 0.1
	x, dy = rng.randn(T, B, insz), rng.randn(T, B, H)
	for reverse in (False, True):
		out, cache = ops.gru_forward(x, params, reverse=reverse)
		dx, dp = ops.gru_backward(x, params, cache, dy, reverse=reverse)

		def loss(xv, pv):
			return float((ops.gru_forward(xv, pv, reverse=reverse)[0] * dy).sum())

		eps = 1e-6
		xp, xm = x.copy(), x.copy()
		xp[1, 1, 2] += eps
		xm[1, 1, 2] -= eps
		assert abs((loss(xp, params) - loss(xm, params)) / (2 * eps) - dx[1, 1, 2]) < 1e-6
		for name, idx in [("wr", (0, 1)), ("rh", (2, 0)), ("ri", (1, 1)), ("brh", (2, )), ("bwh", (0, )), ("brr", (1, ))]:
			pp, pm = {k: v.copy() for k, v in params.items()}, {k: v.copy() for k, v in params.items()}
			pp[name][idx] += eps
			pm[name][idx] -= eps
			assert abs((loss(x, pp) - loss(x, pm)) / (2 * eps) - dp[name][idx]) < 1e-6


def test_lstm_oracle_reverse_direction_is_the_time_flipped_forward():
	rng = np.random.RandomState(2)
	T, B, insz, H = 4, 2, 3, 3
	params = {}
	for g in "ifco":
		params["w" + g], params["r" + g] = rng.randn(H, insz) * 0.5, rng.randn(H, H) * 0.5
		params["bw" + g], params["br" + g] = rng.randn(H) * 0.1, rng.randn(H) * 0.1
	x, dy = rng.randn(T, B, insz), rng.randn(T, B, H)
	out_r, cache_r = ops.lstm_forward(x, params, reverse=True)
	out_f, cache_f = ops.lstm_forward(x[::-1], params)
	assert np.allclose(out_r, out_f[::-1])
	dx_r, dp_r = ops.lstm_backward(x, params, cache_r, dy)
	dx_f, dp_f = ops.lstm_backward(x[::-1], params, cache_f, dy[::-1])
	assert np.allclose(dx_r, dx_f[::-1])
	assert all(np.allclose(dp_r[k], dp_f[k]) for k in dp_r)


def test_cross_entropy_oracle_gradient_is_the_negated_derivative_of_the_error():
	# the reference's costs emit ascent-direction gradients: grad = -d(error / N) / d(scores)  
	rng = np.random.RandomState(4)
	N, C, S = 3, 5, 2
	scores, labels = rng.randn(N, C, S), rng.randint(0, C, (N, S))
	weights = rng.rand(C) + 0.5
	err, grad = ops.cross_entropy(scores, labels)
	eps = 1e-6
	for idx in [(0, 1, 0), (2, 4, 1), (1, 0, 1)]:
		sp, sm = scores.copy(), scores.copy()
		sp[idx] += eps
		sm[idx] -= eps
		num = (ops.cross_entropy(sp, labels)[0] - ops.cross_entropy(sm, labels)[0]) / (2 * eps)
		assert abs(-num * S / N - grad[idx]) < 1e-6
	# the weighted variant scales the gradient by the weight of the OUTPUT class c, not of the label (Costs.py:147), and
	# the error by the weight of the label
	werr, wgrad = ops.cross_entropy(scores, labels, weights)
	assert np.allclose(wgrad, grad * weights.reshape(1, C, 1))
	assert werr != err
	# a 2-d problem is the spatial one with a single position
	e2, g2 = ops.cross_entropy(scores[:, :, 0], labels[:, 0])
	e4, g4 = ops.cross_entropy(scores[:, :, :1], labels[:, :1])
	assert abs(e2 - e4) < 1e-12 and np.allclose(g2, g4[:, :, 0])


def test_optimizer_update_oracles():
	rng = np.random.RandomState(6)
	p, g, m = rng.randn(7), rng.randn(7), rng.randn(7)
	p1, m1 = ops.nesterov_update(p, g, m, 0.1, 0.9)
	# Nesterov in the reference's form: classic momentum step followed by a look-ahead of the NEW momentum
	assert np.allclose(m1, 0.9 * m + 0.1 * g) and np.allclose(p1, p + 0.9 * m1 + 0.1 * g)
	a, s = rng.randn(7), rng.rand(7) + 1.0
	p2, a2, s2 = ops.adam_update(p, g, a, s, 0.01, 0.1, 0.001, 1e-8)
	assert np.allclose(a2, 0.9 * a + 0.1 * g) and np.allclose(s2, 0.999 * s + 0.001 * g * g)
	assert np.allclose(p2, p + 0.01 * a2 / (np.sqrt(s2) + 1e-8))


def test_lrn_oracle_gradient_matches_finite_differences_for_odd_windows():
	rng = np.random.RandomState(10)
	x, g = rng.randn(2, 6, 4, 5), rng.randn(2, 6, 4, 5)
	for across in (True, False):
		y, dx = ops.lrn(x, 3, 0.7, 0.75, 2.0, across, grad=g)
		eps = 1e-6
		for idx in [(0, 0, 0, 0), (1, 3, 2, 4), (0, 5, 3, 1)]:
			xp, xm = x.copy(), x.copy()
			xp[idx] += eps
			xm[idx] -= eps
			num = ((ops.lrn(xp, 3, 0.7, 0.75, 2.0, across) - ops.lrn(xm, 3, 0.7, 0.75, 2.0, across)) * g).sum() / (2 * eps)
			assert abs(num - dx[idx]) < 1e-6


def _train_gold():
	import os
	return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cpu_train.npz"))


def test_optimizer_and_dropout_oracles_match_the_reference_cpu_kernels():
	# tests/golden/ref_cpu_train.npz: outputs of the reference's own CPU kernels and optimizer objects (tools/gen_golden_train.py)
	g = _train_gold()
	w, dw, mom, mg, ms = (g[k] for k in ("w", "dw", "mom", "mg", "ms"))
	f32 = np.float32
	pw, pm = ops.nesterov_update(w, dw, mom, f32(0.01), f32(0.9), dtype=f32)
	assert np.allclose(pw, g["nesterov_w"], rtol=0, atol=1e-6) and np.allclose(pm, g["nesterov_mom"], rtol=0, atol=1e-6)
	cm = f32(0.9) * mom + f32(0.01) * dw
	assert np.allclose(cm, g["classic_mom"], rtol=0, atol=1e-6) and np.allclose(w + cm, g["classic_w"], rtol=0, atol=1e-6)
	lr, fix1, fix2, eps = g["adam_args"]
	aw, amg, ams = ops.adam_update(w, dw, mg, ms, lr, fix1, fix2, eps, dtype=f32)
	assert np.allclose(aw, g["adam_w"], rtol=0, atol=1e-6) and np.allclose(amg, g["adam_mg"], rtol=0, atol=1e-6)
	assert np.allclose(ams, g["adam_ms"], rtol=0, atol=1e-6)

	x, words, v, p = g["drop_x"], g["drop_words"], g["drop_v"][0], g["drop_p"][0]
	# (the reference's gcc build may multiply by 1/p instead of dividing: one ulp)
	assert np.allclose((x * (words.reshape(x.shape) < v) / p).astype(f32), g["drop_y"], rtol=3e-7, atol=0)
	assert np.array_equal(g["drop_y"] == 0, ~(words.reshape(x.shape) < v) | (x == 0))
	mw = g["drop2d_words"].reshape(x.shape[0], x.shape[1], 1, 1)
	assert np.allclose((x * (mw < v) / p).astype(f32), g["drop2d_y"], rtol=3e-7, atol=0)

	# the optimizer objects: three updates with the reference's bias-corrected Adam step size etc. (Optimizers/Adam.py:36-45)
	grads = g["opt_grads"]
	wa, a, s2 = w.astype(np.float64), np.zeros_like(w, np.float64), np.zeros_like(w, np.float64)
	wn, mn = w.astype(np.float64), np.zeros_like(w, np.float64)
	wm, mm = w.astype(np.float64), np.zeros_like(w, np.float64)
	for t, gr in enumerate(grads, start=1):
		lr = 1e-2 * np.sqrt(1.0 - 0.999 ** t) / (1.0 - 0.9 ** t)
		wa, a, s2 = ops.adam_update(wa, gr, a, s2, lr, 0.1, 0.001, 1e-8)
		wn, mn = ops.nesterov_update(wn, gr, mn, 1e-1, 0.9)
		mm = 0.9 * mm + 1e-1 * gr
		wm = wm + mm
	assert np.allclose(wa, g["opt_adam_w"], atol=2e-5) and np.allclose(wn, g["opt_nesterov_w"], atol=2e-5)
	assert np.allclose(wm, g["opt_momentum_w"], atol=2e-5)
