"""CPU tier: the numpy oracle (oracle/ops.py) held to the outputs of the reference's OWN CUDA backend.

tests/golden/ref_cuda_ops.npz was produced on a B200 by tools/gen_golden_cuda.py --impl ref: the seeded case table of
tests/golden_cases.py run through `PuzzleLib.Backend.*` with PuzzleLib's cuDNN 9 / cuBLAS 12 / NVRTC backend (built from the
unmodified reference tree by baseline/build_ref.py).  It holds inputs (`in_*`) and outputs of forward AND backward of
every operator family on the path, so this is what pins the oracle for the functions the reference's CPU backend cannot
run (conv dgrad / wgrad / bgrad, batch-norm train / backward incl. the running variance, max-pool backward tie routing,
softmax backward, LRN).  On this stack cuDNN computes float32 convolutions in full fp32 (error vs fp64 ~1e-7), so the
fp32 bar here is 2e-5 relative; float16 cases are held to the reference's own 16-bit bar (1e-2 absolute scaled).
"""
import os

import numpy as np
import pytest

from oracle import ops
import golden_cases as gc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cuda_ops.npz")


@pytest.fixture(scope="module")
def gold():
	data = np.load(GOLDEN)
	return {key: data[key] for key in data.files}


def rel(got, want):
	got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
	return float(np.abs(got - want).max() / (np.abs(want).max() + 1e-30))


def case(gold, name):
	prefix = name + "/"
	out = {key[len(prefix):]: val for key, val in gold.items() if key.startswith(prefix)}
	assert out, "golden file has no case %s" % name
	return out


def tol(name):
	return 4e-3 if name.endswith("_f16") else 2e-5


CONV2D = [n for n in gc.CASES if n.startswith("conv") and not n.startswith("conv3d")]


@pytest.mark.parametrize("name", CONV2D)
def test_conv2d_forward_backward(gold, name):
	g = case(gold, name)
	_, _, _, _, _, stride, pad, dilation, groups, withbias = gc.CONV_GEOMETRIES[name.rsplit("_", 1)[0]]
	x, W, dy = g["in_x"], g["in_W"], g["in_dy"]
	b = g["in_b"].ravel() if withbias else None
	t = tol(name)

	assert rel(ops.conv2d(x, W, b, stride, pad, dilation, groups), g["y"]) < t
	assert rel(ops.conv2d_bwd_data(dy, W, x.shape, None, stride, pad, dilation, 0, groups), g["dx"]) < t

	want = ops.conv2d_bwd_params(x, dy, W.shape, stride, pad, dilation, groups, withbias=withbias)
	assert rel(want[0] if withbias else want, g["wgrad"]) < t
	acc = ops.conv2d_bwd_params(x, dy, W.shape, stride, pad, dilation, groups, withbias=withbias, wgrad=g["in_w0"],
								bgrad=g["in_b0"] if withbias else None, scale=0.5, momentum=0.9)
	assert rel(acc[0] if withbias else acc, g["wacc"]) < t
	if withbias:
		assert rel(want[1], g["bgrad"].ravel()) < t
		assert rel(acc[1], g["bacc"].ravel()) < t


def test_conv3d_forward_backward(gold):
	g = case(gold, "conv3d_f32")
	_, _, _, _, _, stride, pad, dilation, groups, _ = gc.CONV_GEOMETRIES["conv3d"]
	x, W, dy, b = g["in_x"], g["in_W"], g["in_dy"], g["in_b"].ravel()
	assert rel(ops.conv3d(x, W, b, stride, pad, dilation, groups), g["y"]) < 2e-5
	assert rel(ops.conv3d_bwd_data(dy, W, x.shape, stride, pad, dilation, groups), g["dx"]) < 2e-5
	dw, db = ops.conv3d_bwd_params(x, dy, W.shape, stride, pad, dilation, groups)
	assert rel(dw, g["wgrad"]) < 2e-5 and rel(db, g["bgrad"].ravel()) < 2e-5


@pytest.mark.parametrize("name", [n for n in gc.CASES if n.startswith("deconv")])
def test_deconv2d_forward_backward(gold, name):
	g = case(gold, name)
	_, _, _, _, _, stride, pad, dilation, postpad, groups, withbias = gc.DECONV_GEOMETRIES[name.rsplit("_", 1)[0]]
	x, W, dy = g["in_x"], g["in_W"], g["in_dy"]
	b = g["in_b"].ravel() if withbias else None
	t = tol(name)

	# deconv forward IS conv backward-data (+ bias), its data gradient IS conv forward, its filter gradient swaps the roles
	assert rel(ops.conv2d_bwd_data(x, W, None, b, stride, pad, dilation, postpad, groups), g["y"]) < t
	assert rel(ops.conv2d(dy, W, None, stride, pad, dilation, groups), g["dx"]) < t
	want = ops.conv2d_bwd_params(dy, x, W.shape, stride, pad, dilation, groups, withbias=withbias, deconv=True)
	assert rel(want[0] if withbias else want, g["wgrad"]) < t
	if withbias:
		assert rel(want[1], g["bgrad"].ravel()) < t


@pytest.mark.parametrize("name", [n for n in gc.CASES if n.startswith("bn")])
def test_batchnorm_train_backward_inference(gold, name):
	g = case(gold, name)
	x, dy = g["in_x"], g["in_dy"]
	if name.startswith("bnperact"):
		pass      # (N, C, 1, 1) data: per-activation statistics are the spatial ones
	factor = {"bn2d_big_f32": 0.1, "bn1d_f32": 1.0}.get(name, 0.3)
	t = tol(name)

	y, mu, inv, rmean, rvar = ops.batchnorm_train(x, g["in_scale"].ravel(), g["in_bias"].ravel(), g["in_mean0"].ravel(),
												  g["in_var0"].ravel(), 1e-5, factor)
	assert rel(y, g["y"]) < t
	assert rel(mu, g["savemean"].ravel()) < 2e-5 and rel(inv, g["saveinvvar"].ravel()) < 2e-5
	# the running variance IS the unbiased estimate (cuDNN) -- pinned here by the reference's own backend
	assert rel(rmean, g["runmean"].ravel()) < 2e-5 and rel(rvar, g["runvar"].ravel()) < 2e-5

	dx, dscale, dbias = ops.batchnorm_bwd(x, dy, g["in_scale"].ravel(), mu, inv)
	assert rel(dx, g["dx"]) < t
	assert rel(dscale, g["dscale"].ravel()) < max(t, 1e-4) and rel(dbias, g["dbias"].ravel()) < max(t, 1e-4)
	assert rel(ops.batchnorm_infer(x, g["in_scale"].ravel(), g["in_bias"].ravel(), rmean, rvar), g["yinfer"]) < t


POOLS = {
	"maxpool3s2_ties_f32": (3, 2, 0, "max"), "maxpool2s2_f32": (2, 2, 0, "max"), "maxpool3s2p1_ties_f32": (3, 2, 1, "max"),
	"maxpool3s1p1_ties_f32": (3, 1, 1, "max"), "avgpadpool_f32": (3, 2, 1, "avgWithPad"), "avgnopadpool_f32": (3, 2, 1, "avgNoPad"),
	"avgpool7_f32": (7, 1, 0, "avgWithPad"), "maxpool2s2_f16": (2, 2, 0, "max"),
}


@pytest.mark.parametrize("name", sorted(POOLS))
def test_pool2d_forward_backward_with_ties(gold, name):
	g = case(gold, name)
	size, stride, pad, mode = POOLS[name]
	x, dy = g["in_x"], g["in_dy"]
	y = ops.pool2d(x, size, stride, pad, mode)
	assert rel(y, g["y"]) < (1e-3 if name.endswith("f16") else 1e-6)
	# ties: the gradient of a window goes to the FIRST maximum in row-major order -- what cuDNN does on this stack
	dx = ops.pool2d_bwd(x, g["y"], dy, size, stride, pad, mode)
	assert rel(dx, g["dx"]) < (2e-3 if name.endswith("f16") else 1e-6)


def test_pool3d_max_with_ties(gold):
	g = case(gold, "maxpool3d_ties_f32")
	y = ops.pool3d(g["in_x"], 2, 2, 0, "max")
	y = y[0] if isinstance(y, tuple) else y
	assert np.array_equal(np.asarray(y, np.float32), g["y"])
	assert rel(ops.pool3d_bwd(g["in_x"], g["in_dy"], 2, 2, 0, "max"), g["dx"]) < 1e-6


@pytest.mark.parametrize("name,geo", [("maxpoolmask_f32", (3, 2, 1)), ("maxpoolmask_ties_f32", (3, 2, 0))])
def test_maxpool_mask_is_bit_exact_with_the_reference_kernels(gold, name, geo):
	g = case(gold, name)
	size, stride, pad = geo
	x = g["in_x"]
	y, mask = ops.maxpool2d_mask(x, size, stride, pad)
	assert np.array_equal(y, g["y"]) and np.array_equal(mask, g["mask"])
	assert np.array_equal(ops.maxpool2d_mask_bwd(g["in_dy"], x.shape, mask, size, stride, pad), g["dx"])
	assert np.array_equal(ops.maxunpool2d(g["y"], x.shape, mask), g["unpool"])
	assert np.array_equal(ops.maxunpool2d_bwd(x, g["y"].shape, mask), g["unpoolgrad"])


@pytest.mark.parametrize("name", ["softmax_flat_f32", "softmax_spatial_f32", "softmax_flat_f16"])
def test_softmax(gold, name):
	g = case(gold, name)
	t = 2e-3 if name.endswith("f16") else 1e-6
	assert rel(ops.softmax(g["in_x"]), g["y"]) < t
	assert rel(ops.softmax_bwd(g["y"], g["in_dy"]), g["dx"]) < t


@pytest.mark.parametrize("name,across,N", [("lrn_cross_f32", True, 5), ("lrn_map_f32", False, 3)])
def test_lrn(gold, name, across, N):
	g = case(gold, name)
	y, dx = ops.lrn(g["in_x"], N, 1e-2, 0.75, 2.0, across, grad=g["in_dy"])
	assert rel(y, g["y"]) < 1e-5 and rel(dx, g["dx"]) < 1e-5


@pytest.mark.parametrize("name", ["gemm_f32", "gemm_f16"])
def test_gemm(gold, name):
	g = case(gold, name)
	A, B, C0 = g["in_A"], g["in_B"], g["in_C0"]
	t = 2e-3 if name.endswith("f16") else 2e-5
	assert rel(ops.gemm(A, B), g["nn"]) < t and rel(ops.gemm(A, B), g["tn"]) < t and rel(ops.gemm(A, B), g["nt"]) < t
	assert rel(ops.gemm(A, B, out=C0, alpha=0.5, beta=0.9), g["acc"]) < t
	assert rel(A.astype(np.float64).sum(0), g["colsum"]) < t and rel(A.astype(np.float64).sum(1), g["rowsum"]) < t


def test_matvec_helpers(gold):
	g = case(gold, "matvec_f32")
	A = g["in_A"]
	assert np.array_equal(ops.add_vec_to_mat(g["in_u"], A, 1).astype(np.float32), g["addrow"])
	assert np.array_equal(ops.add_vec_to_mat(g["in_v"], A, 0).astype(np.float32), g["addcol"])
	assert np.array_equal((A + np.tile(g["in_w"], 4)[None, :]).astype(np.float32), g["addtile"])
	assert np.array_equal(ops.argmax(A, 1), g["argmax1"]) and np.array_equal(ops.argmax(A, 0), g["argmax0"])
	assert np.array_equal(ops.argmax(g["in_ties"], 1), g["argmaxties"])      # ties: the reference's butterfly order, not the first


@pytest.mark.parametrize("name", ["act_f32", "act_f16"])
def test_activations(gold, name):
	g = case(gold, name)
	x, dy = g["in_x"], g["in_dy"]
	# float32: the reference compiles its kernels with -use_fast_math (tanh.approx / ex2.approx: ~1e-5 absolute)
	t = 4e-3 if name.endswith("f16") else 2e-5
	for kind, args in gc.ACTIVATIONS.items():
		y = g[kind]
		want = ops.activation(kind, x, *args)
		assert np.abs(want - y).max() < t * max(1.0, np.abs(want).max()), kind
		ref = x if kind == "gelu" else y
		wantg = ops.activation_bwd(kind, dy, ref, *args)
		assert np.abs(wantg - g[kind + "Der"]).max() < t * max(1.0, np.abs(wantg).max()), kind


@pytest.mark.parametrize("name", ["optim_f32", "optim_f16"])
def test_optimizer_kernels(gold, name):
	g = case(gold, name)
	p, gr, m = g["in_p"], g["in_g"], g["in_m"]
	t = 2e-3 if name.endswith("f16") else 1e-6
	wp, wm = ops.sgd_momentum(p, gr, m, 0.01, 0.9)
	assert rel(wp, g["momsgd_p"]) < t and rel(wm, g["momsgd_m"]) < t
	wp, wm = ops.nesterov_update(p, gr, m, 0.01, 0.9)
	assert rel(wp, g["nesterov_p"]) < t and rel(wm, g["nesterov_m"]) < t
	wp, wg, ws = ops.adam_update(p, gr, g["in_mg"], g["in_ms"], 1e-3, 0.1, 0.001, 1e-8)
	assert rel(wp, g["adam_p"]) < t and rel(wg, g["adam_mg"]) < 1e-6 and rel(ws, g["adam_ms"]) < 1e-6


@pytest.mark.parametrize("name", ["xent_flat", "xent_spatial"])
def test_cross_entropy(gold, name):
	g = case(gold, name)
	err, grad = ops.cross_entropy(g["in_scores"], g["in_labels"])
	assert abs(err - float(g["error"])) < 1e-4 * abs(err) and rel(grad, g["grad"]) < 1e-5
	err, grad = ops.cross_entropy(g["in_scores"], g["in_labels"], g["in_weights"])
	assert abs(err - float(g["werror"])) < 1e-4 * abs(err) and rel(grad, g["wgrad"]) < 1e-5


def test_instance_norm(gold):
	g = case(gold, "instnorm_f32")
	x, dy = g["in_x"], g["in_dy"]
	N, C, H, W = x.shape
	scale, bias = np.tile(g["in_scale"].ravel(), N), np.tile(g["in_bias"].ravel(), N)
	xr = x.reshape(1, N * C, H, W)
	y, mu, inv, _, _ = ops.batchnorm_train(xr, scale, bias, np.zeros(N * C), np.ones(N * C))
	assert rel(y.reshape(x.shape), g["y"]) < 2e-5
	dx, dscale, dbias = ops.batchnorm_bwd(xr, dy.reshape(xr.shape), scale, mu, inv)
	assert rel(dx.reshape(x.shape), g["dx"]) < 2e-5
	assert rel(dscale.reshape(N, C).sum(0), g["dscale"].ravel()) < 1e-4 and rel(dbias.reshape(N, C).sum(0), g["dbias"].ravel()) < 1e-4


def test_memory_ops(gold):
	g = case(gold, "memory_f32")
	x = g["in_x"]
	assert np.array_equal(np.transpose(x, (2, 0, 3, 1)), g["transpose"])
	assert np.array_equal(np.moveaxis(x, 1, 3), g["moveaxis"]) and np.array_equal(np.swapaxes(x, 0, 2), g["swapaxes"])
	a, b = g["in_a"], g["in_b"]
	cat = np.zeros((2, 5, 6, 6), np.float32)
	cat[:, :3] = a
	cat[:, 3:, 1:5, 1:5] = b
	assert np.array_equal(cat, g["depthconcat"])
	assert np.array_equal(g["in_g"][:, :3], g["splita"]) and np.array_equal(g["in_g"][:, 3:, 1:5, 1:5], g["splitb"])
