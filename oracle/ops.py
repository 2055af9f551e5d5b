"""CPU oracle of the hot path -- numpy restatement of what the reference's GPU backend computes, op by op.

TEST INFRASTRUCTURE ONLY.  Nothing under puzzlelib_b200/ imports this package; only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` legs do, and there only as the checker or the timed CPU baseline.

Where the arithmetic lives.  In the reference these ops are calls into closed third-party libraries that are NOT in
/root/reference and are not version-pinned by it (no version file; requirements.txt lists Python deps only):
cuDNN (conv / pool / batch-norm / softmax; this image: 9.10.2) and cuBLAS (GEMM; this image: 12.9), plus the
reference's own NVRTC-JIT kernels (activations, axpy-style updates, max-pool with mask, mat-vec helpers) whose source
IS in the tree.  Each function below therefore follows either
  * the JIT kernel source itself (cited file:line), or
  * the host recomputation the reference's own unit test holds for the library call (cited file:line) -- those test
    loops are the reference's specification of the cuDNN / cuBLAS result.

Pinning.  (a) tests/golden/ref_cpu_*.npz hold outputs of the REFERENCE ITSELF (its numpy CPU backend, imported from
/root/reference by tools/gen_golden.py in the build container) for conv2d / pool2d / batch-norm-inference forward,
Linear forward+backward and every activation forward+backward; tests/test_oracle.py checks this oracle against them.
(b) tests/golden/ref_cuda_ops.npz holds outputs of the reference's OWN CUDA backend (cuDNN 9.10 / cuBLAS 12.9 / NVRTC),
built from the unmodified tree by baseline/build_ref.py and run on a B200 by tools/gen_golden_cuda.py: forward AND backward
of conv (dgrad / wgrad / bgrad incl. scale / momentum accumulation, groups, dilation, 3-d), deconv, batch-norm train /
backward / inference (incl. the running variance: the UNBIASED estimate), pooling backward on tied maxima (the first maximum
of a window in row-major order takes the gradient), softmax, LRN, GEMM, cross-entropy, the optimizer kernels, argmax (whose
tie order along the last axis is the reference kernel's butterfly order, not the first occurrence).
tests/test_oracle_cuda_golden.py checks every function of this oracle against them: PARITY PINNED for forward and backward.
(c) Not pinned by reference output: recurrent layers (the reference's cuDNN-7 RNN API does not exist in cuDNN 9, so its RNN
cannot run on this stack; the oracle follows the host formulas of Cuda/Wrappers/CuDnnRnn.py:178-300) and TF32 rounding
(cuDNN computes float32 convolutions in full fp32 on this stack; this backend's tcgen05 path uses TF32 products and is held
to the 1e-3 relative bar of BASELINE.json's north_star).

All functions take and return numpy arrays; `dtype` selects the accumulation type (float64 for parity checks,
float32 for the timed CPU baseline).
"""
import numpy as np
from numpy.lib.stride_tricks import as_strided

FLT_MAX = np.finfo(np.float32).max


def _pair(v):
	return (int(v), int(v)) if isinstance(v, (int, np.integer)) else (int(v[0]), int(v[1]))


# ================================================================================================ convolution
def conv_out_size(insize, fsize, stride, pad, dilation):
	"""reference: CuDnn.c:242-266"""
	return (insize + 2 * pad - dilation * (fsize - 1) - 1) // stride + 1


def _im2col(xp, R, S, P, Q, stride, dilation):
	"""(N, C, Hp, Wp) padded input -> view (N, C, R, S, P, Q) without copying"""
	sN, sC, sH, sW = xp.strides
	N, C = xp.shape[:2]
	return as_strided(
		xp, shape=(N, C, R, S, P, Q),
		strides=(sN, sC, dilation[0] * sH, dilation[1] * sW, stride[0] * sH, stride[1] * sW), writeable=False
	)


def _padded(x, pad, dtype):
	N, C, H, W = x.shape
	if pad == (0, 0) and x.dtype == dtype:
		return x
	xp = np.zeros((N, C, H + 2 * pad[0], W + 2 * pad[1]), dtype=dtype)
	xp[:, :, pad[0]:pad[0] + H, pad[1]:pad[1] + W] = x
	return xp


def conv2d(x, w, bias=None, stride=1, pad=0, dilation=1, groups=1, dtype=np.float64):
	"""y[n,k,p,q] = sum_{c,r,s} x[n, g*Cg+c, p*sh-ph+r*dh, q*sw-pw+s*dw] * w[k,c,r,s] (+ b[k]) -- cross-correlation.
	reference: cudnnConvolutionForward call site CuDnn.c:397-449; host loops of the unit tests
	Cuda/Wrappers/CuDnn.py:29-80 (conv2dTest), :137-201 (convGroupTest); Modules/Conv2D.py:92-327."""
	stride, pad, dilation = _pair(stride), _pair(pad), _pair(dilation)
	N, C, H, W = x.shape
	K, Cg, R, S = w.shape
	assert C == Cg * groups and K % groups == 0
	Kg = K // groups
	P, Q = conv_out_size(H, R, stride[0], pad[0], dilation[0]), conv_out_size(W, S, stride[1], pad[1], dilation[1])

	xp = _padded(x, pad, dtype)
	y = np.empty((N, K, P, Q), dtype=dtype)
	wm = w.astype(dtype, copy=False)

	for g in range(groups):
		cols = _im2col(xp[:, g * Cg:(g + 1) * Cg], R, S, P, Q, stride, dilation).reshape(N, Cg * R * S, P * Q)
		y[:, g * Kg:(g + 1) * Kg] = np.matmul(wm[g * Kg:(g + 1) * Kg].reshape(Kg, Cg * R * S), cols).reshape(N, Kg, P, Q)

	if bias is not None:
		y += bias.astype(dtype).reshape(1, K, 1, 1)
	return y


def conv2d_bwd_data(dy, w, inshape=None, bias=None, stride=1, pad=0, dilation=1, postpad=0, groups=1, dtype=np.float64):
	"""dx = transposed convolution of dy with w; also IS the Deconv forward (+bias over the produced maps).
	reference: cudnnConvolutionBackwardData call site CuDnn.c:517-571; in-shape with `postpad` :269-284;
	host loops Cuda/Wrappers/CuDnn.py:29-80 (conv2dTest, bwd data), :204-252 (deconv2dTest)."""
	stride, pad, dilation, postpad = _pair(stride), _pair(pad), _pair(dilation), _pair(postpad)
	N, K, P, Q = dy.shape
	Kw, Cg, R, S = w.shape
	assert K == Kw
	C, Kg = Cg * groups, K // groups

	if inshape is None:
		H = (P - 1) * stride[0] + dilation[0] * (R - 1) - 2 * pad[0] + 1 + postpad[0]
		W = (Q - 1) * stride[1] + dilation[1] * (S - 1) - 2 * pad[1] + 1 + postpad[1]
	else:
		H, W = inshape[2], inshape[3]

	Hp, Wp = H + 2 * pad[0], W + 2 * pad[1]
	dxp = np.zeros((N, C, Hp, Wp), dtype=dtype)
	wm = w.astype(dtype, copy=False)
	dym = dy.astype(dtype, copy=False).reshape(N, K, P * Q)

	for g in range(groups):
		wg = wm[g * Kg:(g + 1) * Kg].reshape(Kg, Cg * R * S)
		dcols = np.matmul(wg.T, dym[:, g * Kg:(g + 1) * Kg]).reshape(N, Cg, R, S, P, Q)
		for r in range(R):
			for s in range(S):
				h0, w0 = r * dilation[0], s * dilation[1]
				dxp[:, g * Cg:(g + 1) * Cg, h0:h0 + (P - 1) * stride[0] + 1:stride[0], w0:w0 + (Q - 1) * stride[1] + 1:stride[1]] += \
					dcols[:, :, r, s]

	dx = dxp[:, :, pad[0]:pad[0] + H, pad[1]:pad[1] + W]
	if bias is not None:
		dx = dx + bias.astype(dtype).reshape(1, C, 1, 1)
	return np.ascontiguousarray(dx)


def conv2d_bwd_params(x, dy, wshape, stride=1, pad=0, dilation=1, groups=1, withbias=False, deconv=False, wgrad=None,
					  bgrad=None, scale=1.0, momentum=0.0, dtype=np.float64):
	"""dW = scale * sum_{n,p,q} x (*) dy + momentum * dW_old;  db = scale * sum dy + momentum * db_old, where the
	bias side is `dy` for a conv and `x` for a deconv.
	reference: cudnnConvolutionBackwardFilter alpha/beta CuDnn.c:682-685, BackwardBias :375-394, deconv side :689-690,774;
	host loops Cuda/Wrappers/CuDnn.py:29-80 (conv2dTest, wgrad / bgrad), :204-252 (deconv side)."""
	stride, pad, dilation = _pair(stride), _pair(pad), _pair(dilation)
	N, C, H, W = x.shape
	K, Cg, R, S = wshape
	Kg = K // groups
	P, Q = dy.shape[2:]

	xp = _padded(x, pad, dtype)
	dym = dy.astype(dtype, copy=False).reshape(N, K, P * Q)
	dw = np.empty(wshape, dtype=dtype)

	for g in range(groups):
		cols = _im2col(xp[:, g * Cg:(g + 1) * Cg], R, S, P, Q, stride, dilation).reshape(N, Cg * R * S, P * Q)
		acc = np.matmul(dym[:, g * Kg:(g + 1) * Kg], cols.transpose(0, 2, 1)).sum(axis=0)
		dw[g * Kg:(g + 1) * Kg] = acc.reshape(Kg, Cg, R, S)

	dw = scale * dw + (momentum * wgrad.astype(dtype) if wgrad is not None and momentum != 0.0 else 0.0)
	if not withbias:
		return dw

	side = x if deconv else dy
	db = scale * side.astype(dtype).sum(axis=(0, 2, 3))
	if bgrad is not None and momentum != 0.0:
		db = db + momentum * bgrad.astype(dtype).ravel()
	return dw, db


# ================================================================================================ GEMM / mat-vec
def gemm(A, B, out=None, transpA=False, transpB=False, alpha=1.0, beta=0.0, dtype=np.float64):
	"""reference: cublasGemmEx call site CuBlas.c:327-403; unit test Cuda/Wrappers/CuBlas.py:32-47 (matrixTest)"""
	a = A.astype(dtype, copy=False)
	b = B.astype(dtype, copy=False)
	r = alpha * np.matmul(a.T if transpA else a, b.T if transpB else b)
	if out is not None and beta != 0.0:
		r = r + beta * out.astype(dtype)
	return r


def add_vec_to_mat(vec, mat, axis=1):
	"""reference: Cuda/Kernels/MatVec.py:128-171 (opRowVecToMat / opRowOneVecToMat / opColVecToMat)"""
	if axis == 1:
		if vec.shape[-1] == mat.shape[-1]:
			return mat + vec[..., None, :]
		reps = mat.shape[-1] // vec.shape[-1]
		return mat + np.tile(vec, reps)[..., None, :]
	return mat + vec[..., :, None]


def matsum(tensor, axis, out=None, alpha=1.0, beta=0.0, dtype=np.float64):
	"""reference: Cuda/Kernels/MatVec.py:60-91 (out = beta * out + alpha * sum)"""
	r = alpha * tensor.astype(dtype).sum(axis=axis)
	if out is not None:
		r = r + beta * out.astype(dtype)
	return r


def argmax(tensor, axis):
	"""reference: Cuda/Kernels/MatVec.py:8-57.  Along any axis but the last the kernel scans sequentially with a strict
	comparison: the first occurrence wins.  Along the LAST axis (minMaxOnRow) 32 lanes scan i = lane, lane + 32, ... (strict),
	merge in an xor butterfly (strict: a lane keeps its own survivor on a tie) and lane 0's survivor is what lands in memory
	-- pinned by the reference's own CUDA backend (tests/golden/ref_cuda_ops.npz, matvec_f32/argmaxties)."""
	tensor = np.asarray(tensor)
	if axis % tensor.ndim != tensor.ndim - 1:
		return np.argmax(tensor, axis=axis).astype(np.int32)

	w = tensor.shape[-1]
	rows = tensor.reshape(-1, w).astype(np.float64)
	padded = np.full((rows.shape[0], -(-w // 32) * 32), -np.inf)
	padded[:, :w] = rows
	lanes = padded.reshape(rows.shape[0], -1, 32)                       # [row][round][lane]
	best = lanes.max(axis=1)
	idx = (lanes.argmax(axis=1) * 32 + np.arange(32)).astype(np.int64)     # first (strict) occurrence inside every lane
	idx[np.isneginf(best)] = -1
	mask = 16
	while mask:
		partner = np.arange(32) ^ mask
		take = best[:, partner] > best
		best, idx = np.where(take, best[:, partner], best), np.where(take, idx[:, partner], idx)
		mask //= 2
	return idx[:, 0].reshape(tensor.shape[:-1]).astype(np.int32)


# ================================================================================================ batch norm
def batchnorm_train(x, scale, bias, mean, var, epsilon=1e-5, factor=1.0, dtype=np.float64):
	"""Spatial batch norm, training.  Returns (y, savemean, saveinvvar, new_running_mean, new_running_var).
	reference: cudnnBatchNormalizationForwardTraining call site CuDnnNorm.c:31-71; unit test formulas
	Cuda/Wrappers/CuDnnNorm.py:36-48 (y, saveinvvar = 1/sqrt(biased var + eps), running mean at factor 1).
	The running VARIANCE follows cuDNN's documented behaviour (unbiased estimate) -- not asserted by the reference."""
	axes = (0, ) + tuple(range(2, x.ndim))
	shape = (1, -1) + (1, ) * (x.ndim - 2)
	xd = x.astype(dtype)
	m = xd.size // xd.shape[1]

	mu = xd.mean(axis=axes)
	varb = xd.var(axis=axes)
	invstd = 1.0 / np.sqrt(varb + epsilon)

	y = (xd - mu.reshape(shape)) * invstd.reshape(shape) * scale.astype(dtype).reshape(shape) + bias.astype(dtype).reshape(shape)

	varu = varb * m / (m - 1) if m > 1 else varb
	newmean = (1.0 - factor) * mean.astype(dtype).ravel() + factor * mu
	newvar = (1.0 - factor) * var.astype(dtype).ravel() + factor * varu
	return y, mu, invstd, newmean, newvar


def batchnorm_infer(x, scale, bias, mean, var, epsilon=1e-5, dtype=np.float64):
	"""reference: cudnnBatchNormalizationForwardInference CuDnnNorm.c:55; test Cuda/Wrappers/CuDnnNorm.py:70-77;
	CPU backend CPU/Wrappers/NumpyDnn.py:115-129"""
	shape = (1, -1) + (1, ) * (x.ndim - 2)
	a = scale.astype(dtype).ravel() / np.sqrt(var.astype(dtype).ravel() + epsilon)
	b = bias.astype(dtype).ravel() - mean.astype(dtype).ravel() * a
	return x.astype(dtype) * a.reshape(shape) + b.reshape(shape)


def batchnorm_bwd(x, dy, scale, savemean, saveinvvar, dtype=np.float64):
	"""dbias = sum dy; dscale = sum dy * xhat; dx = scale * invstd * (dy - dbias/m - xhat * dscale/m)
	reference: cudnnBatchNormalizationBackward CuDnnNorm.c:158-194; test formulas Cuda/Wrappers/CuDnnNorm.py:50-68"""
	axes = (0, ) + tuple(range(2, x.ndim))
	shape = (1, -1) + (1, ) * (x.ndim - 2)
	xd, g = x.astype(dtype), dy.astype(dtype)
	m = xd.size // xd.shape[1]

	invstd = saveinvvar.astype(dtype).reshape(shape)
	xhat = (xd - savemean.astype(dtype).reshape(shape)) * invstd
	dbias = g.sum(axis=axes)
	dscale = (g * xhat).sum(axis=axes)
	dx = scale.astype(dtype).reshape(shape) * invstd * (g - dbias.reshape(shape) / m - xhat * dscale.reshape(shape) / m)
	return dx, dscale, dbias


# ================================================================================================ pooling
def pool_out_size(insize, fsize, stride, pad):
	"""reference: CuDnnPool.c:24-41; Cuda/Kernels/Pool.py:131-132"""
	return (insize + 2 * pad - fsize) // stride + 1


def _windows(H, W, size, stride, pad):
	OH, OW = pool_out_size(H, size[0], stride[0], pad[0]), pool_out_size(W, size[1], stride[1], pad[1])
	for oh in range(OH):
		h0 = oh * stride[0] - pad[0]
		h1 = min(h0 + size[0], H)
		h0 = max(h0, 0)
		for ow in range(OW):
			w0 = ow * stride[1] - pad[1]
			w1 = min(w0 + size[1], W)
			w0 = max(w0, 0)
			yield oh, ow, h0, h1, w0, w1


def maxpool2d_mask(x, size=2, stride=2, pad=0):
	"""Max pool returning (y, int32 mask): mask = h*W + w of the FIRST strict maximum of a row-major scan of the
	window clipped to the plane, starting from -FLT_MAX; -1 for an empty window.
	reference: kernel Cuda/Kernels/Pool.py:10-46; unit test :229-263 (exact mask equality)."""
	size, stride, pad = _pair(size), _pair(stride), _pair(pad)
	N, C, H, W = x.shape
	OH, OW = pool_out_size(H, size[0], stride[0], pad[0]), pool_out_size(W, size[1], stride[1], pad[1])
	y = np.full((N, C, OH, OW), -FLT_MAX, dtype=np.float32)
	mask = np.full((N, C, OH, OW), -1, dtype=np.int32)

	for oh, ow, h0, h1, w0, w1 in _windows(H, W, size, stride, pad):
		if h1 <= h0 or w1 <= w0:
			continue
		win = x[:, :, h0:h1, w0:w1].reshape(N, C, -1)
		valid = win > -FLT_MAX                      # NaN and -FLT_MAX never win the strict '>' test
		cand = np.where(valid, win, -np.inf)
		idx = np.argmax(cand, axis=2)               # first occurrence of the maximum
		has = valid.any(axis=2)
		hh, ww = h0 + idx // (w1 - w0), w0 + idx % (w1 - w0)
		y[:, :, oh, ow] = np.where(has, np.take_along_axis(win, idx[..., None], axis=2)[..., 0], -FLT_MAX)
		mask[:, :, oh, ow] = np.where(has, hh * W + ww, -1)
	return y, mask


def maxpool2d_mask_bwd(dy, inshape, mask, size=2, stride=2, pad=0):
	"""dx[h,w] = sum over the candidate windows, in (ph, pw) order, of dy where mask == h*W + w (fp32 adds).
	reference: kernel Cuda/Kernels/Pool.py:66-96; unit test :265-281."""
	size, stride, pad = _pair(size), _pair(stride), _pair(pad)
	N, C, H, W = inshape
	dx = np.zeros((N, C, H * W), dtype=np.float32)
	OH, OW = dy.shape[2:]
	n_idx, c_idx = np.meshgrid(np.arange(N), np.arange(C), indexing="ij")

	for oh in range(OH):                            # (ph, pw) ascending = the kernel's accumulation order per element
		for ow in range(OW):
			m = mask[:, :, oh, ow]
			ok = m >= 0
			np.add.at(dx, (n_idx[ok], c_idx[ok], m[ok]), dy[:, :, oh, ow][ok].astype(np.float32))
	return dx.reshape(N, C, H, W)


def maxunpool2d(x, outshape, mask):
	"""reference: kernel Cuda/Kernels/Pool.py:48-63"""
	N, C = x.shape[:2]
	y = np.zeros((N, C, outshape[2] * outshape[3]), dtype=x.dtype)
	np.put_along_axis(y, mask.reshape(N, C, -1).astype(np.int64), x.reshape(N, C, -1), axis=2)
	return y.reshape(N, C, outshape[2], outshape[3])


def maxunpool2d_bwd(dy, poolshape, mask):
	"""reference: kernel Cuda/Kernels/Pool.py:98-112"""
	N, C = dy.shape[:2]
	dx = np.take_along_axis(dy.reshape(N, C, -1), mask.reshape(N, C, -1).astype(np.int64), axis=2)
	return dx.reshape(N, C, poolshape[2], poolshape[3])


def pool2d(x, size=2, stride=2, pad=0, mode="max", dtype=np.float64):
	"""cuDNN-style pooling forward.  mode: max | avgWithPad (divide by fh*fw) | avgNoPad (divide by the valid count).
	reference: cudnnPoolingForward call site CuDnnPool.c:64-96 (NOT_PROPAGATE_NAN); host loops of the unit test
	Cuda/Wrappers/CuDnn.py:376-393 (maxpool2dTest); CPU backend CPU/Wrappers/NumpyDnn.py:83-112 (max, avg-with-pad)."""
	size, stride, pad = _pair(size), _pair(stride), _pair(pad)
	N, C, H, W = x.shape
	OH, OW = pool_out_size(H, size[0], stride[0], pad[0]), pool_out_size(W, size[1], stride[1], pad[1])
	xd = x.astype(dtype)
	y = np.empty((N, C, OH, OW), dtype=dtype)

	for oh, ow, h0, h1, w0, w1 in _windows(H, W, size, stride, pad):
		win = xd[:, :, h0:h1, w0:w1]
		if mode == "max":
			y[:, :, oh, ow] = np.fmax.reduce(win.reshape(N, C, -1), axis=2)
		else:
			cnt = size[0] * size[1] if mode == "avgWithPad" else (h1 - h0) * (w1 - w0)
			y[:, :, oh, ow] = win.sum(axis=(2, 3)) / cnt
	return y


def pool2d_bwd(x, y, dy, size=2, stride=2, pad=0, mode="max", dtype=np.float64):
	"""cuDNN-style pooling backward.  max: the window's dy goes to the FIRST element equal to the window maximum
	(row-major); with no ties this is the reference test's expectation (Cuda/Wrappers/CuDnn.py:395-410).  Tie routing
	itself is not asserted by any reference test -- "parity unpinned" for ties.
	reference: cudnnPoolingBackward call site CuDnnPool.c:155-190."""
	size, stride, pad = _pair(size), _pair(stride), _pair(pad)
	N, C, H, W = x.shape
	dx = np.zeros((N, C, H, W), dtype=dtype)
	g = dy.astype(dtype)
	n_idx, c_idx = np.meshgrid(np.arange(N), np.arange(C), indexing="ij")

	for oh, ow, h0, h1, w0, w1 in _windows(H, W, size, stride, pad):
		if mode == "max":
			win = x[:, :, h0:h1, w0:w1].reshape(N, C, -1)
			idx = np.argmax(win == y[:, :, oh, ow][..., None].astype(x.dtype), axis=2)
			hh, ww = h0 + idx // (w1 - w0), w0 + idx % (w1 - w0)
			np.add.at(dx, (n_idx, c_idx, hh, ww), g[:, :, oh, ow])
		else:
			cnt = size[0] * size[1] if mode == "avgWithPad" else (h1 - h0) * (w1 - w0)
			dx[:, :, h0:h1, w0:w1] += (g[:, :, oh, ow] / cnt)[..., None, None]
	return dx


# ================================================================================================ softmax
def softmax(x, mode="spatial", dtype=np.float64):
	"""SOFTMAX_ACCURATE: exp(x - max) / sum.  spatial = over axis 1 for every (n, spatial...) position
	(CUDNN_SOFTMAX_MODE_CHANNEL); perActivation = over all non-batch axes (MODE_INSTANCE).
	reference: cudnnSoftmaxForward call site CuDnn.c:974-997; unit test Cuda/Wrappers/CuDnn.py:454-470 (softmax2dTest)"""
	xd = x.astype(dtype)
	axes = 1 if mode == "spatial" else tuple(range(1, x.ndim))
	e = np.exp(xd - xd.max(axis=axes, keepdims=True))
	return e / e.sum(axis=axes, keepdims=True)


def softmax_bwd(y, dy, mode="spatial", dtype=np.float64):
	"""dx = y * (dy - sum(y * dy)).  reference: cudnnSoftmaxBackward CuDnn.c:1053-1079; test Cuda/Wrappers/CuDnn.py:472-485"""
	yd, g = y.astype(dtype), dy.astype(dtype)
	axes = 1 if mode == "spatial" else tuple(range(1, y.ndim))
	return yd * (g - (yd * g).sum(axis=axes, keepdims=True))


# ================================================================================================ elementwise
def _erf(x):
	from math import erf
	return np.vectorize(erf, otypes=[np.float64])(x)


def activation(kind, x, a=None, b=None, dtype=np.float64):
	"""reference formulas: Cuda/Kernels/ElementWise.py:18 (sigmoid), :73 (tanh), :128 (relu), :184-188 (leakyRelu),
	:249-254 (elu), :313 (softPlus), :369-373 (clip), :436-441 (gelu); same in CPU/Kernels/ElementWise.py"""
	x = x.astype(dtype)
	if kind == "sigmoid":
		return 1.0 / (1.0 + np.exp(-x))
	if kind == "tanh":
		return np.tanh(x)
	if kind == "relu":
		return x * (x > 0)
	if kind == "leakyRelu":
		a = 0.01 if a is None else a
		return x * ((x > 0) + a * (x <= 0))
	if kind == "elu":
		a = 1.0 if a is None else a
		return x * (x > 0) + a * (np.exp(x) - 1.0) * (x <= 0)
	if kind == "softPlus":
		return np.log(1.0 + np.exp(x))
	if kind == "clip":
		a, b = (0.0 if a is None else a), (6.0 if b is None else b)
		return np.minimum(b, np.maximum(a, x))
	if kind == "gelu":
		return 0.5 * x * (1.0 + _erf(x / np.sqrt(2.0)))
	raise ValueError(kind)


def activation_bwd(kind, g, ref, a=None, b=None, dtype=np.float64):
	"""`ref` is the activation OUTPUT (the INPUT for gelu, whose Gaussian term uses 1/sqrt(pi), sic).
	reference: Cuda/Kernels/ElementWise.py:45, :100, :156, :215-220, :281-286, :341, :403-407, :468-477"""
	g, d = g.astype(dtype), ref.astype(dtype)
	if kind == "sigmoid":
		return g * d * (1.0 - d)
	if kind == "tanh":
		return g * (1.0 - d * d)
	if kind == "relu":
		return g * (d > 0)
	if kind == "leakyRelu":
		a = 0.01 if a is None else a
		return g * ((d > 0) + a * (d <= 0))
	if kind == "elu":
		a = 1.0 if a is None else a
		return g * ((d > 0) + (d + a) * (d <= 0))
	if kind == "softPlus":
		return g * (1.0 - np.exp(-d))
	if kind == "clip":
		a, b = (0.0 if a is None else a), (6.0 if b is None else b)
		return g * ((d > a) & (d < b))
	if kind == "gelu":
		return g * (0.5 * (1.0 + _erf(d / np.sqrt(2.0))) + d / np.sqrt(np.pi) * np.exp(-0.5 * d * d))
	raise ValueError(kind)


def sgd_momentum(param, grad, mom, learnRate, momRate, dtype=np.float64):
	"""mom = momRate * mom + learnRate * grad; param += mom (gradients are ascent direction).
	reference: Cuda/Kernels/ElementWise.py:771-800; Optimizers/MomentumSGD.py:24-27"""
	mom = momRate * mom.astype(dtype) + learnRate * grad.astype(dtype)
	return param.astype(dtype) + mom, mom


def grid_mean(tensors, dtype=np.float64):
	"""Grid.sumTensor: the mean of the per-rank buffers (reference: Grid.py:123-135, beta = 1/P)"""
	acc = np.zeros_like(tensors[0], dtype=dtype)
	for t in tensors:
		acc += t.astype(dtype)
	return acc / len(tensors)


# ================================================================================================ recurrent layers
def _sigmoid(x):
	return 1.0 / (1.0 + np.exp(-x))


def lstm_forward(x, params, h0=None, c0=None, dtype=np.float64, reverse=False):
	"""One uni-directional LSTM layer.  Follows the reference's own host loop (Cuda/Wrappers/CuDnnRnn.py:178-236): gates
	i, f, o sigmoid, candidate c tanh; c_t = f*c_{t-1} + i*g; h_t = o*tanh(c_t); double bias bw* + br*.
	x (T, B, in); params: dict wi wf wc wo (H, in), ri rf rc ro (H, H), bw*/br* (H,).  Returns (out (T,B,H), cache)."""
	x = np.asarray(x, dtype)
	T, B, _ = x.shape
	H = params["ri"].shape[0]
	p = {k: np.asarray(v, dtype) for k, v in params.items()}
	h = np.zeros((B, H), dtype) if h0 is None else np.asarray(h0, dtype)
	c = np.zeros((B, H), dtype) if c0 is None else np.asarray(c0, dtype)
	out = np.empty((T, B, H), dtype)
	cache = {k: [None] * T for k in ("i", "f", "g", "o", "c", "hprev", "cprev")}
	cache["reverse"] = reverse
	for t in (range(T - 1, -1, -1) if reverse else range(T)):
		pre = {g: x[t] @ p["w" + g].T + h @ p["r" + g].T + p["bw" + g] + p["br" + g] for g in "ifco"}
		i, f, o, g = _sigmoid(pre["i"]), _sigmoid(pre["f"]), _sigmoid(pre["o"]), np.tanh(pre["c"])
		cache["hprev"][t] = h
		cache["cprev"][t] = c
		c = f * c + i * g
		h = o * np.tanh(c)
		out[t] = h
		for k, v in (("i", i), ("f", f), ("g", g), ("o", o), ("c", c)):
			cache[k][t] = v
	return out, cache


def lstm_backward(x, params, cache, dy, dtype=np.float64):
	"""Gradients of lstm_forward (reference host loop: CuDnnRnn.py:238-300): returns (dx, dparams dict)."""
	x, dy = np.asarray(x, dtype), np.asarray(dy, dtype)
	T, B, insz = x.shape
	p = {k: np.asarray(v, dtype) for k, v in params.items()}
	H = p["ri"].shape[0]
	dp = {k: np.zeros_like(v) for k, v in p.items()}
	dx = np.zeros((T, B, insz), dtype)
	dhn, dcn = np.zeros((B, H), dtype), np.zeros((B, H), dtype)
	for t in (range(T) if cache.get("reverse") else range(T - 1, -1, -1)):
		i, f, g, o, c = (cache[k][t] for k in ("i", "f", "g", "o", "c"))
		hprev, cprev = cache["hprev"][t], cache["cprev"][t]
		dh = dy[t] + dhn
		tc = np.tanh(c)
		dc = dh * o * (1.0 - tc * tc) + dcn
		dpre = {"i": dc * g * i * (1 - i), "f": dc * cprev * f * (1 - f), "c": dc * i * (1 - g * g), "o": dh * tc * o * (1 - o)}
		dhn = np.zeros((B, H), dtype)
		for gname, d in dpre.items():
			dp["w" + gname] += d.T @ x[t]
			dp["r" + gname] += d.T @ hprev
			dp["bw" + gname] += d.sum(axis=0)
			dp["br" + gname] += d.sum(axis=0)
			dx[t] += d @ p["w" + gname]
			dhn += d @ p["r" + gname]
		dcn = dc * f
	return dx, dp


def rnn_forward(x, params, mode="tanh", h0=None, dtype=np.float64):
	"""Plain RNN layer h_t = act(x_t wi^T + h_{t-1} ri^T + bwi + bri) (reference host loop: CuDnnRnn.py:28-58, 95-118)."""
	x = np.asarray(x, dtype)
	T, B, _ = x.shape
	p = {k: np.asarray(v, dtype) for k, v in params.items()}
	H = p["ri"].shape[0]
	h = np.zeros((B, H), dtype) if h0 is None else np.asarray(h0, dtype)
	out = np.empty((T, B, H), dtype)
	for t in range(T):
		pre = x[t] @ p["wi"].T + h @ p["ri"].T + p["bwi"] + p["bri"]
		h = np.maximum(pre, 0.0) if mode == "relu" else np.tanh(pre)
		out[t] = h
	return out


def rnn_backward(x, params, out, dy, mode="tanh", h0=None, dtype=np.float64):
	x, out, dy = np.asarray(x, dtype), np.asarray(out, dtype), np.asarray(dy, dtype)
	T, B, insz = x.shape
	p = {k: np.asarray(v, dtype) for k, v in params.items()}
	H = p["ri"].shape[0]
	dp = {k: np.zeros_like(v) for k, v in p.items()}
	dx = np.zeros((T, B, insz), dtype)
	dhn = np.zeros((B, H), dtype)
	for t in range(T - 1, -1, -1):
		d = (dy[t] + dhn) * ((out[t] > 0) if mode == "relu" else (1.0 - out[t] ** 2))
		hprev = out[t - 1] if t > 0 else (np.zeros((B, H), dtype) if h0 is None else np.asarray(h0, dtype))
		dp["wi"] += d.T @ x[t]
		dp["ri"] += d.T @ hprev
		dp["bwi"] += d.sum(axis=0)
		dp["bri"] += d.sum(axis=0)
		dx[t] = d @ p["wi"]
		dhn = d @ p["ri"]
	return dx, dp


def gru_forward(x, params, h0=None, reverse=False, dtype=np.float64):
	"""One GRU layer direction in cuDNN's formulation (reference host loop Cuda/Wrappers/CuDnnRnn.py:303-352):
	r = sigm(wr x + rr h + bwr + brr), i = sigm(wi x + ri h + bwi + bri), h~ = tanh(wh x + bwh + r * (rh h + brh)),
	h' = (1 - i) h~ + i h.  Returns (out (T,B,H), cache)."""
	x = np.asarray(x, dtype)
	T, B, _ = x.shape
	p = {k: np.asarray(v, dtype) for k, v in params.items()}
	H = p["ri"].shape[0]
	h = np.zeros((B, H), dtype) if h0 is None else np.asarray(h0, dtype)
	out = np.empty((T, B, H), dtype)
	cache = {}
	for t in (range(T - 1, -1, -1) if reverse else range(T)):
		r = _sigmoid(x[t] @ p["wr"].T + h @ p["rr"].T + p["bwr"] + p["brr"])
		i = _sigmoid(x[t] @ p["wi"].T + h @ p["ri"].T + p["bwi"] + p["bri"])
		q = h @ p["rh"].T + p["brh"]
		ht = np.tanh(x[t] @ p["wh"].T + p["bwh"] + r * q)
		cache[t] = (r, i, ht, q, h)
		h = (1.0 - i) * ht + i * h
		out[t] = h
	return out, cache


def gru_backward(x, params, cache, dy, reverse=False, dtype=np.float64):
	"""Gradients of gru_forward (reference host loop: CuDnnRnn.py:354-419): returns (dx, dparams dict)."""
	x, dy = np.asarray(x, dtype), np.asarray(dy, dtype)
	T, B, insz = x.shape
	p = {k: np.asarray(v, dtype) for k, v in params.items()}
	H = p["ri"].shape[0]
	dp = {k: np.zeros_like(v) for k, v in p.items()}
	dx = np.zeros((T, B, insz), dtype)
	dhn = np.zeros((B, H), dtype)
	for t in (range(T) if reverse else range(T - 1, -1, -1)):
		r, i, ht, q, hprev = cache[t]
		dh = dy[t] + dhn
		dpre_h = dh * (1.0 - i) * (1.0 - ht * ht)
		dpre_i = dh * (hprev - ht) * i * (1.0 - i)
		dpre_r = dpre_h * q * r * (1.0 - r)
		dq = dpre_h * r
		for gname, dxside, dhside in (("r", dpre_r, dpre_r), ("i", dpre_i, dpre_i), ("h", dpre_h, dq)):
			dp["w" + gname] += dxside.T @ x[t]
			dp["r" + gname] += dhside.T @ hprev
			dp["bw" + gname] += dxside.sum(axis=0)
			dp["br" + gname] += dhside.sum(axis=0)
			dx[t] += dxside @ p["w" + gname]
		dhn = dh * i + dpre_r @ p["rr"] + dpre_i @ p["ri"] + dq @ p["rh"]
	return dx, dp


# ================================================================================================ 3-d convolution / pooling
def _triple(v):
	return (v, v, v) if isinstance(v, (int, np.integer)) else tuple(v)


def _windows3d(xp, fsize, stride, dilation, outshape):
	"""View (N, C, Do, P, Q, T, R, S) of the padded input: element [.., do, p, q, t, r, s] = xp[.., do*sd + t*dd, ...]."""
	N, C = xp.shape[:2]
	sN, sC, sD, sH, sW = xp.strides
	(T, R, S), (sd, sh, sw), (dd, dh, dw) = fsize, stride, dilation
	Do, P, Q = outshape
	return np.lib.stride_tricks.as_strided(
		xp, shape=(N, C, Do, P, Q, T, R, S), strides=(sN, sC, sD * sd, sH * sh, sW * sw, sD * dd, sH * dh, sW * dw), writeable=False)


def conv3d(x, w, bias=None, stride=1, pad=0, dilation=1, groups=1, dtype=np.float64):
	"""y[n,k,d,p,q] = sum_{c,t,r,s} x[n, g*Cg+c, d*sd-pd+t*dd, ...] * w[k,c,t,r,s] (+ b[k]); out size as CuDnn.c:242-266 with
	nd = 3; the reference pins it with a host loop in Cuda/Wrappers/CuDnn.py:106-144."""
	x, w = np.asarray(x, dtype), np.asarray(w, dtype)
	stride, pad, dilation = _triple(stride), _triple(pad), _triple(dilation)
	N, C = x.shape[:2]
	K, Cg = w.shape[:2]
	fsize = w.shape[2:]
	outshape = tuple(conv_out_size(x.shape[2 + i], fsize[i], stride[i], pad[i], dilation[i]) for i in range(3))
	xp = np.pad(x, ((0, 0), (0, 0)) + tuple((p, p) for p in pad))
	win = _windows3d(xp, fsize, stride, dilation, outshape)
	Kg = K // groups
	y = np.empty((N, K) + outshape, dtype)
	for g in range(groups):
		y[:, g * Kg:(g + 1) * Kg] = np.einsum("ncdpqtrs,kctrs->nkdpq", win[:, g * Cg:(g + 1) * Cg], w[g * Kg:(g + 1) * Kg], optimize=True)
	if bias is not None:
		y += np.asarray(bias, dtype).reshape(1, K, 1, 1, 1)
	return y


def conv3d_bwd_params(x, dy, wshape, stride=1, pad=0, dilation=1, groups=1, dtype=np.float64):
	x, dy = np.asarray(x, dtype), np.asarray(dy, dtype)
	stride, pad, dilation = _triple(stride), _triple(pad), _triple(dilation)
	K, Cg = wshape[:2]
	fsize = tuple(wshape[2:])
	xp = np.pad(x, ((0, 0), (0, 0)) + tuple((p, p) for p in pad))
	win = _windows3d(xp, fsize, stride, dilation, dy.shape[2:])
	Kg = K // groups
	dw = np.empty(wshape, dtype)
	for g in range(groups):
		dw[g * Kg:(g + 1) * Kg] = np.einsum("ncdpqtrs,nkdpq->kctrs", win[:, g * Cg:(g + 1) * Cg], dy[:, g * Kg:(g + 1) * Kg], optimize=True)
	return dw, dy.sum(axis=(0, 2, 3, 4))


def conv3d_bwd_data(dy, w, inshape, stride=1, pad=0, dilation=1, groups=1, dtype=np.float64):
	dy, w = np.asarray(dy, dtype), np.asarray(w, dtype)
	stride, pad, dilation = _triple(stride), _triple(pad), _triple(dilation)
	N, C, D, H, W = inshape
	K, Cg, T, R, S = w.shape
	Kg = K // groups
	Do, P, Q = dy.shape[2:]
	dxp = np.zeros((N, C, D + 2 * pad[0], H + 2 * pad[1], W + 2 * pad[2]), dtype)
	for g in range(groups):
		dyg, wg = dy[:, g * Kg:(g + 1) * Kg], w[g * Kg:(g + 1) * Kg]
		for t in range(T):
			for r in range(R):
				for s in range(S):
					contrib = np.einsum("nkdpq,kc->ncdpq", dyg, wg[:, :, t, r, s], optimize=True)
					d0, h0, w0 = t * dilation[0], r * dilation[1], s * dilation[2]
					dxp[:, g * Cg:(g + 1) * Cg, d0:d0 + Do * stride[0]:stride[0], h0:h0 + P * stride[1]:stride[1],
						w0:w0 + Q * stride[2]:stride[2]] += contrib
	return dxp[:, :, pad[0]:pad[0] + D, pad[1]:pad[1] + H, pad[2]:pad[2] + W]


def pool3d(x, size=2, stride=2, pad=0, mode="max", dtype=np.float64):
	"""(fd, fh, fw) pooling, out size (in + 2 pad - size) / stride + 1 (CuDnnPool.c:24-41 with nd = 3).  Returns (y, argmax) where
	argmax is the flat in-volume index of the first maximum (row-major scan), -1 for the average modes."""
	x = np.asarray(x, dtype)
	size, stride, pad = _triple(size), _triple(stride), _triple(pad)
	N, C, D, H, W = x.shape
	out = tuple((x.shape[2 + i] + 2 * pad[i] - size[i]) // stride[i] + 1 for i in range(3))
	y = np.zeros((N, C) + out, dtype)
	arg = np.full((N, C) + out, -1, np.int64)
	for do in range(out[0]):
		d0, d1 = max(do * stride[0] - pad[0], 0), min(do * stride[0] - pad[0] + size[0], D)
		for p in range(out[1]):
			h0, h1 = max(p * stride[1] - pad[1], 0), min(p * stride[1] - pad[1] + size[1], H)
			for q in range(out[2]):
				w0, w1 = max(q * stride[2] - pad[2], 0), min(q * stride[2] - pad[2] + size[2], W)
				box = x[:, :, d0:d1, h0:h1, w0:w1].reshape(N, C, -1)
				if mode == "max":
					idx = box.argmax(axis=2)
					y[:, :, do, p, q] = np.take_along_axis(box, idx[..., None], axis=2)[..., 0]
					bd, bh, bw = d1 - d0, h1 - h0, w1 - w0
					arg[:, :, do, p, q] = ((d0 + idx // (bh * bw)) * H + h0 + (idx // bw) % bh) * W + w0 + idx % bw
				else:
					cnt = size[0] * size[1] * size[2] if mode == "avgWithPad" else box.shape[2]
					y[:, :, do, p, q] = box.sum(axis=2) / cnt
	return y, arg


def pool3d_bwd(x, dy, size=2, stride=2, pad=0, mode="max", dtype=np.float64):
	x, dy = np.asarray(x, dtype), np.asarray(dy, dtype)
	size, stride, pad = _triple(size), _triple(stride), _triple(pad)
	N, C, D, H, W = x.shape
	dx = np.zeros_like(x)
	_, arg = pool3d(x, size, stride, pad, mode, dtype)
	out = dy.shape[2:]
	flat = dx.reshape(N, C, -1)
	for do in range(out[0]):
		d0, d1 = max(do * stride[0] - pad[0], 0), min(do * stride[0] - pad[0] + size[0], D)
		for p in range(out[1]):
			h0, h1 = max(p * stride[1] - pad[1], 0), min(p * stride[1] - pad[1] + size[1], H)
			for q in range(out[2]):
				w0, w1 = max(q * stride[2] - pad[2], 0), min(q * stride[2] - pad[2] + size[2], W)
				if mode == "max":
					np.add.at(flat, (np.arange(N)[:, None], np.arange(C)[None, :], arg[:, :, do, p, q]), dy[:, :, do, p, q])
				else:
					cnt = size[0] * size[1] * size[2] if mode == "avgWithPad" else (d1 - d0) * (h1 - h0) * (w1 - w0)
					dx[:, :, d0:d1, h0:h1, w0:w1] += (dy[:, :, do, p, q] / cnt)[:, :, None, None, None]
	return dx


# ================================================================================================ training closure
def cross_entropy(scores, labels, weights=None, dtype=np.float64):
	"""Softmax over axis 1 + cross-entropy cost as the reference computes it (Cuda/Kernels/Costs.py:77-106,133-157,213-247):
	scores (N, C[, ...spatial]), labels (N[, ...spatial]) int.  Returns (error, grad) with
	grad = w_c * ((c == label) - p) / N  (ascent direction) and error = sum(-w_label * log p_label) / spatial."""
	x = np.asarray(scores, dtype)
	N, C = x.shape[:2]
	S = int(np.prod(x.shape[2:])) if x.ndim > 2 else 1
	z = x.reshape(N, C, S)
	p = np.exp(z - z.max(axis=1, keepdims=True))
	p /= p.sum(axis=1, keepdims=True)
	lab = np.asarray(labels).reshape(N, S)
	onehot = (np.arange(C).reshape(1, C, 1) == lab.reshape(N, 1, S)).astype(dtype)
	w = np.ones(C, dtype) if weights is None else np.asarray(weights, dtype)
	grad = w.reshape(1, C, 1) * (onehot - p) / N
	picked = np.take_along_axis(p, lab.reshape(N, 1, S), axis=1)[:, 0, :]
	error = float((-w[lab] * np.log(picked)).sum() / S)
	return error, grad.reshape(x.shape)


def nesterov_update(param, grad, mom, lr, mr, dtype=np.float64):
	"""reference: Cuda/Kernels/ElementWise.py:815-857 -- returns (param', mom')"""
	p, g, m = (np.asarray(a, dtype) for a in (param, grad, mom))
	return p + mr * mr * m + (1.0 + mr) * lr * g, mr * m + lr * g


def adam_update(param, grad, mg, ms, lr, fix1, fix2, eps, dtype=np.float64):
	"""reference: Cuda/Kernels/ElementWise.py:709-755 -- returns (param', mg', ms')"""
	p, g, a, s = (np.asarray(v, dtype) for v in (param, grad, mg, ms))
	a = a + fix1 * (g - a)
	s = s + fix2 * (g * g - s)
	return p + lr * a / (np.sqrt(s) + eps), a, s


def lrn(x, N, alpha, beta, K, across_maps, grad=None, dtype=np.float64):
	"""Local response normalisation and its gradient by the host formulas of the reference's tests
	(Cuda/Wrappers/CuDnnNorm.py:185-268): window [i - (N-1)//2, i + N - (N-1)//2) clipped, over maps (across_maps) or over the
	N x N neighbourhood inside a map.  Returns y, or (y, dx) when grad is given."""
	x = np.asarray(x, dtype)
	B, C, H, W = x.shape
	lb = (N - 1) // 2
	la = N - lb
	scale = alpha / N if across_maps else alpha / N ** 2

	def window_sum(t):
		out = np.zeros_like(t)
		if across_maps:
			for c in range(C):
				out[:, c] = t[:, max(0, c - lb):min(C, c + la)].sum(axis=1)
		else:
			for y in range(H):
				for xx in range(W):
					out[:, :, y, xx] = t[:, :, max(0, y - lb):min(H, y + la), max(0, xx - lb):min(W, xx + la)].sum(axis=(2, 3))
		return out

	norms = K + scale * window_sum(x * x)
	y = x / norms ** beta
	if grad is None:
		return y
	g = np.asarray(grad, dtype)
	dx = g / norms ** beta - 2.0 * beta * scale * x * window_sum(g * x / norms ** (beta + 1))
	return y, dx


# ---------------------------------------------------------------------------------------------------------- side modules
# pinned by tests/golden/ref_cuda_side.npz (the reference's NVRTC kernels / cuDNN on a B200, tests/golden_cases.py SIDE_CASES)
def prelu(x, slopes, shared=False):
	"""reference: Cuda/Kernels/PRelu.py:14-24 -- y = x * (x > 0 ? 1 : slope[c]); one slope for all maps when `shared`"""
	a = slopes.reshape(()) if shared else slopes.reshape((1, -1) + (1, ) * (x.ndim - 2))
	return x * np.where(x > 0, 1.0, a)


def prelu_bwd(x, dy, slopes, shared=False):
	"""reference: Cuda/Kernels/PRelu.py:26-56 -- (dx, dslopes); dslopes sums dy * x * (x <= 0) over everything but the map axis"""
	a = slopes.reshape(()) if shared else slopes.reshape((1, -1) + (1, ) * (x.ndim - 2))
	dx = dy * np.where(x > 0, 1.0, a)
	contrib = dy.astype(np.float64) * x * (x <= 0)
	ds = contrib.sum().reshape(1) if shared else contrib.sum(axis=(0, ) + tuple(range(2, x.ndim)))
	return dx, ds


def _pad_widths(ndim, pad):
	if ndim == 3:
		return ((0, 0), (0, 0), (pad[0], pad[1]))
	return ((0, 0), (0, 0), (pad[0], pad[1]), (pad[2], pad[3]))


def reflectpad(x, pad):
	"""reference: Cuda/Kernels/Pad.py:33-74 -- numpy's "reflect" mode; pad = (left, right) or (up, bottom, left, right)"""
	return np.pad(x, _pad_widths(x.ndim, pad), mode="reflect")


def reflectpad_bwd(dy, pad):
	"""the adjoint of reflectpad (reference: scatter with atomicAdd, Cuda/Kernels/Pad.py:77-139): every padded sample's gradient
	returns to the input element it was copied from"""
	inshape = tuple(n - sum(w) for n, w in zip(dy.shape, _pad_widths(dy.ndim, pad)))
	index = np.arange(int(np.prod(inshape))).reshape(inshape)
	src = np.pad(index, _pad_widths(dy.ndim, pad), mode="reflect")
	dx = np.zeros(int(np.prod(inshape)), dtype=np.float64)
	np.add.at(dx, src.ravel(), dy.astype(np.float64).ravel())
	return dx.reshape(inshape)


def embed(idx, W):
	"""reference: Cuda/Kernels/Embedder.py:11-22 -- rows of W; index -1 leaves a zero row"""
	out = W[np.maximum(idx, 0)].copy()
	out[idx == -1] = 0
	return out


def embed_bwd(idx, dy, W, scale):
	"""reference: Cuda/Kernels/Embedder.py:24-41 -- W[idx] += scale * dy (repeated words accumulate, -1 skipped)"""
	out = W.astype(np.float64).copy()
	keep = idx.ravel() != -1
	np.add.at(out, idx.ravel()[keep], scale * dy.reshape(-1, dy.shape[-1]).astype(np.float64)[keep])
	return out


def _scales(scale, n):
	return (scale, ) * n if isinstance(scale, int) else tuple(scale)


def upsample_nearest(x, scale):
	"""reference: Cuda/Kernels/Upsample.py:9-60 -- every input element fills its block of the output"""
	for axis, s in enumerate(_scales(scale, x.ndim - 2)):
		x = np.repeat(x, s, axis=2 + axis)
	return x


def upsample_nearest_bwd(dy, scale):
	"""reference: Cuda/Kernels/Upsample.py:26-43,62-96 -- block sums"""
	scales = _scales(scale, dy.ndim - 2)
	shape = list(dy.shape[:2])
	for n, s in zip(dy.shape[2:], scales):
		shape += [n // s, s]
	return dy.astype(np.float64).reshape(shape).sum(axis=tuple(range(3, len(shape), 2)))


def _lerp_axis(n_in, n_out):
	"""float32 source coordinates like the kernels (Upsample.py:108-118): i0, i0 + step, weight of the second tap"""
	ratio = np.float32((n_in - 1) / (n_out - 1))
	src = ratio * np.arange(n_out, dtype=np.float32)
	i0 = src.astype(np.int32)
	step = (i0 < n_in - 1).astype(np.int32)
	w1 = (src - i0.astype(np.float32)).astype(np.float32)
	return i0, step, w1


def upsample_linear(x, scale, quirk=True):
	"""reference: Cuda/Kernels/Upsample.py:100-139 (2-d), 186-252 (3-d) -- align-corners (tri)linear interpolation.  `quirk`: one tap
	of the reference's 3-d forward kernel is addressed with d1 * inw * inw instead of d1 * inh * inw (Upsample.py:241); it only
	shows when inh != inw"""
	scales = _scales(scale, x.ndim - 2)
	if x.ndim == 4:
		x = x[:, :, None]
		scales = (1, ) + scales
	N, C, D, H, W = x.shape
	oD, oH, oW = D * scales[0], H * scales[1], W * scales[2]
	three_d = scales[0] != 1 or D != 1
	h0, hs, hw1 = _lerp_axis(H, oH)
	w0, ws, ww1 = _lerp_axis(W, oW)
	if three_d:
		d0, dstep, dw1 = _lerp_axis(D, oD)
	else:
		d0, dstep, dw1 = np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.float32)
	f = x.astype(np.float32)
	flat = f.reshape(N, C, -1)
	D0, H0, W0 = np.meshgrid(d0, h0, w0, indexing="ij")
	DS, HS, WS = np.meshgrid(dstep, hs, ws, indexing="ij")
	Dw, Hw, Ww = np.meshgrid(dw1, hw1, ww1, indexing="ij")

	def tap(d, h, w, quirky=False):
		if quirky and quirk and three_d:
			lin = d * W * W + h * W + w
			lin = np.where(lin >= D * H * W, d * H * W + h * W + w, lin)
		else:
			lin = d * H * W + h * W + w
		return flat[:, :, lin]

	one = np.float32(1)
	near = (one - Hw) * ((one - Ww) * tap(D0, H0, W0) + Ww * tap(D0, H0, W0 + WS, True)) + \
		   Hw * ((one - Ww) * tap(D0, H0 + HS, W0) + Ww * tap(D0, H0 + HS, W0 + WS))
	if not three_d:
		return near[:, :, 0]
	far = (one - Hw) * ((one - Ww) * tap(D0 + DS, H0, W0) + Ww * tap(D0 + DS, H0, W0 + WS)) + \
		  Hw * ((one - Ww) * tap(D0 + DS, H0 + HS, W0) + Ww * tap(D0 + DS, H0 + HS, W0 + WS))
	return (one - Dw) * near + Dw * far


def upsample_linear_bwd(dy, scale):
	"""reference: Cuda/Kernels/Upsample.py:141-184, 254-296 -- each output gradient goes to its 4 / 8 taps with the forward weights"""
	scales = _scales(scale, dy.ndim - 2)
	two_d = dy.ndim == 4
	if two_d:
		dy = dy[:, :, None]
		scales = (1, ) + scales
	N, C, oD, oH, oW = dy.shape
	D, H, W = oD // scales[0], oH // scales[1], oW // scales[2]
	h0, hs, hw1 = _lerp_axis(H, oH)
	w0, ws, ww1 = _lerp_axis(W, oW)
	if two_d:
		d0, dstep, dw1 = np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.float32)
	else:
		d0, dstep, dw1 = _lerp_axis(D, oD)
	D0, H0, W0 = np.meshgrid(d0, h0, w0, indexing="ij")
	DS, HS, WS = np.meshgrid(dstep, hs, ws, indexing="ij")
	Dw, Hw, Ww = (a.astype(np.float64) for a in np.meshgrid(dw1, hw1, ww1, indexing="ij"))
	g = dy.astype(np.float64).reshape(N * C, -1)
	dx = np.zeros((N * C, D * H * W), dtype=np.float64)
	for dd, wd in ((D0, 1 - Dw), (D0 + DS, Dw)):
		for hh, wh in ((H0, 1 - Hw), (H0 + HS, Hw)):
			for ww, wx in ((W0, 1 - Ww), (W0 + WS, Ww)):
				lin = (dd * H * W + hh * W + ww).ravel()
				weight = (wd * wh * wx).ravel()
				for row in range(N * C):
					np.add.at(dx[row], lin, weight * g[row])
	dx = dx.reshape(N, C, D, H, W)
	return dx[:, :, 0] if two_d else dx


def lcn(x, means, N, alpha, beta, K, grad=None):
	"""mapLRN with a means tensor = cudnnDivisiveNormalization (CuDnnNorm.c:329-527; host formulas of Modules/LCN.py:62-143).
	-> y, or (dx, dmeans) when `grad` is given.  Window of position i: [i - lb, i + la) clipped, lb = (N - 1) // 2, la = N - lb."""
	x64, m64 = x.astype(np.float64), means.astype(np.float64)
	B, C, H, W = x.shape
	lb = (N - 1) // 2
	la = N - lb
	norm = np.empty_like(x64)
	sumdiff = np.empty_like(x64)
	for h in range(H):
		for w in range(W):
			win = x64[:, :, max(0, h - lb):min(H, h + la), max(0, w - lb):min(W, w + la)] - m64[:, :, h:h + 1, w:w + 1]
			norm[:, :, h, w] = K + alpha / N ** 2 * (win ** 2).sum(axis=(2, 3))
			sumdiff[:, :, h, w] = win.sum(axis=(2, 3))
	if grad is None:
		return x64 * norm ** -beta
	g = grad.astype(np.float64)
	t = g * x64 * norm ** -(beta + 1)
	k = 2.0 * alpha * beta / N ** 2
	dx = g * norm ** -beta
	for h in range(H):
		for w in range(W):
			sl = (slice(None), slice(None), slice(max(0, h - lb), min(H, h + la)), slice(max(0, w - lb), min(W, w + la)))
			dx[:, :, h, w] -= k * (x64[:, :, h, w] * t[sl].sum(axis=(2, 3)) - (t[sl] * m64[sl]).sum(axis=(2, 3)))
	return dx, k * t * sumdiff
