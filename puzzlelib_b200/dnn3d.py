"""3-d convolution and pooling of the B200 backend, expressed through the 2-d kernels.

Reference surface: the same `dnn.convNd / convNdBackwardData / convNdBackwardParams / poolNd / poolNdBackward` entry
points called with 5-d tensors (CuDnn.c:242-322 shape rules for nd = 3, CuDnnPool.c:24-41; tests
Cuda/Wrappers/CuDnn.py:106-201,414-451).

Convolution.  For one output depth d' the 3-d filter sees the T input slices d0 .. d0+T-1 (d0 = d'*stride_d - pad_d).  Folding
the filter depth into the channel axis makes that a 2-d convolution over C*T channels: the filter (K, C/G, T, R, S) IS a
contiguous (K, (C/G)*T, R, S) tensor, and the folded input is a pitched copy of x[n, c, d0:d0+T] (T*H*W contiguous elements
per (n, c)).  So each output depth costs one pitched copy in, one tcgen05 implicit-GEMM launch with a T-times longer
reduction, and one pitched copy out; backward-data scatters the folded gradient back with a pitched add, backward-filter
accumulates over the output depths with beta = 1.

Pooling is separable: a (fd, fh, fw) window is an (fh, fw) pooling of every slice followed by an (fd, 1) pooling along the
depth axis -- two launches of the 2-d kernels on reshaped views, no copies.  Max, average-with-pad and average-without-pad
all factor exactly (the valid count of a clipped box is the product of the per-axis counts); backward is the chain rule
through the two stages, and the first-maximum tie rule composes to the row-major (d, h, w) first maximum.
"""
import numpy as np

from .driver import lib, check, dtypeCode, Conv2dDesc
from .gpuarray import GPUArray
from ctypes import byref


def _outsize(insize, fsize, stride, pad, dilation):
	ext = insize + 2 * pad - dilation * (fsize - 1) - 1
	if ext < 0:
		raise ValueError("invalid input map size")
	return ext // stride + 1


def _copy2d(dst, dpitch, src, spitch, width, rows, itemsize):
	if width > 0 and rows > 0:
		check(lib.pz_memcpy2d(dst, dpitch * itemsize, src, spitch * itemsize, width * itemsize, rows, 0, None))


class _Fold:
	"""Geometry of the depth folding for one 3-d convolution."""

	def __init__(self, xshape, Wshape, yshape, stride, pad, dilation, groups):
		self.N, self.C, self.D, self.H, self.W = xshape
		self.K, self.Cg, self.T, self.R, self.S = Wshape
		_, _, self.Do, self.P, self.Q = yshape
		self.stride, self.pad, self.dilation, self.groups = stride, pad, dilation, groups
		self.HW, self.PQ = self.H * self.W, self.P * self.Q

	def desc(self):
		return Conv2dDesc(self.N, self.C * self.T, self.H, self.W, self.K, self.R, self.S, self.P, self.Q, self.stride[1],
						  self.stride[2], self.pad[1], self.pad[2], self.dilation[1], self.dilation[2], self.groups)

	def taps(self, do):
		"""(t, d) pairs of the filter taps that fall inside the input for output depth `do`."""
		d0 = do * self.stride[0] - self.pad[0]
		return [(t, d0 + t * self.dilation[0]) for t in range(self.T) if 0 <= d0 + t * self.dilation[0] < self.D]

	def runs(self, do):
		"""The valid taps grouped into runs of consecutive slices: (t_first, d_first, count).  With depth dilation 1 there is
		one run (one pitched copy); otherwise one run per tap."""
		taps = self.taps(do)
		if not taps:
			return []
		if self.dilation[0] == 1:
			return [(taps[0][0], taps[0][1], len(taps))]
		return [(t, d, 1) for t, d in taps]


def _gather(fold, x, xs, do, itemsize):
	"""xs[n, (c, t)] = x[n, c, d0 + t*dil] (zero where the tap leaves the input)."""
	runs = fold.runs(do)
	covered = sum(r[2] for r in runs)
	if covered < fold.T:
		check(lib.pz_memset8(xs.ptr, 0, xs.nbytes, None))
	for t0, d, count in runs:
		_copy2d(xs.ptr + t0 * fold.HW * itemsize, fold.T * fold.HW, x.ptr + d * fold.HW * itemsize, fold.D * fold.HW,
				count * fold.HW, fold.N * fold.C, itemsize)


def conv3d(dnn, data, W, bias, stride, pad, dilation, groups, out, allocator):
	if data.shape[1] != W.shape[1] * groups:
		raise ValueError("invalid number of input maps")
	outshape = (data.shape[0], W.shape[0]) + tuple(
		_outsize(data.shape[2 + i], W.shape[2 + i], stride[i], pad[i], dilation[i]) for i in range(3))
	if out is None:
		out = GPUArray(outshape, data.dtype, allocator=allocator)
	elif out.shape != outshape or out.dtype != data.dtype:
		raise ValueError("invalid output gpuarray data layout")

	fold = _Fold(data.shape, W.shape, outshape, stride, pad, dilation, groups)
	itemsize, code = data.dtype.itemsize, dtypeCode(data.dtype)
	xs = GPUArray((fold.N, fold.C * fold.T, fold.H, fold.W), data.dtype, allocator=allocator)
	ys = GPUArray((fold.N, fold.K, fold.P, fold.Q), data.dtype, allocator=allocator)
	desc = fold.desc()
	for do in range(fold.Do):
		_gather(fold, data, xs, do, itemsize)
		check(lib.pz_conv2d_fprop(code, byref(desc), xs.ptr, W.ptr, bias.ptr if bias is not None else None, ys.ptr, None))
		_copy2d(out.ptr + do * fold.PQ * itemsize, fold.Do * fold.PQ, ys.ptr, fold.PQ, fold.PQ, fold.N * fold.K, itemsize)
	return out


def conv3dBackwardData(dnn, grad, W, bias, data, stride, pad, dilation, postpad, groups, out, allocator):
	inmaps = W.shape[1] * groups
	if data is not None:
		inshape = data.shape
	else:
		inshape = (grad.shape[0], inmaps) + tuple(
			(grad.shape[2 + i] - 1) * stride[i] + dilation[i] * (W.shape[2 + i] - 1) - 2 * pad[i] + 1 + postpad[i] for i in range(3))
	if out is None:
		out = GPUArray(inshape, grad.dtype, allocator=allocator)
	elif out.shape != inshape or out.dtype != grad.dtype:
		raise ValueError("invalid output gpuarray data layout")

	fold = _Fold(inshape, W.shape, grad.shape, stride, pad, dilation, groups)
	itemsize, code = grad.dtype.itemsize, dtypeCode(grad.dtype)
	check(lib.pz_memset8(out.ptr, 0, out.nbytes, None))
	dxs = GPUArray((fold.N, fold.C * fold.T, fold.H, fold.W), grad.dtype, allocator=allocator)
	dys = GPUArray((fold.N, fold.K, fold.P, fold.Q), grad.dtype, allocator=allocator)
	desc = fold.desc()
	for do in range(fold.Do):
		runs = fold.runs(do)
		if not runs:
			continue
		_copy2d(dys.ptr, fold.PQ, grad.ptr + do * fold.PQ * itemsize, fold.Do * fold.PQ, fold.PQ, fold.N * fold.K, itemsize)
		check(lib.pz_conv2d_dgrad(code, byref(desc), dys.ptr, W.ptr, None, dxs.ptr, None, 0, None))
		for t0, d, count in runs:
			check(lib.pz_add2d(code, out.ptr + d * fold.HW * itemsize, fold.D * fold.HW, dxs.ptr + t0 * fold.HW * itemsize,
							   fold.T * fold.HW, count * fold.HW, fold.N * fold.C, None))
	if bias is not None:
		# deconvolution forward: the bias of the (N, C, D, H, W) result, added per channel
		_addChannelBias(out, bias)
	return out


def _addChannelBias(tensor, bias):
	"""tensor[n, c, ...] += bias[c]: the (N*C, S) view with the bias vector tiled down the rows (pz_addvec2mat, axis 0)."""
	N, C = tensor.shape[:2]
	S = int(np.prod(tensor.shape[2:]))
	check(lib.pz_addvec2mat(dtypeCode(tensor.dtype), tensor.ptr, tensor.ptr, bias.ptr, 1, N * C, S, 0, C, None))


def conv3dBackwardParams(dnn, data, grad, W, stride, pad, dilation, groups, withbias, deconv, wgrad, bgrad, scale, momentum,
						 allocator):
	if wgrad is None:
		wgrad = GPUArray.zeros(W.shape, W.dtype, allocator=allocator)
	elif wgrad.shape != W.shape or wgrad.dtype != W.dtype:
		raise ValueError("invalid output gpuarray data layout")

	fold = _Fold(data.shape, W.shape, grad.shape, stride, pad, dilation, groups)
	itemsize, code = data.dtype.itemsize, dtypeCode(data.dtype)
	xs = GPUArray((fold.N, fold.C * fold.T, fold.H, fold.W), data.dtype, allocator=allocator)
	dys = GPUArray((fold.N, fold.K, fold.P, fold.Q), data.dtype, allocator=allocator)
	desc = fold.desc()
	for do in range(fold.Do):
		_gather(fold, data, xs, do, itemsize)
		_copy2d(dys.ptr, fold.PQ, grad.ptr + do * fold.PQ * itemsize, fold.Do * fold.PQ, fold.PQ, fold.N * fold.K, itemsize)
		check(lib.pz_conv2d_wgrad(code, byref(desc), xs.ptr, dys.ptr, wgrad.ptr, scale, momentum if do == 0 else 1.0, None))

	if not withbias:
		return wgrad
	side = data if deconv else grad
	if bgrad is None:
		bgrad = GPUArray.zeros((side.shape[1], ), side.dtype, allocator=allocator)
	check(lib.pz_bias_grad(code, side.ptr, bgrad.ptr, side.shape[0], side.shape[1], int(np.prod(side.shape[2:])), scale, momentum,
						   None))
	return wgrad, bgrad


# ------------------------------------------------------------------------------------------------------------ pooling
def _poolShapes(shape, size, stride, pad):
	out = []
	for i in range(3):
		ext = shape[2 + i] + 2 * pad[i]
		if ext < size[i]:
			raise ValueError("invalid input map size on dim #%d" % (i + 1))
		out.append((ext - size[i]) // stride[i] + 1)
	return tuple(out)


def _pool2d(code, mode, src, dst, planes, H, W, OH, OW, size, stride, pad):
	check(lib.pz_pool2d_fwd(code, int(mode), src, dst, planes, H, W, OH, OW, size[0], size[1], stride[0], stride[1], pad[0], pad[1],
							None))


def pool3d(dnn, data, size, stride, pad, mode, out, allocator, keep=None):
	N, C, D, H, W = data.shape
	Do, P, Q = _poolShapes(data.shape, size, stride, pad)
	outshape = (N, C, Do, P, Q)
	if out is None:
		out = GPUArray(outshape, data.dtype, allocator=allocator)
	elif out.shape != outshape or out.dtype != data.dtype:
		raise ValueError("invalid output gpuarray data layout")
	code = dtypeCode(data.dtype)
	# stage 1: (fh, fw) pooling of every slice; stage 2: (fd, 1) pooling along the depth axis of the (N*C, D, P*Q) view
	mid = GPUArray((N, C, D, P, Q), data.dtype, allocator=allocator)
	_pool2d(code, mode, data.ptr, mid.ptr, N * C * D, H, W, P, Q, size[1:], stride[1:], pad[1:])
	_pool2d(code, mode, mid.ptr, out.ptr, N * C, D, P * Q, Do, P * Q, (size[0], 1), (stride[0], 1), (pad[0], 0))
	if keep is not None:
		keep.append(mid)
	return out


def pool3dBackward(dnn, grad, indata, outdata, size, stride, pad, mode, out, allocator):
	N, C, D, H, W = indata.shape
	_, _, Do, P, Q = outdata.shape
	if out is None:
		out = GPUArray(indata.shape, indata.dtype, allocator=allocator)
	elif out.shape != indata.shape or out.dtype != indata.dtype:
		raise ValueError("invalid output gpuarray data layout")
	code = dtypeCode(indata.dtype)
	# recompute the stage-1 result (the reference API hands only x, y, dy to the backward pass)
	mid = GPUArray((N, C, D, P, Q), indata.dtype, allocator=allocator)
	_pool2d(code, mode, indata.ptr, mid.ptr, N * C * D, H, W, P, Q, size[1:], stride[1:], pad[1:])
	dmid = GPUArray((N, C, D, P, Q), indata.dtype, allocator=allocator)
	check(lib.pz_pool2d_bwd(code, int(mode), mid.ptr, outdata.ptr, grad.ptr, dmid.ptr, N * C, D, P * Q, Do, P * Q, size[0], 1,
							stride[0], 1, pad[0], 0, None))
	check(lib.pz_pool2d_bwd(code, int(mode), indata.ptr, mid.ptr, dmid.ptr, out.ptr, N * C * D, H, W, P, Q, size[1], size[2], stride[1],
							stride[2], pad[1], pad[2], None))
	return out
