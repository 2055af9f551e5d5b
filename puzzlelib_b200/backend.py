"""The `CudaBackend`-shaped object of the B200 backend -- the drop-in seam.

The reference reaches the GPU only through `PuzzleLib.Cuda.Backend.getBackend(deviceIdx, initmode)` and the
attributes of the object it returns (reference: Cuda/Backend.py:50-66,353-370; Cuda/GPUBackend.py:17-215;
consumers: Backend/gpuarray.py:60-113, Backend/Dnn.py:159-338, Backend/Blas.py:43-102, Backend/Kernels/*.py).
This module provides the same function and an object with the same attribute names and call signatures, where
every operation is one call into the C-ABI of libpzb200.so (hand-written sm_100a kernels).  There is no cuDNN,
cuBLAS, NVRTC or CPU fallback behind it.

`dnn`      mirrors CuDnn.DnnContext    (Cuda/Source/Libs/CuDnn.c, CuDnnPool.c, CuDnnNorm.c)
`blas`     mirrors CuBlas.BlasContext  (Cuda/Source/Libs/CuBlas.c)
`matmod`   mirrors MatModule           (Cuda/Kernels/MatVec.py)
`poolmod`  mirrors PoolModule          (Cuda/Kernels/Pool.py)
`*Ker`     mirror the kernel factories of Cuda/Kernels/ElementWise.py
"""
from enum import Enum
from collections import OrderedDict
from ctypes import byref

import ctypes

import numpy as np

from . import driver, gpuarray
from .driver import lib, check, dtypeCode, Conv2dDesc, CuDnnError, CuBlasError
from .gpuarray import GPUArray

_f32 = np.dtype(np.float32)
_i32 = np.dtype(np.int32)


def prod(seq):
	n = 1
	for d in seq:
		n *= int(d)
	return n


def _seq(val, nd, default, name):
	"""int / sequence / None -> tuple of nd ints (reference: Libs.h:91-119 CuDnn_unpackIntSequence)."""
	if val is None:
		return (default, ) * nd
	if isinstance(val, (int, np.integer)):
		return (int(val), ) * nd
	val = tuple(int(v) for v in val)
	if len(val) != nd:
		raise ValueError("%s must be int or %d-elem tuple" % (name, nd))
	return val


def _requireArray(ary, name):
	if not isinstance(ary, GPUArray):
		raise TypeError("%s must be a GPUArray (got %s)" % (name, type(ary).__name__))
	if not ary.contiguous:
		raise ValueError("invalid %s gpuarray data layout" % name)


def _checkOut(out, shape, dtype):
	# reference: CuDnn.c:77-134 -- an `out` array must match shape and dtype exactly
	_requireArray(out, "output")
	if out.shape != tuple(shape) or out.dtype != dtype:
		raise ValueError("invalid output gpuarray data layout")
	return out


# ============================================================================================================ dnn
class DnnContext:
	"""Hand-written sm_100a replacements of the cuDNN entry points the reference uses."""

	def __init__(self, backend):
		self.backend = backend
		self.tensorOps = True

	@staticmethod
	def getVersion():
		return int(lib.pz_version())

	def enableTensorOps(self, enable):
		"""reference: CuDnn_Context_enableTensorOps (CuDnn.c:61-74; the reference switches it on, Cuda/GPUBackend.py:152).
		True (default): float32 tensors are contracted as TF32 tensor-core products with fp32 accumulation -- within 1e-3 of
		full fp32.  False: exact mode -- float32 convolutions and GEMMs run as fp32 FMAs on the CUDA cores: the accuracy cuDNN /
		cuBLAS give the reference on this stack (its own unit tests assume it), at a fraction of the speed.  Process-wide:
		`dnn` and `blas` share the switch."""
		self.tensorOps = bool(enable)
		if self.backend.blas is not None:
			self.backend.blas.tensorOps = self.tensorOps
		check(lib.pz_set_exact_fp32(0 if enable else 1))
		return self

	# ------------------------------------------------------------------------------------------ convolution
	@staticmethod
	def _convDesc(datashape, Wshape, outshape, stride, pad, dilation, groups):
		N, C, H, W = datashape
		K, _, R, S = Wshape
		return Conv2dDesc(N, C, H, W, K, R, S, outshape[2], outshape[3], stride[0], stride[1], pad[0], pad[1],
						  dilation[0], dilation[1], groups)

	@staticmethod
	def _check4d(ary, name):
		_requireArray(ary, name)
		if ary.ndim != 4:
			raise ValueError("invalid %s gpuarray dims" % name)

	@staticmethod
	def _is3d(*arys):
		"""5-d tensors take the 3-d decomposition of dnn3d.py (cuDNN descriptors are 4-d or 5-d, Libs.h:46-56)."""
		for ary in arys:
			_requireArray(ary, "tensor")
		nd = {ary.ndim for ary in arys}
		if nd == {5}:
			return True
		if 5 in nd:
			raise ValueError("mixed 4-d / 5-d gpuarray dims")
		return False

	def convNd(self, data, W, bias=None, stride=1, pad=0, dilation=1, groups=1, algo=0, out=None, allocator=None):
		"""reference: CuDnn_Context_pyConvNd, CuDnn.c:457-514 (out shape :242-266)"""
		if self._is3d(data, W):
			from . import dnn3d
			if data.dtype != W.dtype:
				raise ValueError("invalid W gpuarray data layout")
			return dnn3d.conv3d(self, data, W, bias, _seq(stride, 3, 1, "stride"), _seq(pad, 3, 0, "pad"),
								_seq(dilation, 3, 1, "dilation"), groups, out, allocator)
		self._check4d(data, "data")
		self._check4d(W, "W")
		stride, pad, dilation = _seq(stride, 2, 1, "stride"), _seq(pad, 2, 0, "pad"), _seq(dilation, 2, 1, "dilation")

		if data.dtype != W.dtype:
			raise ValueError("invalid W gpuarray data layout")
		if data.shape[1] != W.shape[1] * groups:
			raise ValueError("invalid number of input maps")
		if W.shape[0] % groups != 0:
			raise ValueError("invalid number of output maps")

		outshape = [data.shape[0], W.shape[0]]
		for i in range(2):
			ext = data.shape[2 + i] + 2 * pad[i] - dilation[i] * (W.shape[2 + i] - 1) - 1
			if ext < 0:
				raise ValueError("invalid input map size on dim #%d" % (i + 1))
			outshape.append(ext // stride[i] + 1)
		outshape = tuple(outshape)

		if bias is not None:
			_requireArray(bias, "bias")
			if bias.ndim != 1 or bias.shape[0] != W.shape[0] or bias.dtype != data.dtype:
				raise ValueError("invalid bias gpuarray data layout")

		out = GPUArray(outshape, data.dtype, allocator=allocator) if out is None else _checkOut(out, outshape, data.dtype)
		desc = self._convDesc(data.shape, W.shape, outshape, stride, pad, dilation, groups)

		check(lib.pz_conv2d_fprop(dtypeCode(data.dtype), byref(desc), data.ptr, W.ptr, bias.ptr if bias is not None else None,
								  out.ptr, None))
		return out

	def convNdBackwardData(self, grad, W, bias=None, data=None, stride=1, pad=0, dilation=1, postpad=0, groups=1,
						   algo=0, out=None, allocator=None):
		"""reference: CuDnn_Context_pyConvNdBackwardData, CuDnn.c:579-649 (in shape :269-322)"""
		if self._is3d(grad, W):
			from . import dnn3d
			if grad.dtype != W.dtype or grad.shape[1] != W.shape[0]:
				raise ValueError("invalid W gpuarray data layout")
			return dnn3d.conv3dBackwardData(self, grad, W, bias, data, _seq(stride, 3, 1, "stride"), _seq(pad, 3, 0, "pad"),
											_seq(dilation, 3, 1, "dilation"), _seq(postpad, 3, 0, "postpad"), groups, out, allocator)
		self._check4d(grad, "grad")
		self._check4d(W, "W")
		stride, pad, dilation = _seq(stride, 2, 1, "stride"), _seq(pad, 2, 0, "pad"), _seq(dilation, 2, 1, "dilation")
		postpad = _seq(postpad, 2, 0, "postpad")

		if grad.dtype != W.dtype:
			raise ValueError("invalid W gpuarray data layout")
		if grad.shape[1] != W.shape[0]:
			raise ValueError("invalid number of output maps")

		inmaps = W.shape[1] * groups
		if data is not None:
			self._check4d(data, "data")
			inshape = data.shape
			if inshape[0] != grad.shape[0] or inshape[1] != inmaps:
				raise ValueError("invalid data gpuarray dims")
		else:
			inshape = (grad.shape[0], inmaps) + tuple(
				(grad.shape[2 + i] - 1) * stride[i] + dilation[i] * (W.shape[2 + i] - 1) - 2 * pad[i] + 1 + postpad[i]
				for i in range(2)
			)

		if bias is not None:
			_requireArray(bias, "bias")
			if bias.ndim != 1 or bias.shape[0] != inmaps or bias.dtype != grad.dtype:
				raise ValueError("invalid bias gpuarray data layout")

		out = GPUArray(inshape, grad.dtype, allocator=allocator) if out is None else _checkOut(out, inshape, grad.dtype)
		desc = self._convDesc(inshape, W.shape, grad.shape, stride, pad, dilation, groups)

		code = dtypeCode(grad.dtype)
		wsbytes = int(lib.pz_conv2d_dgrad_workspace(code, byref(desc)))
		workspace = GPUArray((max(wsbytes, 1), ), np.uint8, allocator=allocator) if wsbytes > 0 else None

		check(lib.pz_conv2d_dgrad(code, byref(desc), grad.ptr, W.ptr, bias.ptr if bias is not None else None, out.ptr,
								  workspace.ptr if workspace is not None else None, wsbytes, None))
		return out

	def convNdBackwardParams(self, data, grad, W, stride=1, pad=0, dilation=1, groups=1, withbias=False, deconv=False,
							 wgrad=None, bgrad=None, scale=1.0, momentum=0.0, algo=0, allocator=None):
		"""reference: CuDnn_Context_pyConvNdBackwardParams, CuDnn.c:722-800; wgrad / bgrad accumulate IN PLACE with
		alpha = scale, beta = momentum (:682-685, :388)"""
		if self._is3d(data, grad, W):
			from . import dnn3d
			if data.dtype != grad.dtype or data.dtype != W.dtype:
				raise ValueError("invalid gpuarray data layout")
			return dnn3d.conv3dBackwardParams(self, data, grad, W, _seq(stride, 3, 1, "stride"), _seq(pad, 3, 0, "pad"),
											  _seq(dilation, 3, 1, "dilation"), groups, withbias, deconv, wgrad, bgrad, scale, momentum,
											  allocator)
		self._check4d(data, "data")
		self._check4d(grad, "grad")
		self._check4d(W, "W")
		stride, pad, dilation = _seq(stride, 2, 1, "stride"), _seq(pad, 2, 0, "pad"), _seq(dilation, 2, 1, "dilation")

		if data.dtype != grad.dtype or data.dtype != W.dtype:
			raise ValueError("invalid gpuarray data layout")
		if data.shape[1] != W.shape[1] * groups or grad.shape[1] != W.shape[0] or data.shape[0] != grad.shape[0]:
			raise ValueError("invalid number of maps")

		if wgrad is None:
			wgrad = GPUArray.zeros(W.shape, W.dtype, allocator=allocator)
		else:
			_checkOut(wgrad, W.shape, W.dtype)

		desc = self._convDesc(data.shape, W.shape, grad.shape, stride, pad, dilation, groups)
		code = dtypeCode(data.dtype)
		if driver.gradientWriteHook is not None:
			driver.gradientWriteHook(wgrad)
		check(lib.pz_conv2d_wgrad(code, byref(desc), data.ptr, grad.ptr, wgrad.ptr, scale, momentum, None))

		if not withbias:
			return wgrad

		# the bias belongs to the OUTPUT side of the layer: `grad` for a conv, `data` for a deconv (CuDnn.c:689-690,774)
		side = data if deconv else grad
		if bgrad is None:
			bgrad = GPUArray.zeros((side.shape[1], ), side.dtype, allocator=allocator)
		else:
			_checkOut(bgrad, (side.shape[1], ), side.dtype)

		if driver.gradientWriteHook is not None:
			driver.gradientWriteHook(bgrad)
		check(lib.pz_bias_grad(code, side.ptr, bgrad.ptr, side.shape[0], side.shape[1], prod(side.shape[2:]), scale,
							   momentum, None))
		return wgrad, bgrad

	# ------------------------------------------------------------------------------------------ pooling
	def poolNd(self, data, size=2, stride=2, pad=0, mode=0, out=None, allocator=None):
		"""reference: CuDnn_Context_pyPoolNd, CuDnnPool.c:102-152"""
		if self._is3d(data):
			from . import dnn3d
			return dnn3d.pool3d(self, data, _seq(size, 3, 2, "size"), _seq(stride, 3, 2, "stride"), _seq(pad, 3, 0, "pad"), mode, out,
								allocator)
		self._check4d(data, "data")
		size, stride, pad = _seq(size, 2, 2, "size"), _seq(stride, 2, 2, "stride"), _seq(pad, 2, 0, "pad")

		outshape = [data.shape[0], data.shape[1]]
		for i in range(2):
			ext = data.shape[2 + i] + 2 * pad[i]
			if ext < size[i]:
				raise ValueError("invalid input map size on dim #%d" % (i + 1))
			outshape.append((ext - size[i]) // stride[i] + 1)
		outshape = tuple(outshape)

		out = GPUArray(outshape, data.dtype, allocator=allocator) if out is None else _checkOut(out, outshape, data.dtype)
		check(lib.pz_pool2d_fwd(dtypeCode(data.dtype), int(mode), data.ptr, out.ptr, data.shape[0] * data.shape[1],
								data.shape[2], data.shape[3], outshape[2], outshape[3], size[0], size[1], stride[0], stride[1],
								pad[0], pad[1], None))
		return out

	def poolNdBackward(self, grad, indata, outdata, size=2, stride=2, pad=0, mode=0, out=None, allocator=None):
		"""reference: CuDnn_Context_pyPoolNdBackward, CuDnnPool.c:197-245"""
		if self._is3d(grad, indata, outdata):
			from . import dnn3d
			return dnn3d.pool3dBackward(self, grad, indata, outdata, _seq(size, 3, 2, "size"), _seq(stride, 3, 2, "stride"),
										_seq(pad, 3, 0, "pad"), mode, out, allocator)
		self._check4d(grad, "grad")
		self._check4d(indata, "indata")
		self._check4d(outdata, "outdata")
		size, stride, pad = _seq(size, 2, 2, "size"), _seq(stride, 2, 2, "stride"), _seq(pad, 2, 0, "pad")

		if grad.shape != outdata.shape or grad.dtype != indata.dtype or outdata.dtype != indata.dtype:
			raise ValueError("invalid grad gpuarray data layout")

		out = GPUArray(indata.shape, indata.dtype, allocator=allocator) if out is None else \
			_checkOut(out, indata.shape, indata.dtype)
		check(lib.pz_pool2d_bwd(dtypeCode(indata.dtype), int(mode), indata.ptr, outdata.ptr, grad.ptr, out.ptr,
								indata.shape[0] * indata.shape[1], indata.shape[2], indata.shape[3], outdata.shape[2],
								outdata.shape[3], size[0], size[1], stride[0], stride[1], pad[0], pad[1], None))
		return out

	# ------------------------------------------------------------------------------------------ softmax
	@staticmethod
	def _ncs(ary):
		if ary.ndim < 2:
			raise ValueError("invalid data gpuarray dims")
		return ary.shape[0], ary.shape[1], prod(ary.shape[2:])

	def softmaxNd(self, data, mode=1, algo=1, out=None, allocator=None):
		"""reference: CuDnn_Context_pySoftmaxNd, CuDnn.c:1005-1048 (SOFTMAX_ACCURATE)"""
		_requireArray(data, "data")
		N, C, S = self._ncs(data)
		out = GPUArray(data.shape, data.dtype, allocator=allocator) if out is None else _checkOut(out, data.shape, data.dtype)
		check(lib.pz_softmax_fwd(dtypeCode(data.dtype), int(mode), data.ptr, out.ptr, N, C, S, None))
		return out

	def softmaxNdBackward(self, grad, outdata, mode=1, algo=1, out=None, allocator=None):
		"""reference: CuDnn_Context_pySoftmaxNdBackward, CuDnn.c:1087-1131"""
		_requireArray(grad, "grad")
		_requireArray(outdata, "outdata")
		if grad.shape != outdata.shape or grad.dtype != outdata.dtype:
			raise ValueError("invalid grad gpuarray data layout")
		N, C, S = self._ncs(grad)
		out = GPUArray(grad.shape, grad.dtype, allocator=allocator) if out is None else _checkOut(out, grad.shape, grad.dtype)
		check(lib.pz_softmax_bwd(dtypeCode(grad.dtype), int(mode), outdata.ptr, grad.ptr, out.ptr, N, C, S, None))
		return out

	# ------------------------------------------------------------------------------------------ memory reorganisation
	# (reference: Cuda/Source/Libs/CuDnnMemory.c, Backend/Memory.py:43-66)
	def transpose(self, data, axes=None, out=None, allocator=None):
		_requireArray(data, "data")
		data.enforceContiguous()
		ndim = data.ndim
		axes = tuple(reversed(range(ndim))) if axes is None else tuple(int(a) % ndim if ndim else 0 for a in axes)
		if sorted(axes) != list(range(ndim)):
			raise ValueError("axes is not a permutation of the tensor's axes")

		shape = tuple(data.shape[a] for a in axes)
		if out is None:
			out = GPUArray(shape, data.dtype, allocator=allocator)
		else:
			_checkOut(out, shape, data.dtype)

		if ndim == 0 or data.size == 0:
			return out

		elemstrides = [st // data.dtype.itemsize for st in data.strides]
		oshape = (ctypes.c_int64 * ndim)(*shape)
		istride = (ctypes.c_int64 * ndim)(*[elemstrides[a] for a in axes])
		check(lib.pz_permute(data.dtype.itemsize, out.ptr, data.ptr, ndim, oshape, istride, None))
		return out

	def moveaxis(self, data, src, dst, out=None, allocator=None):
		ndim = data.ndim
		if not (0 <= src < ndim and 0 <= dst < ndim):
			raise ValueError("axis is out of range")
		axes = [a for a in range(ndim) if a != src]
		axes.insert(dst, src)
		return self.transpose(data, tuple(axes), out=out, allocator=allocator)

	def swapaxes(self, data, axis1, axis2, out=None, allocator=None):
		ndim = data.ndim
		if not (0 <= axis1 < ndim and 0 <= axis2 < ndim):
			raise ValueError("axis is out of range")
		axes = list(range(ndim))
		axes[axis1], axes[axis2] = axes[axis2], axes[axis1]
		return self.transpose(data, tuple(axes), out=out, allocator=allocator)

	@staticmethod
	def _depthLayout(data):
		maps = [ary.shape[1] for ary in data]
		height, width = max(ary.shape[2] for ary in data), max(ary.shape[3] for ary in data)
		return maps, height, width

	def depthConcat(self, data, out=None, allocator=None):
		"""maps of different sizes stacked along the channel axis, each centred in the largest plane, zero elsewhere"""
		for ary in data:
			self._check4d(ary, "data")
		batchsize, dtype = data[0].shape[0], data[0].dtype
		if any(ary.shape[0] != batchsize or ary.dtype != dtype for ary in data):
			raise ValueError("tensors must share batch size and datatype")

		maps, height, width = self._depthLayout(data)
		shape = (batchsize, sum(maps), height, width)
		if out is None:
			out = GPUArray.zeros(shape, dtype, allocator=allocator)
		else:
			_checkOut(out, shape, dtype)
			out.fill(0)

		c0 = 0
		for ary in data:
			_, c, h, w = ary.shape
			oh, ow = (height - h) // 2, (width - w) // 2
			if ary.size:
				ary.copy(out=out[:, c0:c0 + c, oh:oh + h, ow:ow + w])
			c0 += c
		return out

	def depthSplit(self, grad, indata, allocator=None):
		self._check4d(grad, "grad")
		maps, height, width = self._depthLayout(indata)
		if grad.shape != (indata[0].shape[0], sum(maps), height, width):
			raise ValueError("grad has invalid shape %s" % (grad.shape, ))

		ingrads, c0 = [], 0
		for ary in indata:
			_, c, h, w = ary.shape
			oh, ow = (height - h) // 2, (width - w) // 2
			ingrads.append(grad[:, c0:c0 + c, oh:oh + h, ow:ow + w].copy(allocator=allocator))
			c0 += c
		return ingrads

	# ------------------------------------------------------------------------------------------ local response normalisation
	# (reference: Backend/Dnn.py:93-111, Cuda/Source/Libs/CuDnnNorm.c:329-690)
	def _lrn(self, mode, data, N, alpha, beta, K, out, allocator):
		self._check4d(data, "data")
		data.enforceContiguous()
		out = GPUArray(data.shape, data.dtype, allocator=allocator) if out is None else _checkOut(out, data.shape, data.dtype)
		check(lib.pz_lrn_fwd(dtypeCode(data.dtype), mode, data.ptr, out.ptr, *data.shape, int(N), float(alpha), float(beta), float(K), None))
		return out

	def _lrnBackward(self, mode, data, grad, N, alpha, beta, K, out, allocator):
		self._check4d(data, "data")
		if grad.shape != data.shape or grad.dtype != data.dtype:
			raise ValueError("invalid grad gpuarray data layout")
		out = GPUArray(data.shape, data.dtype, allocator=allocator) if out is None else _checkOut(out, data.shape, data.dtype)
		tmp = GPUArray(data.shape, _f32, allocator=allocator)
		check(lib.pz_lrn_bwd(dtypeCode(data.dtype), mode, data.ptr, grad.ptr, out.ptr, tmp.ptr, *data.shape, int(N), float(alpha),
							 float(beta), float(K), None))
		return out

	def crossMapLRN(self, data, N=5, alpha=1e-4, beta=0.75, K=2.0, out=None, allocator=None):
		return self._lrn(0, data, N, alpha, beta, K, out, allocator)

	def crossMapLRNBackward(self, data, outdata, grad, N=5, alpha=1e-4, beta=0.75, K=2.0, out=None, allocator=None):
		return self._lrnBackward(0, data, grad, N, alpha, beta, K, out, allocator)

	def mapLRN(self, data, means=None, N=5, alpha=1e-4, beta=0.75, K=2.0, out=None, allocator=None):
		"""reference: CuDnn_Context_pyMapLRN, CuDnnNorm.c:329-420 -- within-map LRN; with `means` the divisive normalisation the
		LCN module uses (Modules/LCN.py:33-39)"""
		if means is None:
			return self._lrn(1, data, N, alpha, beta, K, out, allocator)
		self._check4d(data, "data")
		if means.shape != data.shape or means.dtype != data.dtype:
			raise ValueError("invalid means gpuarray data layout")
		out = GPUArray(data.shape, data.dtype, allocator=allocator) if out is None else _checkOut(out, data.shape, data.dtype)
		check(lib.pz_divnorm_fwd(dtypeCode(data.dtype), data.ptr, means.ptr, out.ptr, data.shape[0] * data.shape[1], data.shape[2],
								 data.shape[3], int(N), float(alpha), float(beta), float(K), None))
		return out

	def mapLRNBackward(self, data, grad, means=None, N=5, alpha=1e-4, beta=0.75, K=2.0, dmeans=None, out=None, allocator=None):
		"""reference: CuDnn_Context_pyMapLRNBackward, CuDnnNorm.c:456-527 -- returns the input gradient, or (input gradient, means
		gradient) when a means tensor is given"""
		if means is None:
			return self._lrnBackward(1, data, grad, N, alpha, beta, K, out, allocator)
		self._check4d(data, "data")
		if grad.shape != data.shape or grad.dtype != data.dtype:
			raise ValueError("invalid grad gpuarray data layout")
		if means.shape != data.shape or means.dtype != data.dtype:
			raise ValueError("invalid means gpuarray data layout")
		out = GPUArray(data.shape, data.dtype, allocator=allocator) if out is None else _checkOut(out, data.shape, data.dtype)
		dmeans = GPUArray(data.shape, data.dtype, allocator=allocator) if dmeans is None else _checkOut(dmeans, data.shape, data.dtype)
		tmp = GPUArray(data.shape, _f32, allocator=allocator)
		check(lib.pz_divnorm_bwd(dtypeCode(data.dtype), data.ptr, means.ptr, grad.ptr, out.ptr, dmeans.ptr, tmp.ptr,
								 data.shape[0] * data.shape[1], data.shape[2], data.shape[3], int(N), float(alpha), float(beta), float(K), None))
		return out, dmeans

	# ------------------------------------------------------------------------------------------ spatial transformer
	def spatialTf(self, data, transform, outshape=None, getGrid=False, grid=None, out=None, allocator=None):
		"""reference: CuDnn_Context_pySpatialTf, CuDnnSpatialTf.c:66-150 -- affine grid + bilinear sampler"""
		self._check4d(data, "data")
		outshape = tuple(data.shape) if outshape is None else tuple(int(v) for v in outshape)
		if len(outshape) != 4 or outshape[:2] != tuple(data.shape[:2]):
			raise ValueError("invalid outshape")
		if transform.shape != (data.shape[0], 2, 3) or transform.dtype != data.dtype:
			raise ValueError("invalid transform gpuarray data layout")
		gridshape = (data.shape[0], outshape[2], outshape[3], 2)
		grid = GPUArray(gridshape, data.dtype, allocator=allocator) if grid is None else _checkOut(grid, gridshape, data.dtype)
		out = GPUArray(outshape, data.dtype, allocator=allocator) if out is None else _checkOut(out, outshape, data.dtype)
		check(lib.pz_spatialtf_fwd(dtypeCode(data.dtype), data.ptr, transform.ptr, grid.ptr, out.ptr, data.shape[0], data.shape[1],
								   data.shape[2], data.shape[3], outshape[2], outshape[3], None))
		return (out, grid) if getGrid else out

	def spatialTfBackward(self, grad, indata, grid, getDGrid=False, dgrid=None, dtransform=None, out=None, allocator=None):
		"""reference: CuDnn_Context_pySpatialTfBackward, CuDnnSpatialTf.c:229-285 -> (ingrad, dtransform[, dgrid])"""
		self._check4d(grad, "grad")
		self._check4d(indata, "indata")
		B, C, H, W = indata.shape
		if grad.shape[:2] != (B, C) or grad.dtype != indata.dtype or grid.shape != (B, grad.shape[2], grad.shape[3], 2):
			raise ValueError("invalid grad / grid gpuarray data layout")
		out = GPUArray(indata.shape, indata.dtype, allocator=allocator) if out is None else _checkOut(out, indata.shape, indata.dtype)
		dgrid = GPUArray(grid.shape, indata.dtype, allocator=allocator) if dgrid is None else _checkOut(dgrid, grid.shape, indata.dtype)
		dtransform = GPUArray((B, 2, 3), indata.dtype, allocator=allocator) if dtransform is None else \
			_checkOut(dtransform, (B, 2, 3), indata.dtype)
		check(lib.pz_spatialtf_bwd(dtypeCode(indata.dtype), grad.ptr, indata.ptr, grid.ptr, out.ptr, dtransform.ptr, dgrid.ptr, B, C, H, W,
								   grad.shape[2], grad.shape[3], None))
		return (out, dtransform, dgrid) if getDGrid else (out, dtransform)

	# ------------------------------------------------------------------------------------------ batch norm
	@staticmethod
	def _bnGeometry(data, mode):
		if data.ndim < 2:
			raise ValueError("invalid data gpuarray dims")
		if mode == 0:   # per activation: one statistic per (c, spatial...) position over the batch
			return data.shape[0], prod(data.shape[1:]), 1
		return data.shape[0], data.shape[1], prod(data.shape[2:])

	@staticmethod
	def _checkBnParam(ary, C, name):
		_requireArray(ary, name)
		if ary.dtype != _f32 or ary.size != C:
			raise ValueError("invalid %s gpuarray data layout" % name)

	def batchNormNd(self, data, mean, var, scale, bias, epsilon=1e-5, factor=1.0, test=False, mode=1, out=None,
					allocator=None):
		"""reference: CuDnn_Context_pyBatchNormNd, CuDnnNorm.c:80-155; params and statistics are always fp32 (:118-121)"""
		_requireArray(data, "data")
		N, C, S = self._bnGeometry(data, mode)
		for ary, name in ((mean, "mean"), (var, "var"), (scale, "scale"), (bias, "bias")):
			self._checkBnParam(ary, C, name)

		out = GPUArray(data.shape, data.dtype, allocator=allocator) if out is None else _checkOut(out, data.shape, data.dtype)
		code = dtypeCode(data.dtype)

		if test:
			check(lib.pz_bn_fwd_infer(code, data.ptr, out.ptr, N, C, S, scale.ptr, bias.ptr, mean.ptr, var.ptr, epsilon, None))
			return out

		driver.traceScalar("the batch-norm running-average factor", factor)
		savemean = GPUArray(scale.shape, _f32, allocator=allocator)
		saveinvvar = GPUArray(scale.shape, _f32, allocator=allocator)
		params = (scale, bias, mean, var, savemean, saveinvvar)
		if data.nbytes >= gpuarray.DEFER_MIN_BYTES and data.contiguous and out.contiguous and all(p.contiguous for p in params):
			# held back for one call: a ReLU of `out` that follows is folded into the same pass (gpuarray.DeferredBatchNorm)
			driver.flushDeferred()
			driver.deferred = gpuarray.DeferredBatchNorm(code, data, out, (N, C, S), params, epsilon, factor)
			return out, savemean, saveinvvar
		check(lib.pz_bn_fwd_train(code, data.ptr, out.ptr, N, C, S, scale.ptr, bias.ptr, mean.ptr, var.ptr, savemean.ptr,
								  saveinvvar.ptr, epsilon, factor, None))
		return out, savemean, saveinvvar

	def batchNormNdBackward(self, grad, data, scale, savemean=None, saveinvvar=None, epsilon=1e-5, mode=1,
							scalegrad=None, bgrad=None, out=None, allocator=None):
		"""reference: CuDnn_Context_pyBatchNormNdBackward, CuDnnNorm.c:201-293"""
		_requireArray(grad, "grad")
		_requireArray(data, "data")
		if grad.shape != data.shape or grad.dtype != data.dtype:
			raise ValueError("invalid grad gpuarray data layout")
		N, C, S = self._bnGeometry(data, mode)
		self._checkBnParam(scale, C, "scale")

		if savemean is None or saveinvvar is None:
			# cuDNN recomputes the batch statistics when no saved ones are given (CuDnnNorm.c:176-190)
			tmpmean, tmpvar = GPUArray.zeros(scale.shape, _f32, allocator=allocator), GPUArray.zeros(scale.shape, _f32, allocator=allocator)
			tmpbias = GPUArray.zeros(scale.shape, _f32, allocator=allocator)
			_, savemean, saveinvvar = self.batchNormNd(data, tmpmean, tmpvar, scale, tmpbias, epsilon, 1.0, False, mode,
													   allocator=allocator)
		else:
			self._checkBnParam(savemean, C, "savemean")
			self._checkBnParam(saveinvvar, C, "saveinvvar")

		out = GPUArray(data.shape, data.dtype, allocator=allocator) if out is None else _checkOut(out, data.shape, data.dtype)
		scalegrad = GPUArray(scale.shape, _f32, allocator=allocator) if scalegrad is None else _checkOut(scalegrad, scale.shape, _f32)
		bgrad = GPUArray(scale.shape, _f32, allocator=allocator) if bgrad is None else _checkOut(bgrad, scale.shape, _f32)

		tensors = (data, grad, out, scale, savemean, saveinvvar, scalegrad, bgrad)
		if data.nbytes >= gpuarray.DEFER_MIN_BYTES and all(t.contiguous for t in tensors):
			# held back: the module's two parameter-gradient accumulations that follow ride on the pass (gpuarray.DeferredBatchNormBackward)
			driver.flushDeferred()
			driver.deferred = gpuarray.DeferredBatchNormBackward(dtypeCode(data.dtype), data, grad, out, (N, C, S), (scale, savemean, saveinvvar),
																 scalegrad, bgrad)
			return out, scalegrad, bgrad
		check(lib.pz_bn_bwd(dtypeCode(data.dtype), data.ptr, grad.ptr, out.ptr, N, C, S, scale.ptr, savemean.ptr, saveinvvar.ptr,
							scalegrad.ptr, bgrad.ptr, None))
		return out, scalegrad, bgrad


# ============================================================================================================ blas
class BlasContext:
	def __init__(self, backend):
		self.backend = backend
		self.tensorOps = True

	def enableTensorOps(self, enable):
		"""reference: CuBlas_Context_enableTensorOps (CuBlas.c:91-106).  False selects the exact fp32 mode, see
		DnnContext.enableTensorOps (one process-wide switch)"""
		self.backend.dnn.enableTensorOps(enable)
		return self

	def gemm(self, A, B, out=None, transpA=False, transpB=False, alpha=1.0, beta=0.0, allocator=None):
		"""Row-major out = alpha * op(A) op(B) + beta * out; at most one operand transposed (reference:
		CuBlas_Context_gemm, CuBlas.c:327-403, shape rules :168-203)"""
		_requireArray(A, "A")
		_requireArray(B, "B")
		if A.ndim != 2 or B.ndim != 2 or A.dtype != B.dtype:
			raise ValueError("invalid gemm operands")
		if transpA and transpB:
			raise ValueError("only one of the gemm operands can be transposed")

		M, K = (A.shape[1], A.shape[0]) if transpA else A.shape
		Kb, N = (B.shape[1], B.shape[0]) if transpB else B.shape
		if K != Kb:
			raise ValueError("gemm operand shapes %s and %s do not match" % (A.shape, B.shape))

		if out is None:
			out = GPUArray((M, N), A.dtype, allocator=allocator)
			if beta != 0.0:
				out.fill(0)
		else:
			_checkOut(out, (M, N), A.dtype)
			if driver.gradientWriteHook is not None:
				driver.gradientWriteHook(out)

		check(lib.pz_gemm(dtypeCode(A.dtype), A.ptr, B.ptr, out.ptr, M, N, K, A.shape[1], B.shape[1], N, int(transpA),
						  int(transpB), alpha, beta, None, None))
		return out

	# ---- level-1 reductions returning host scalars (reference: Cuda/Source/Libs/CuBlas.c dot / l1norm / l2norm)
	def _reduce(self, kind, x, y=None):
		_requireArray(x, "x")
		x.enforceContiguous()
		if y is not None:
			_requireArray(y, "y")
			y.enforceContiguous()
			if y.size != x.size or y.dtype != x.dtype:
				raise ValueError("vectors must share size and datatype")
		out = GPUArray.zeros((), _f32, allocator=self.backend.memoryPool)
		check(lib.pz_vec_reduce(dtypeCode(x.dtype), kind, x.ptr, None if y is None else y.ptr, x.size, out.ptr, None))
		return float(out.get())

	def dot(self, x, y):
		return self._reduce(0, x, y)

	def l1norm(self, x):
		return self._reduce(1, x)

	def l2norm(self, x):
		return float(np.sqrt(self._reduce(2, x)))

	def gemmBatched(self, A, B, formatA=0, formatB=0, formatOut=0, transpA=False, transpB=False, alpha=1.0, beta=0.0, out=None,
					allocator=None):
		"""One GEMM per group over 3-d tensors in (group, batch, param) = gbp (0) or (batch, group, param) = bgp (1) layout
		(reference: CuBlas_Context_gemmBatched, CuBlas.c:207-320; tests Cuda/Wrappers/CuBlas.py:60-189).  A bgp operand is a
		pitched matrix per group, so every group is one pz_gemm call with the right leading dimension -- no repacking."""
		_requireArray(A, "A")
		_requireArray(B, "B")
		if A.ndim != 3 or B.ndim != 3 or A.dtype != B.dtype:
			raise ValueError("invalid gemmBatched operands")
		if transpA and transpB:
			raise ValueError("only one of the gemm operands can be transposed")

		def geometry(ary, fmt):
			# -> groups, rows, cols, leading dimension, element offset between groups
			if fmt == 0:
				g, r, c = ary.shape
				return g, r, c, c, r * c
			r, g, c = ary.shape
			return g, r, c, g * c, c

		ga, ra, ca, lda, sa = geometry(A, formatA)
		gb, rb, cb, ldb, sb = geometry(B, formatB)
		if ga != gb and ga != 1 and gb != 1:
			raise ValueError("invalid input gpuarray dims")
		# a single-group operand is shared by every group of the other one (CuBlas.c:255-262,304-305: stride 0)
		if ga == 1 and gb > 1:
			sa = 0
		elif gb == 1 and ga > 1:
			sb = 0
		ga = max(ga, gb)
		M, K = (ca, ra) if transpA else (ra, ca)
		Kb, N = (cb, rb) if transpB else (rb, cb)
		if K != Kb:
			raise ValueError("gemmBatched operand shapes %s and %s do not match" % (A.shape, B.shape))

		outshape = (ga, M, N) if formatOut == 0 else (M, ga, N)
		if out is None:
			out = GPUArray(outshape, A.dtype, allocator=allocator)
			if beta != 0.0:
				out.fill(0)
		else:
			_checkOut(out, outshape, A.dtype)
		_, _, _, ldc, sc = geometry(out, formatOut)

		code, es = dtypeCode(A.dtype), A.dtype.itemsize
		for g in range(ga):
			check(lib.pz_gemm(code, A.ptr + g * sa * es, B.ptr + g * sb * es, out.ptr + g * sc * es, M, N, K, lda, ldb, ldc,
							  int(transpA), int(transpB), alpha, beta, None, None))
		return out

	def gemmBias(self, A, B, bias, out=None, transpB=False, allocator=None):
		"""Linear forward with the bias add folded into the GEMM epilogue (Linear.py:36-40 as one kernel).  An extension: the reference's
		Linear calls gemm + addVecToMat; the bias must have exactly one entry per output column."""
		for ary, name in ((A, "A"), (B, "B"), (bias, "bias")):
			_requireArray(ary, name)
		if A.ndim != 2 or B.ndim != 2 or B.dtype != A.dtype or bias.dtype != A.dtype:
			raise ValueError("gemmBias needs 2-d operands and a bias of one dtype")
		M, K = A.shape
		N = B.shape[0] if transpB else B.shape[1]
		if (B.shape[1] if transpB else B.shape[0]) != K or bias.size != N:
			raise ValueError("gemmBias: shapes %s x %s with a bias of %d entries do not fit" % (A.shape, B.shape, bias.size))
		out = GPUArray((M, N), A.dtype, allocator=allocator) if out is None else _checkOut(out, (M, N), A.dtype)
		check(lib.pz_gemm(dtypeCode(A.dtype), A.ptr, B.ptr, out.ptr, M, N, K, A.shape[1], B.shape[1], N, 0, int(transpB),
						  1.0, 0.0, bias.ptr, None))
		return out


# ============================================================================================================ matmod
class MatModule:
	"""reference: Cuda/Kernels/MatVec.py:216-374"""
	GPUArray = GPUArray

	def __init__(self, backend):
		self.backend = backend

	def matvec(self, mat, vec, axis=0, out=None, alpha=1.0, beta=0.0, allocator=None):
		"""reference: Cuda/Kernels/MatVec.py:311-343 -- mat (..., h, w); axis 1: out (..., h) = mat . vec(..., w);
		axis 0: out (..., w) = mat^T . vec(..., h); out = beta * out + alpha * product"""
		_requireArray(mat, "mat")
		_requireArray(vec, "vec")
		if vec.dtype != mat.dtype or vec.ndim != mat.ndim - 1 or not 0 <= axis < 2 or mat.ndim < 2:
			raise ValueError("invalid matrix / vector layout")
		mat.enforceContiguous()
		vec.enforceContiguous()
		h, w = mat.shape[-2:]
		if vec.shape != mat.shape[:-2] + ((w, ) if axis == 1 else (h, )):
			raise ValueError("vector shape %s does not fit matrix shape %s along axis %d" % (vec.shape, mat.shape, axis))
		shape = mat.shape[:-2] + ((h, ) if axis == 1 else (w, ))
		if out is None:
			out = GPUArray.zeros(shape, mat.dtype, allocator=allocator)
		else:
			_checkOut(out, shape, mat.dtype)
		check(lib.pz_matvec(dtypeCode(mat.dtype), out.ptr, mat.ptr, vec.ptr, prod(mat.shape[:-2]), h, w, 1 if axis == 1 else 0,
							float(alpha), float(beta), None))
		return out

	def addVecToMat(self, vec, mat, axis=0, out=None, allocator=None):
		assert vec.dtype == mat.dtype
		assert vec.ndim == mat.ndim - 1 and 0 <= axis < 2
		assert mat.shape[:-2] == vec.shape[:-1]

		out = GPUArray(mat.shape, mat.dtype, allocator=allocator) if out is None else out
		z = prod(mat.shape[:-2])
		n, m = mat.shape[-2:]

		if axis == 1:
			assert m % vec.shape[-1] == 0
		else:
			assert vec.shape[-1] == n

		check(lib.pz_addvec2mat(dtypeCode(mat.dtype), out.ptr, mat.ptr, vec.ptr, z, n, m, axis, vec.shape[-1], None))
		return out

	def matsum(self, tensor, axis=0, out=None, alpha=1.0, beta=0.0, allocator=None):
		assert 0 <= axis < tensor.ndim
		outshape = tensor.shape[:axis] + tensor.shape[axis + 1:]

		if out is None:
			out = GPUArray.zeros(outshape, tensor.dtype, allocator=allocator)
		else:
			assert out.shape == outshape
			if driver.gradientWriteHook is not None:
				driver.gradientWriteHook(out)

		if axis == tensor.ndim - 1:
			z, h, w, rows = 1, prod(tensor.shape[:-1]), tensor.shape[-1], 1
		else:
			z, h, w, rows = prod(tensor.shape[:axis]), tensor.shape[axis], prod(tensor.shape[axis + 1:]), 0

		check(lib.pz_matsum(dtypeCode(tensor.dtype), out.ptr, tensor.ptr, z, h, w, rows, alpha, beta, None))
		return out

	def argminmax(self, tensor, axis, mode, allocator=None):
		assert 0 <= axis < tensor.ndim
		idx = GPUArray(tensor.shape[:axis] + tensor.shape[axis + 1:], _i32, allocator=allocator)
		z, h, w = prod(tensor.shape[:axis]), tensor.shape[axis], prod(tensor.shape[axis + 1:])
		check(lib.pz_argminmax(dtypeCode(tensor.dtype), idx.ptr, tensor.ptr, z, h, w, 1 if mode == "max" else 0, None))
		return idx

	def argmax(self, tensor, axis=0, allocator=None):
		return self.argminmax(tensor, axis, "max", allocator)

	def argmin(self, tensor, axis=0, allocator=None):
		return self.argminmax(tensor, axis, "min", allocator)


# ============================================================================================================ poolmod
class PoolModule:
	"""reference: Cuda/Kernels/Pool.py:115-213 (bit-exact int32 argmax masks)"""
	GPUArray = GPUArray

	def __init__(self, backend):
		self.backend = backend

	def maxpool2d(self, data, size, stride, pad, allocator=None):
		assert data.dtype == _f32
		batchsize, maps, inh, inw = data.shape
		(fh, fw), (hstride, wstride), (hpad, wpad) = size, stride, pad

		outh = (inh - fh + 2 * hpad) // hstride + 1
		outw = (inw - fw + 2 * wpad) // wstride + 1

		outdata = GPUArray((batchsize, maps, outh, outw), _f32, allocator=allocator)
		mask = GPUArray((batchsize, maps, outh, outw), _i32, allocator=allocator)

		check(lib.pz_maxpool2d_mask_fwd(data.ptr, outdata.ptr, mask.ptr, batchsize * maps, inh, inw, outh, outw, fh, fw,
										hstride, wstride, hpad, wpad, None))
		return outdata, mask

	def maxpool2dBackward(self, grad, origshape, mask, size, stride, pad, allocator=None):
		assert grad.dtype == _f32 and mask.dtype == _i32
		batchsize, maps, outh, outw = grad.shape
		(fh, fw), (hstride, wstride), (hpad, wpad) = size, stride, pad
		inh, inw = origshape[2], origshape[3]

		ingrad = GPUArray((batchsize, maps, inh, inw), _f32, allocator=allocator)
		check(lib.pz_maxpool2d_mask_bwd(grad.ptr, mask.ptr, ingrad.ptr, batchsize * maps, inh, inw, outh, outw, fh, fw,
										hstride, wstride, hpad, wpad, None))
		return ingrad

	def maxunpool2d(self, data, origshape, mask, allocator=None):
		assert data.dtype == _f32
		batchsize, maps, inh, inw = data.shape
		outh, outw = origshape[2], origshape[3]

		outdata = GPUArray.zeros((batchsize, maps, outh, outw), _f32, allocator=allocator)
		check(lib.pz_maxunpool2d_fwd(data.ptr, mask.ptr, outdata.ptr, batchsize * maps, inh * inw, outh * outw, None))
		return outdata

	def maxunpool2dBackward(self, grad, poolshape, mask, allocator=None):
		assert grad.dtype == _f32 and mask.dtype == _i32
		batchsize, maps, outh, outw = grad.shape
		inh, inw = poolshape[2], poolshape[3]

		ingrad = GPUArray((batchsize, maps, inh, inw), _f32, allocator=allocator)
		check(lib.pz_maxunpool2d_bwd(grad.ptr, mask.ptr, ingrad.ptr, batchsize * maps, inh * inw, outh * outw, None))
		return ingrad


# ============================================================================================================ SharedArray
# ============================================================================================================ costmod
class PReluModule:
	"""reference: Cuda/Kernels/PRelu.py:60-132 (float32 only; one slope per map, or one shared by all maps)"""
	GPUArray = GPUArray

	def __init__(self, backend):
		self.backend = backend

	@staticmethod
	def _geometry(data, slopes, sharedMaps):
		assert slopes.shape == (1, ) if sharedMaps else data.shape[1] == slopes.shape[0]
		return data.shape[0], data.shape[1], prod(data.shape[2:])

	def prelu(self, data, slopes, inplace=False, sharedMaps=False, allocator=None):
		assert data.dtype == slopes.dtype and slopes.dtype == _f32
		N, C, S = self._geometry(data, slopes, sharedMaps)
		outdata = data if inplace else GPUArray(data.shape, _f32, allocator=allocator)
		check(lib.pz_prelu_fwd(data.ptr, slopes.ptr, outdata.ptr, N, C, S, 1 if sharedMaps else 0, None))
		return outdata

	def preluBackwardData(self, grad, slopes, indata, sharedMaps=False, allocator=None):
		assert grad.dtype == slopes.dtype and slopes.dtype == indata.dtype and indata.dtype == _f32
		assert grad.shape == indata.shape
		N, C, S = self._geometry(grad, slopes, sharedMaps)
		ingrad = GPUArray(grad.shape, _f32, allocator=allocator)
		check(lib.pz_prelu_bwd_data(grad.ptr, slopes.ptr, indata.ptr, ingrad.ptr, N, C, S, 1 if sharedMaps else 0, None))
		return ingrad

	def preluBackwardParams(self, indata, outgrad, sharedMaps=False, allocator=None):
		assert indata.dtype == outgrad.dtype and outgrad.dtype == _f32
		assert indata.shape == outgrad.shape
		N, C, S = indata.shape[0], indata.shape[1], prod(indata.shape[2:])
		slopegrad = GPUArray((1, ) if sharedMaps else (C, ), _f32, allocator=allocator)
		check(lib.pz_prelu_bwd_params(indata.ptr, outgrad.ptr, slopegrad.ptr, N, C, S, 1 if sharedMaps else 0, None))
		return slopegrad


class PadModule:
	"""reference: Cuda/Kernels/Pad.py:145-229 (reflection padding of 3-d / 4-d tensors, float32 / float16)"""
	GPUArray = GPUArray

	def __init__(self, backend):
		self.backend = backend

	@staticmethod
	def _geometry(shape, pad, grad):
		if len(shape) == 3:
			lpad, rpad = pad
			upad = bpad = 0
			batchsize, maps, h, w = shape[0], shape[1], 1, shape[2]
		elif len(shape) == 4:
			upad, bpad, lpad, rpad = pad
			batchsize, maps, h, w = shape
		else:
			raise NotImplementedError(len(shape))
		if grad:
			h, w = h - upad - bpad, w - lpad - rpad
		else:
			assert h >= max(upad, bpad) + 1 and w >= max(lpad, rpad) + 1
		return batchsize, maps, h, w, upad, bpad, lpad, rpad

	def reflectpad(self, data, pad, allocator=None):
		batchsize, maps, h, w, upad, bpad, lpad, rpad = self._geometry(data.shape, pad, False)
		outshape = (batchsize, maps, w + lpad + rpad) if data.ndim == 3 else (batchsize, maps, h + upad + bpad, w + lpad + rpad)
		outdata = GPUArray(outshape, data.dtype, allocator=allocator)
		check(lib.pz_reflectpad_fwd(dtypeCode(data.dtype), data.ptr, outdata.ptr, batchsize * maps, h, w, upad, bpad, lpad, rpad, None))
		return outdata

	def reflectpadBackward(self, grad, pad, allocator=None):
		batchsize, maps, h, w, upad, bpad, lpad, rpad = self._geometry(grad.shape, pad, True)
		inshape = (batchsize, maps, w) if grad.ndim == 3 else (batchsize, maps, h, w)
		ingrad = GPUArray(inshape, grad.dtype, allocator=allocator)
		check(lib.pz_reflectpad_bwd(dtypeCode(grad.dtype), grad.ptr, ingrad.ptr, batchsize * maps, h, w, upad, bpad, lpad, rpad, None))
		return ingrad


class EmbedModule:
	"""reference: Cuda/Kernels/Embedder.py:45-87 (int32 word indices, -1 = padding; float32 / float16 vocabulary)"""
	GPUArray = GPUArray

	def __init__(self, backend):
		self.backend = backend

	def embed(self, data, W, allocator=None):
		assert data.dtype == _i32 and (W.dtype == _f32 or W.dtype == np.float16)
		batchsize, sentlen = data.shape
		_, embsize = W.shape
		outdata = GPUArray((batchsize, sentlen, embsize), W.dtype, allocator=allocator)
		check(lib.pz_embed_fwd(dtypeCode(W.dtype), data.ptr, W.ptr, outdata.ptr, batchsize * sentlen, embsize, None))
		return outdata

	def embedBackwardParams(self, indata, grad, W, scale):
		assert indata.shape == grad.shape[:2] and W.shape[1] == grad.shape[2]
		assert indata.dtype == _i32 and grad.dtype == W.dtype
		batchsize, sentlen = indata.shape
		if driver.gradientWriteHook is not None:
			driver.gradientWriteHook(W)
		check(lib.pz_embed_bwd(dtypeCode(W.dtype), indata.ptr, grad.ptr, W.ptr, float(scale), batchsize * sentlen, W.shape[1], None))


class UpsampleModule:
	"""reference: Cuda/Kernels/Upsample.py:300-454 (float32; integer scale factors; "nearest" and align-corners "linear")"""
	GPUArray = GPUArray

	def __init__(self, backend):
		self.backend = backend

	@staticmethod
	def _scale(scale, n):
		return (scale, ) * n if isinstance(scale, int) else tuple(scale)

	def _forward(self, data, dims, scales, mode, allocator):
		batchsize, maps = data.shape[:2]
		outdims = tuple(s * d for s, d in zip(scales, dims))
		outdata = GPUArray((batchsize, maps) + outdims[3 - (data.ndim - 2):], data.dtype, allocator=allocator)
		if mode == "nearest":
			check(lib.pz_upsample_nearest_fwd(data.ptr, outdata.ptr, batchsize * maps, *dims, *scales, None))
		elif mode == "linear":
			ratios = [(i - 1) / (o - 1) if data.ndim == 5 or k > 0 else 0.0 for k, (i, o) in enumerate(zip(dims, outdims))]
			check(lib.pz_upsample_linear_fwd(data.ptr, outdata.ptr, batchsize * maps, *dims, *outdims, *ratios, 1 if data.ndim == 5 else 0, None))
		else:
			raise NotImplementedError(mode)
		return outdata

	def _backward(self, grad, outdims, scales, mode, allocator):
		batchsize, maps = grad.shape[:2]
		dims = tuple(o // s for s, o in zip(scales, outdims))
		inshape = (batchsize, maps) + dims[3 - (grad.ndim - 2):]
		if mode == "nearest":
			ingrad = GPUArray(inshape, grad.dtype, allocator=allocator)
			check(lib.pz_upsample_nearest_bwd(grad.ptr, ingrad.ptr, batchsize * maps, *dims, *scales, None))
		elif mode == "linear":
			ingrad = GPUArray.zeros(inshape, grad.dtype, allocator=allocator)
			ratios = [(i - 1) / (o - 1) if grad.ndim == 5 or k > 0 else 0.0 for k, (i, o) in enumerate(zip(dims, outdims))]
			check(lib.pz_upsample_linear_bwd(grad.ptr, ingrad.ptr, batchsize * maps, *dims, *outdims, *ratios, 1 if grad.ndim == 5 else 0, None))
		else:
			raise NotImplementedError(mode)
		return ingrad

	def upsample2d(self, data, scale, mode="nearest", allocator=None):
		assert data.dtype == _f32
		hscale, wscale = self._scale(scale, 2)
		return self._forward(data, (1, ) + tuple(data.shape[2:]), (1, hscale, wscale), mode, allocator)

	def upsample2dBackward(self, grad, scale, mode="nearest", allocator=None):
		assert grad.dtype == _f32
		hscale, wscale = self._scale(scale, 2)
		return self._backward(grad, (1, ) + tuple(grad.shape[2:]), (1, hscale, wscale), mode, allocator)

	def upsample3d(self, data, scale, mode="nearest", allocator=None):
		assert data.dtype == _f32
		return self._forward(data, tuple(data.shape[2:]), self._scale(scale, 3), mode, allocator)

	def upsample3dBackward(self, grad, scale, mode="nearest", allocator=None):
		assert grad.dtype == _f32
		return self._backward(grad, tuple(grad.shape[2:]), self._scale(scale, 3), mode, allocator)


class CTCModule:
	"""reference: Cuda/Kernels/CTC.py:195-269 (connectionist temporal classification loss and its gradient, float32)"""
	GPUArray = GPUArray

	def __init__(self, backend):
		self.backend, self.dnn = backend, backend.dnn

	def ctcLoss(self, data, datalen, labels, lengths, blank, error=None, normalized=False, returnAlphas=False, allocator=None):
		assert data.dtype == _f32 and datalen.dtype == _i32 and labels.dtype == _i32
		T, batchsize, vocabsize = data.shape
		lengths = np.asarray(lengths)

		if not normalized:
			data = self.backend.dnn.softmaxNd(data.reshape(T * batchsize, vocabsize, 1, 1), allocator=allocator).reshape(T, batchsize, vocabsize)

		extOffsets = np.zeros((batchsize + 1, ), dtype=np.int32)
		extOffsets[1:] = np.cumsum(lengths, dtype=np.int32)

		alphas = GPUArray((T * (2 * int(extOffsets[-1]) + batchsize), ), _f32, allocator=allocator)
		offsets = GPUArray.toGpu(extOffsets, allocator=allocator)
		nll = GPUArray((batchsize, ), _f32, allocator=allocator)
		error = GPUArray.zeros((), _f32, allocator=allocator) if error is None else error
		grad = GPUArray.zeros(data.shape, _f32, allocator=allocator)

		check(lib.pz_ctc_loss(data.ptr, datalen.ptr, labels.ptr, offsets.ptr, alphas.ptr, nll.ptr, error.ptr, grad.ptr, T, batchsize, vocabsize,
							  int(lengths.max()) if lengths.size else 0, int(blank), None))
		return (error, grad) if not returnAlphas else (error, grad, alphas)


class CostModule:
	"""reference: Cuda/Kernels/Costs.py:160-247 (the cross-entropy entry and the accuracy reduction of the training closure)"""
	GPUArray = GPUArray

	def __init__(self, backend):
		self.backend = backend

	def crossEntropy(self, scores, labels, weights=None, error=None, allocator=None):
		"""softmax over axis 1 + cost; returns (error scalar on the device, grad) -- Costs.py:213-247"""
		_requireArray(scores, "scores")
		_requireArray(labels, "labels")
		if scores.dtype != _f32 or labels.dtype != np.int32:
			raise ValueError("crossEntropy needs float32 scores and int32 labels")

		shape = scores.shape
		if scores.ndim < 4:
			scores = scores.reshape(*shape, *(1 for _ in range(4 - scores.ndim)))

		softmax = self.backend.dnn.softmaxNd(scores, mode=self.backend.SoftMaxMode.spatial.value, allocator=allocator)
		grad = GPUArray.empty(shape, _f32, allocator=allocator)
		if error is None:
			error = GPUArray.empty((), _f32, allocator=allocator)
		error.fill(0.0)

		samples, cases, spatial = scores.shape[0], scores.shape[1], prod(scores.shape[2:])
		if labels.size != samples * spatial:
			raise ValueError("labels must hold one class index per sample and position")
		if weights is not None and (weights.dtype != _f32 or weights.size != cases):
			raise ValueError("weights must be a float32 vector with one entry per class")

		check(lib.pz_cross_entropy(softmax.ptr, labels.ptr, None if weights is None else weights.ptr, samples, cases, spatial,
								   error.ptr, grad.ptr, None))
		return error, grad


	def svm(self, scores, labels, mode, error=None, allocator=None):
		"""reference: Cuda/Kernels/Costs.py:249-279 -- mode "l1" (hinge) or "l2" (squared hinge); -> (error, grad)"""
		_requireArray(scores, "scores")
		_requireArray(labels, "labels")
		if scores.dtype != _f32 or labels.dtype != np.int32:
			raise ValueError("svm needs float32 scores and int32 labels")
		if mode not in ("l1", "l2"):
			raise KeyError(mode)

		grad = GPUArray.empty(scores.shape, _f32, allocator=allocator)
		if error is None:
			error = GPUArray.empty((), _f32, allocator=allocator)
		error.fill(0.0)

		samples, cases, spatial = scores.shape[0], scores.shape[1], prod(scores.shape[2:])
		if labels.size != samples * spatial:
			raise ValueError("labels must hold one class index per sample and position")
		check(lib.pz_svm(1 if mode == "l2" else 0, scores.ptr, labels.ptr, samples, cases, spatial, error.ptr, grad.ptr, None))
		return error, grad

	def getAccuracyKernel(self, name):
		"""reference: Cuda/Kernels/Costs.py:172-210 -- reductions returning a float32 device scalar"""
		if name == "calcAccuracy":
			def ker(x, y, allocator=None):
				if x.dtype != np.int32 or y.dtype != np.int32 or x.size != y.size:
					raise ValueError("calcAccuracy needs two int32 tensors of one size")
				out = GPUArray.zeros((), _f32, allocator=allocator)
				check(lib.pz_count_mismatch(x.ptr, y.ptr, x.size, out.ptr, None))
				return out
			return ker

		kinds = {"calcBCEAccuracy": (0, _i32), "klDivergence": (1, _f32), "l1HingeAccuracy": (2, _i32)}
		if name not in kinds:
			raise NotImplementedError(name)
		kind, ytype = kinds[name]

		def ker(x, y, grad=None, gradnorm=0.0, allocator=None):
			if x.dtype != _f32 or y.dtype != ytype or x.size != y.size or (kind == 1) != (grad is not None):
				raise ValueError("%s: invalid arguments" % name)
			out = GPUArray.zeros((), _f32, allocator=allocator)
			check(lib.pz_cost_reduce(kind, x.ptr, y.ptr, grad.ptr if grad is not None else None, float(gradnorm), x.size, out.ptr, None))
			return out
		return ker


# ============================================================================================================ rng
class RandomNumberGenerator:
	"""reference: Cuda/Source/Libs/CuRand.c (the generator object behind `gpuarray.globalRng`): fillInteger / fillUniform /
	fillNormal of a whole gpuarray.  Philox4x32-10 on the device; (seed, offset) is the whole state.  The offset lives in
	DEVICE memory and is advanced by a kernel after every fill, so a captured graph (driver.StepGraph) that replays a fill
	draws new numbers every time."""

	def __init__(self, type=None, seed=0):
		self.type = "philox4x32-10" if type is None else type
		self.seed = int(seed) & 0xffffffffffffffff
		self.state = None            # device uint64: the offset, in units of 4 random words

	@property
	def offset(self):
		return 0 if self.state is None else int(self.state.get()[0])

	@offset.setter
	def offset(self, value):
		if self.state is None:
			self.state = GPUArray((1, ), np.uint64)
		self.state.set(np.array([int(value)], np.uint64))

	def _fill(self, kind, ary, a, b, dtype):
		_requireArray(ary, "ary")
		ary.enforceContiguous()
		if ary.dtype != dtype:
			raise ValueError("unsupported gpuarray dtype")
		if self.state is None:
			self.offset = 0
		check(lib.pz_rng_fill_dev(kind, ary.ptr, ary.size, self.seed, self.state.ptr, float(a), float(b), None))

	def fillInteger(self, ary):
		self._fill(0, ary, 0.0, 0.0, np.dtype(np.uint32) if ary.dtype == np.uint32 else np.dtype(np.int32))

	def fillUniform(self, ary, minval=0.0, maxval=1.0):
		self._fill(1, ary, minval, maxval, _f32)

	def fillNormal(self, ary, mean=0.0, stddev=1.0):
		self._fill(2, ary, mean, stddev, _f32)


class SharedArray:
	"""One flat buffer with 16-byte aligned named views (reference: Cuda/Utils.py:19-64); the per-dtype flat
	parameter / gradient buffers of Optimizer.setupGlobalState and the payload of the DP all-reduce."""
	alignment = 16

	def __init__(self, dtype=np.float32, allocator=None):
		self.ary = None
		self.blocks = OrderedDict()
		self.dtype = np.dtype(dtype)
		self.allocator = allocator

	def register(self, shape, dtype, name):
		assert name not in self.blocks
		assert dtype == self.dtype
		self.blocks[name] = (shape, prod(shape) * self.dtype.itemsize)

	def build(self):
		totalbytes = sum(self.align(nbytes) for _, nbytes in self.blocks.values())
		self.ary = GPUArray((totalbytes // self.dtype.itemsize, ), self.dtype, allocator=self.allocator)

		blocks, offset = OrderedDict(), 0
		for name, (shape, nbytes) in self.blocks.items():
			blocks[name] = GPUArray(shape, self.dtype, gpudata=self.ary.gpudata[offset:offset + nbytes])
			offset += self.align(nbytes)
		self.blocks = blocks

	def __getitem__(self, item):
		return self.blocks[item]

	@classmethod
	def align(cls, nbytes):
		return (nbytes + cls.alignment - 1) // cls.alignment * cls.alignment


class QueueManager:
	"""Borrow / give pool of streams or events (reference: Cuda/Utils.py:67-94)."""

	def __init__(self, objtype):
		self.objtype = objtype
		self.items = []

	def reserve(self, nitems):
		self.items.extend(self.objtype() for _ in range(nitems))

	def borrow(self, nitems):
		if len(self.items) < nitems:
			self.reserve(nitems - len(self.items))
		borrowed, self.items = self.items[:nitems], self.items[nitems:]
		return borrowed

	def give(self, items):
		self.items.extend(items)

	def clear(self):
		self.items = []


# ============================================================================================================ kernels
_ACT_KINDS = {"sigmoid": 0, "tanh": 1, "relu": 2, "leakyRelu": 3, "elu": 4, "softPlus": 5, "clip": 6, "gelu": 7}


def _noSlice(kwargs):
	if kwargs.get("slice") is not None:
		raise NotImplementedError("this kernel has no strided `slice=` form in the B200 backend")


def _slice(kwargs, size):
	"""`slice=` of an elementwise launch -> (start, stop, step) or None (reference: Cuda/SourceModule.py:162-173)"""
	slc = kwargs.get("slice")
	if slc is None:
		return None
	start = 0 if slc.start is None else int(slc.start)
	stop = size if slc.stop is None else int(slc.stop)
	step = 1 if slc.step is None else int(slc.step)
	if start < 0 or step < 1:
		raise ValueError("invalid elementwise slice %s" % (slc, ))
	return start, min(stop, size), step


def _actFactory(kind, nscalars):
	code = _ACT_KINDS[kind]

	def factory(dtype):
		dt = dtypeCode(dtype)

		def fwd(out, inp, *scalars, **kwargs):
			a = float(scalars[0]) if nscalars > 0 else 0.0
			b = float(scalars[1]) if nscalars > 1 else 0.0
			slc = _slice(kwargs, out.size)
			if slc is None:
				if kind == "relu" and driver.deferred is not None and (gpuarray.reluAfterSum(out, inp) or gpuarray.reluAfterBatchNorm(out, inp)):
					return          # fused with the pending Add / Replicate sum, or the pending batch-norm launch, whose output it reads
				check(lib.pz_act_fwd(code, dt, out.ptr, inp.ptr, out.size, a, b, None))
			else:
				check(lib.pz_act_fwd_slice(code, dt, out.ptr, inp.ptr, out.size, a, b, *slc, None))

		return fwd

	def derFactory(dtype):
		dt = dtypeCode(dtype)

		def bwd(ingrad, outgrad, ref, *scalars, **kwargs):
			a = float(scalars[0]) if nscalars > 0 else 0.0
			b = float(scalars[1]) if nscalars > 1 else 0.0
			slc = _slice(kwargs, ingrad.size)
			if slc is None:
				if kind == "relu" and driver.deferred is not None and gpuarray.reluDerAfterSum(ingrad, outgrad, ref):
					return          # fused with the pending Replicate / Add gradient sum it reads
				check(lib.pz_act_bwd(code, dt, ingrad.ptr, outgrad.ptr, ref.ptr, ingrad.size, a, b, None))
			else:
				check(lib.pz_act_bwd_slice(code, dt, ingrad.ptr, outgrad.ptr, ref.ptr, ingrad.size, a, b, *slc, None))

		return bwd

	return factory, derFactory


_EW_OPS = {
	"absKer": 0, "weightDecayKer": 1, "l1penaltyKer": 2, "l1gradKer": 3, "rbmKer": 4, "rmspropKer": 5, "rmspropGravesKer": 6,
	"adagradKer": 7, "adadeltaKer": 8, "smorms3Ker": 9, "bceKer": 10, "hingeKer": 11, "smoothL1Ker": 12, "l1HingeKer": 13
}


def _genericKernel(name, nptrs, nscalars, naux=0, dtype=None, f32state=()):
	"""A kernel object with the reference's calling convention `(ptr args..., scalar args..., slice=None)` over pz_eltwise
	(reference objects: Cuda/Kernels/ElementWise.py, Costs.py:8-74).  `dtype` None -> taken from the first array."""
	op = _EW_OPS[name]

	def ker(*args, **kwargs):
		if len(args) != nptrs + nscalars + naux:
			raise TypeError("%s takes %d arguments (%d given)" % (name, nptrs + nscalars + naux, len(args)))
		arrays, rest = args[:nptrs], args[nptrs:]
		dt = arrays[0].dtype if dtype is None else dtype
		for i, ary in enumerate(arrays):
			_requireArray(ary, "argument #%d" % (i + 1))
			want = _f32 if i in f32state else (dt if ary.dtype != _i32 else _i32)
			if ary.dtype != want:
				raise ValueError("%s: argument #%d has dtype %s" % (name, i + 1, ary.dtype))

		# the launch size is the size of the first array argument (Cuda/SourceModule.py:128-137)
		size = arrays[0].size
		start, stop, step = _slice(kwargs, size) or (0, size, 1)

		ptrs = (ctypes.c_void_p * nptrs)(*[ary.ptr for ary in arrays])
		if nscalars and op in (5, 6, 7, 8, 9):
			driver.traceScalar("the %s rates" % name, *rest)
		if naux:       # pointwise costs: scalars are (int, int) or (float, float) after the pointers
			scalars = (ctypes.c_float * 4)()
			aux = (ctypes.c_int * 2)(int(rest[0]), int(rest[1]))
			check(lib.pz_eltwise(op, dtypeCode(dt), ptrs, nptrs, scalars, 0, aux, size, start, stop, step, None))
		else:
			scalars = (ctypes.c_float * max(1, nscalars))(*[float(v) for v in rest])
			check(lib.pz_eltwise(op, dtypeCode(dt), ptrs, nptrs, scalars, nscalars, None, size, start, stop, step, None))

	ker.__name__ = name
	return ker


def _genericFactory(name, nptrs, nscalars, f32state=()):
	"""`fooKer(dtype)` -> kernel, like the reference's memoized factories"""
	cache = {}

	def factory(dtype):
		dtype = np.dtype(dtype)
		if dtype not in cache:
			dtypeCode(dtype)
			cache[dtype] = _genericKernel(name, nptrs, nscalars, f32state=f32state)
		return cache[dtype]

	factory.__name__ = name
	return factory


class ConvPerf:
	def __init__(self, algo, tm, memory):
		self.algo, self.time, self.memory = algo, tm, memory
		self.determinism = True

	def __repr__(self):
		return "Algo %s time %.6f secs memory %.6f mbytes" % (self.algo, self.time, self.memory / 1024 ** 2)


class NotBehindTheSeam:
	"""An attribute of the reference backend object that this backend does not implement: binding it (the reference's
	Backend/Kernels/*.py do so at import time) works, USING it raises -- loudly, never a CPU or library fallback."""

	def __init__(self, name):
		self._name = name

	def __call__(self, *args, **kwargs):
		raise NotImplementedError("%s is not implemented in the B200 backend" % self._name)

	def __getattr__(self, attr):
		if attr.startswith("__"):
			raise AttributeError(attr)
		return NotBehindTheSeam("%s.%s" % (self._name, attr))

	def __repr__(self):
		return "<%s: not implemented in the B200 backend>" % self._name


# ============================================================================================================ backend
class B200Backend:
	BackendName = "B200"
	warpSize, nthreads = 32, 1024

	GPUArray = GPUArray
	Driver = driver
	SharedArray = SharedArray

	class GroupFormat(Enum):
		gbp = 0
		bgp = 1

	class ConvFwdAlgo(Enum):
		implicitGemm = 0
		implicitPrecompGemm = 1
		gemm = 2
		direct = 3
		fft = 4
		fftTiling = 5
		winograd = 6
		winogradNonfused = 7

	class ConvBwdDataAlgo(Enum):
		algo0 = 0
		algo1 = 1
		fft = 2
		fftTiling = 3
		winograd = 4
		winogradNonfused = 5

	class ConvBwdFilterAlgo(Enum):
		algo0 = 0
		algo1 = 1
		fft = 2
		algo3 = 3
		winograd = 4
		winogradNonfused = 5
		fftTiling = 6

	class PoolMode(Enum):
		max = 0
		avgWithPad = 1
		avgNoPad = 2
		maxDeterminism = 3

	class SoftMaxMode(Enum):
		perActivation = 0
		spatial = 1

	class BatchNormMode(Enum):
		perActivation = 0
		spatial = 1
		spatialPersistent = 2

	class RNNAlgo(Enum):
		standard = 0
		persistStatic = 1
		persistDynamic = 2

	class RNNMode(Enum):
		relu = 0
		tanh = 1
		lstm = 2
		gru = 3

	class DirectionMode(Enum):
		uni = 0
		bi = 1

	notImplemented = ()

	def __init__(self, deviceIdx, initmode=0, logger=None):
		self.deviceIdx = deviceIdx
		self.logger = logger

		ndevices = driver.Device.count()
		if ndevices == 0:
			raise driver.CudaError("no CUDA device is visible")

		self.device = driver.Device(deviceIdx % ndevices).set()
		major, minor = self.device.computeCapability()
		if major != 10:
			raise driver.CudaError(
				"libpzb200 holds sm_100a code only; device %s is sm_%d%d" % (self.device.name(), major, minor)
			)

		if logger is not None:
			logger.debug("Using device #%s (%s), %d SMs", deviceIdx, self.device.name(), driver.Device.smCount())

		self.memoryPool = driver.MemoryPool()
		self.globalRng = RandomNumberGenerator(seed=int(np.random.randint(np.iinfo(np.int64).max, dtype=np.int64)))
		self.streamManager = QueueManager(driver.Stream)
		self.eventManager = QueueManager(driver.Event)

		self.initmode = 0
		self.blas, self.dnn = None, None
		self.matmod, self.poolmod, self.costmod = None, None, None
		self.prelumod, self.padmod, self.embedmod, self.upsamplemod, self.ctcmod = None, None, None, None, None

		# attributes the reference's Backend/Kernels/*.py bind at import time (Cuda/GPUBackend.py:74-131) that sit outside the
		# hot path and are not implemented: binding works, use raises NotImplementedError
		for name in self.notImplemented:
			if getattr(self, name, None) is None:
				setattr(self, name, NotBehindTheSeam(name))

		self.updateBackend(initmode)

	def updateBackend(self, initmode):
		if initmode >= 1 and self.dnn is None:
			self.blas, self.dnn = BlasContext(self), DnnContext(self)
			self.memmod = self.dnn          # depthConcat / depthSplit / moveaxis / swapaxes / transpose (Cuda/Kernels/Memory.py's surface)
		if initmode >= 2 and self.matmod is None:
			self.matmod, self.poolmod, self.costmod = MatModule(self), PoolModule(self), CostModule(self)
			self.prelumod, self.padmod, self.embedmod, self.upsamplemod = PReluModule(self), PadModule(self), EmbedModule(self), UpsampleModule(self)
			self.ctcmod = CTCModule(self)
		self.initmode = max(self.initmode, initmode)

	# ---- kernel factories (reference attribute names: Cuda/GPUBackend.py:85-131)
	sigmoidKer, sigmoidDerKer = (staticmethod(f) for f in _actFactory("sigmoid", 0))
	tanhKer, tanhDerKer = (staticmethod(f) for f in _actFactory("tanh", 0))
	reluKer, reluDerKer = (staticmethod(f) for f in _actFactory("relu", 0))
	leakyReluKer, leakyReluDerKer = (staticmethod(f) for f in _actFactory("leakyRelu", 1))
	eluKer, eluDerKer = (staticmethod(f) for f in _actFactory("elu", 1))
	softPlusKer, softPlusDerKer = (staticmethod(f) for f in _actFactory("softPlus", 0))
	clipKer, clipDerKer = (staticmethod(f) for f in _actFactory("clip", 2))
	geluKer, geluDerKer = (staticmethod(f) for f in _actFactory("gelu", 0))

	@staticmethod
	def toVectorAddVectorKer(dtype):
		dt = dtypeCode(dtype)

		def ker(y, x, alpha, **kwargs):
			_noSlice(kwargs)
			if driver.gradientWriteHook is not None:
				driver.gradientWriteHook(y)
			if driver.deferred is not None and gpuarray.accumulate(y, x, alpha):
				return          # absorbed by / fused with the pending zero fill of y (Add.py:15-23, Replicate.py:18-29)
			check(lib.pz_axpy(dt, y.ptr, x.ptr, float(alpha), y.size, None))

		return ker

	@staticmethod
	def addKer(dtype):
		dt = dtypeCode(dtype)

		def ker(out, x, alpha, y, beta, **kwargs):
			slc = _slice(kwargs, out.size)
			if driver.gradientWriteHook is not None:
				driver.gradientWriteHook(out)
			if slc is None:
				if driver.deferred is not None and gpuarray.accumulateAfterBatchNormBackward(out, x, alpha, y, beta):
					return      # rides on the pending batch-norm backward pass (BatchNormND.py:74-83)
				check(lib.pz_axpby(dt, out.ptr, x.ptr, float(alpha), y.ptr, float(beta), out.size, None))
			else:
				check(lib.pz_axpby_slice(dt, out.ptr, x.ptr, float(alpha), y.ptr, float(beta), out.size, *slc, None))

		return ker

	@staticmethod
	def mulKer(dtype):
		dt = dtypeCode(dtype)

		def ker(out, a, b, **kwargs):
			slc = _slice(kwargs, out.size)
			if slc is None:
				check(lib.pz_mul(dt, out.ptr, a.ptr, b.ptr, out.size, None))
			else:
				check(lib.pz_mul_slice(dt, out.ptr, a.ptr, b.ptr, out.size, *slc, None))

		return ker

	@staticmethod
	def linearKer(dtype):
		dt = dtypeCode(dtype)

		def ker(out, inp, a, b, **kwargs):
			_noSlice(kwargs)
			check(lib.pz_scale_shift(dt, out.ptr, inp.ptr, float(a), float(b), out.size, None))

		return ker

	@staticmethod
	def classicMomSGDKer(dtype):
		dt = dtypeCode(dtype)

		def ker(param, grad, mom, learnRate, momRate, **kwargs):
			driver.traceScalar("the momentum-SGD rates", learnRate, momRate)
			check(lib.pz_sgd_momentum(dt, param.ptr, grad.ptr, mom.ptr, float(learnRate), float(momRate), param.size, None))

		return ker

	@staticmethod
	def nesterovMomSGDKer(dtype):
		dt = dtypeCode(dtype)

		def ker(param, grad, mom, learnRate, momRate, **kwargs):
			driver.traceScalar("the Nesterov-SGD rates", learnRate, momRate)
			check(lib.pz_sgd_nesterov(dt, param.ptr, grad.ptr, mom.ptr, float(learnRate), float(momRate), param.size, None))

		return ker

	@staticmethod
	def adamKer(dtype):
		dt = dtypeCode(dtype)

		def ker(param, grad, mg, ms, learnRate, fix1, fix2, epsilon, **kwargs):
			if mg.dtype != _f32 or ms.dtype != _f32:
				raise ValueError("adam moments must be float32")
			driver.traceScalar("Adam's bias-corrected rates", learnRate, fix1, fix2)
			check(lib.pz_adam(dt, param.ptr, grad.ptr, mg.ptr, ms.ptr, float(learnRate), float(fix1), float(fix2), float(epsilon),
							  param.size, None))

		return ker

	# reference: Cuda/Kernels/ElementWise.py:1100-1138 (kernel objects, float32 only) and :614-706,860-1003 (factories)
	absKer = staticmethod(_genericKernel("absKer", 2, 0, dtype=_f32))
	weightDecayKer = staticmethod(_genericKernel("weightDecayKer", 2, 1, dtype=_f32))
	l1penaltyKer = staticmethod(_genericKernel("l1penaltyKer", 3, 1, dtype=_f32))
	l1gradKer = staticmethod(_genericKernel("l1gradKer", 3, 1, dtype=_f32))
	rbmKer = staticmethod(_genericKernel("rbmKer", 3, 0, dtype=_f32))

	rmspropKer = staticmethod(_genericFactory("rmspropKer", 3, 3))
	rmspropGravesKer = staticmethod(_genericFactory("rmspropGravesKer", 5, 4))
	adagradKer = staticmethod(_genericFactory("adagradKer", 3, 2))
	adadeltaKer = staticmethod(_genericFactory("adadeltaKer", 4, 2))
	smorms3Ker = staticmethod(_genericFactory("smorms3Ker", 5, 2, f32state=(2, 3, 4)))

	# reference: Cuda/Kernels/Costs.py:8-74 -- (scores / pred, labels / target, totalError, grad..., two ints or two floats)
	bceKer = staticmethod(_genericKernel("bceKer", 4, 0, naux=2, dtype=_f32))
	hingeKer = staticmethod(_genericKernel("hingeKer", 4, 0, naux=2, dtype=_f32))
	smoothL1Ker = staticmethod(_genericKernel("smoothL1Ker", 4, 2, dtype=_f32))
	l1HingeKer = staticmethod(_genericKernel("l1HingeKer", 6, 0, naux=2, dtype=_f32))

	@staticmethod
	def castFP16toFP32(outdata, indata, **kwargs):
		"""reference: Cuda/Kernels/ElementWise.py castFP16toFP32 -- a kernel object, not a factory"""
		_noSlice(kwargs)
		if outdata.dtype != _f32 or indata.dtype != np.float16 or outdata.size != indata.size:
			raise ValueError("castFP16toFP32 needs a float32 output and a float16 input of one size")
		check(lib.pz_cast(driver.PZ_F32, outdata.ptr, driver.PZ_F16, indata.ptr, indata.size, None))

	@staticmethod
	def castFP32toFP16(outdata, indata, **kwargs):
		_noSlice(kwargs)
		if outdata.dtype != np.float16 or indata.dtype != _f32 or outdata.size != indata.size:
			raise ValueError("castFP32toFP16 needs a float16 output and a float32 input of one size")
		check(lib.pz_cast(driver.PZ_F16, outdata.ptr, driver.PZ_F32, indata.ptr, indata.size, None))

	def getAccuracyKernel(self, name):
		"""reference: Cuda/GPUBackend.py:159 -- bound from costmod (Cuda/Kernels/Costs.py:172-210)"""
		self.updateBackend(2)
		return self.costmod.getAccuracyKernel(name)

	@staticmethod
	def _dropoutKer(dtype, mapped):
		dt = dtypeCode(dtype)
		parttype = np.dtype(np.uint32) if np.dtype(dtype) == _f32 else np.dtype(np.uint16)

		def ker(outdata, indata, b, v, p, mapsize=1, **kwargs):
			if b.dtype != parttype or b.size * (mapsize if mapped else 1) < indata.size:
				raise ValueError("dropout needs one %s random word per %s" % (parttype, "map" if mapped else "element"))
			start, stop, step = _slice(kwargs, indata.size) or (0, indata.size, 1)
			check(lib.pz_dropout_slice(dt, outdata.ptr, indata.ptr, b.ptr, int(v), float(p), indata.size, int(mapsize) if mapped else 1,
									   start, stop, step, None))

		return ker

	@classmethod
	def dropoutKer(cls, dtype):
		"""reference: Cuda/Kernels/ElementWise.py:495-536 -- ker(outdata, indata, rands, partition, p)"""
		return cls._dropoutKer(dtype, False)

	@classmethod
	def dropout2dKer(cls, dtype):
		"""reference: ElementWise.py:539-580 -- ker(outdata, indata, rands, partition, p, mapsize): one word per feature map"""
		return cls._dropoutKer(dtype, True)

	@staticmethod
	def add2Ker(dtype):
		"""out = a + b in one pass: the value Add.updateData / Replicate.updateGrad build with fill(0) + 2 axpy
		(reference Modules/Add.py:15-23), 3 tensor passes instead of 5"""
		dt = dtypeCode(dtype)

		def ker(out, a, b):
			check(lib.pz_add2(dt, out.ptr, a.ptr, b.ptr, out.size, None))

		return ker

	# ---- misc surface used by Backend/gpuarray.py
	@staticmethod
	def dtypesSupported():
		"""the calculation types with the tolerance the reference's own tests hold them to (Cuda/GPUBackend.py:218-220);
		bfloat16 works everywhere float16 does but is not advertised here: the reference's Modules / tests iterate
		this list and key python dicts by it (Modules/Cast.py:76, Dropout.py:43)"""
		return [(np.float32, 1e-5), (np.float16, 1e-2)]

	@staticmethod
	def dtypesExtra():
		return [(driver.bfloat16, 5e-2)] if driver.bfloat16 is not None else []

	def fillUniform(self, data, minval=0.0, maxval=1.0, rng=None):
		"""reference: Cuda/GPUBackend.py:234-241"""
		if data.dtype != _f32:
			raise ValueError("fillUniform needs a float32 gpuarray")
		rng = self.globalRng if rng is None else rng
		rng.fillUniform(data, minval, maxval)

	def fillNormal(self, data, mean=0.0, stddev=1.0, rng=None):
		"""reference: Cuda/GPUBackend.py:244-246"""
		rng = self.globalRng if rng is None else rng
		rng.fillNormal(data, mean=mean, stddev=stddev)

	@staticmethod
	def copy(dest, source, allocator=None):
		if dest is None:
			return source.copy(allocator=allocator)
		dest.set(source)
		return dest

	def concatenate(self, tup, axis, out=None, allocator=None):
		ary = tup[0]
		reduced = ary.shape[:axis] + ary.shape[axis + 1:]
		assert all(a.dtype == ary.dtype and a.shape[:axis] + a.shape[axis + 1:] == reduced for a in tup[1:])

		shape = ary.shape[:axis] + (sum(a.shape[axis] for a in tup), ) + ary.shape[axis + 1:]
		if out is None:
			out = GPUArray(shape, ary.dtype, allocator=allocator)
		else:
			assert out.shape == shape and out.dtype == ary.dtype

		dstPitch = out.strides[axis - 1] if axis > 0 else out.nbytes
		height, offset = prod(shape[:axis]), 0

		for a in tup:
			width = a.strides[axis - 1] if axis > 0 else a.nbytes
			check(lib.pz_memcpy2d(out.ptr + offset, dstPitch, a.ptr, width, width, height, 0, None))
			offset += width
		return out

	def split(self, ary, sections, axis, allocator=None):
		assert sum(sections) == ary.shape[axis]
		outs = [GPUArray(ary.shape[:axis] + (sec, ) + ary.shape[axis + 1:], ary.dtype, allocator=allocator) for sec in sections]

		srcPitch = ary.strides[axis - 1] if axis > 0 else ary.nbytes
		height, offset = prod(ary.shape[:axis]), 0

		for out in outs:
			width = out.strides[axis - 1] if axis > 0 else out.nbytes
			check(lib.pz_memcpy2d(out.ptr, width, ary.ptr + offset, srcPitch, width, height, 0, None))
			offset += width
		return outs

	def tile(self, ary, times, axis, allocator=None):
		return self.concatenate([ary] * times, axis, allocator=allocator)

	def timeKernel(self, func, args, kwargs=None, looplength=1000, log=True, logname=None, normalize=False, hotpass=True):
		"""CUDA-event timing of a python-side launch loop (reference: Cuda/GPUBackend.py:332-368)"""
		kwargs = {} if kwargs is None else kwargs
		if hotpass:
			func(*args, **kwargs)

		start, end = driver.Event(), driver.Event()
		start.record()
		for _ in range(looplength):
			func(*args, **kwargs)
		end.record()
		end.synchronize()

		secs = start.timeTill(end) * 1e-3
		if normalize:
			secs /= looplength
		if log and self.logger is not None:
			self.logger.info("%s time: %s secs", logname or func.__name__, secs)
		return secs

	def convNdbenchmark(self, datashape, Wshape, dtype, stride=1, pad=0, dilation=1, groups=1, algoCount=10):
		"""One implementation per pass, so each list holds one timed entry (reference: GPUBackend.py:371-378)."""
		stride, pad, dilation = _seq(stride, 2, 1, "stride"), _seq(pad, 2, 0, "pad"), _seq(dilation, 2, 1, "dilation")
		data, W = GPUArray.zeros(datashape, dtype, allocator=self.memoryPool), GPUArray.zeros(Wshape, dtype, allocator=self.memoryPool)
		out = self.dnn.convNd(data, W, None, stride, pad, dilation, groups, allocator=self.memoryPool)
		wgrad = GPUArray.zeros(Wshape, dtype, allocator=self.memoryPool)

		def bench(fn):
			return self.timeKernel(fn, (), looplength=5, log=False, normalize=True)

		fwd = bench(lambda: self.dnn.convNd(data, W, None, stride, pad, dilation, groups, out=out))
		bwdData = bench(lambda: self.dnn.convNdBackwardData(out, W, None, data, stride, pad, dilation, None, groups,
															allocator=self.memoryPool))
		bwdParam = bench(lambda: self.dnn.convNdBackwardParams(data, out, W, stride, pad, dilation, groups, wgrad=wgrad))
		return [ConvPerf(0, fwd, 0)], [ConvPerf(0, bwdData, 0)], [ConvPerf(0, bwdParam, 0)]

	# ---- recurrent layers (reference: Cuda/Backend.py:171-350)
	def createRnn(self, insize, hsize, dtype, layers=1, algo=None, mode=None, direction=None, dropout=0.0, seed=0, batchsize=0):
		from .rnn import Rnn
		algo = self.RNNAlgo.standard if algo is None else algo
		mode = self.RNNMode.lstm if mode is None else mode
		direction = self.DirectionMode.uni if direction is None else direction

		self.updateBackend(2)
		rnn = Rnn(self, insize, hsize, np.dtype(dtype), layers, algo.value, mode.value, direction.value, dropout, seed, batchsize)
		W = GPUArray.empty((rnn.wsize, ), dtype=dtype)
		return rnn, W, self.acquireRnnParams(rnn, W)

	def acquireRnnParams(self, rnn, W):
		return rnn.acquireParams(W)

	def updateRnnParams(self, rnn, W, params):
		# the named parameters ARE views into W here (no cuDNN-owned copy to refresh, Cuda/Backend.py:320-350)
		pass

	def deviceSupportsBatchHint(self):
		return self.device.computeCapability() >= (6, 1)

	def instanceNorm2d(self, data, scale, bias, epsilon=1e-5, out=None, allocator=None):
		"""BN over a (1, N*C, H, W) view with the affine parameters tiled N times (reference: GPUBackend.py:381-398)"""
		batchsize, maps, height, width = data.shape
		extmaps = batchsize * maps

		indata = data.reshape(1, extmaps, height, width)
		mean, var = GPUArray.zeros((extmaps, ), _f32, allocator=allocator), GPUArray.zeros((extmaps, ), _f32, allocator=allocator)

		if batchsize > 1:
			scale, bias = self.tile(scale, batchsize, axis=0, allocator=allocator), self.tile(bias, batchsize, axis=0, allocator=allocator)

		outdata, savemean, saveinvvar = self.dnn.batchNormNd(indata, mean, var, scale, bias, epsilon, 1.0, False, 1, out=out,
															 allocator=allocator)
		return outdata.reshape(data.shape), savemean, saveinvvar, scale

	def instanceNorm2dBackward(self, grad, data, extscale, savemean, saveinvvar, epsilon, affine=True, out=None, allocator=None):
		"""reference: GPUBackend.py:401-416"""
		batchsize, maps, height, width = grad.shape
		extmaps = batchsize * maps

		outgrad, scalegrad, biasgrad = self.dnn.batchNormNdBackward(
			grad.reshape(1, extmaps, height, width), data.reshape(1, extmaps, height, width), extscale, savemean, saveinvvar,
			epsilon, 1, out=out, allocator=allocator
		)
		outgrad = outgrad.reshape(grad.shape)

		if not affine:
			return outgrad

		if batchsize > 1:
			scalegrad = self.matmod.matsum(scalegrad.reshape(batchsize, -1), axis=0, allocator=allocator)
			biasgrad = self.matmod.matsum(biasgrad.reshape(batchsize, -1), axis=0, allocator=allocator)
		return outgrad, scalegrad, biasgrad


_backends = {}


def getDeviceCount():
	return driver.Device.count()


def getBackend(deviceIdx=0, initmode=0, logger=None):
	"""reference: Cuda/Backend.py:360-370 -- one cached backend object per device, upgraded in place by initmode"""
	bnd = _backends.get(deviceIdx)
	if bnd is None:
		bnd = B200Backend(deviceIdx, initmode, logger)
		_backends[deviceIdx] = bnd
	else:
		bnd.updateBackend(initmode)
	return bnd
