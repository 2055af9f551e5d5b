"""Late-bound function tables over the backend object -- the host-side mirror of the reference's `Backend/*` shim
(reference: Backend/gpuarray.py:60-113, Backend/Dnn.py:159-288, Backend/Blas.py:43-102, Backend/Kernels/*.py).

The reference binds module-level names at import time; here the tables are classes whose attributes resolve the
backend on first use, so importing the package never touches the GPU (the CPU-only test tier imports it freely).
Function names, argument order and defaults are the reference's.
"""
import numpy as np

from . import Config
from .backend import getBackend, B200Backend
from .gpuarray import GPUArray

_state = {"backend": None}


def backend():
	bnd = _state["backend"]
	if bnd is None:
		bnd = getBackend(Config.deviceIdx, initmode=2, logger=Config.getLogger() if Config.systemLog else None)
		_state["backend"] = bnd
	return bnd


def memoryPool():
	return backend().memoryPool


class gpuarray:
	"""reference: Backend/gpuarray.py"""
	GPUArray = GPUArray

	@staticmethod
	def to_gpu(ary, allocator=None):
		return GPUArray.toGpu(ary, allocator=allocator)

	@staticmethod
	def empty(shape, dtype, allocator=None):
		return GPUArray.empty(shape, dtype, allocator=allocator)

	@staticmethod
	def zeros(shape, dtype, allocator=None):
		return GPUArray.zeros(shape, dtype, allocator=allocator)

	@staticmethod
	def dtypesSupported():
		return B200Backend.dtypesSupported()

	@staticmethod
	def copy(dest, source):
		return backend().copy(dest, source, allocator=memoryPool())

	@staticmethod
	def concatenate(tup, axis, out=None):
		return backend().concatenate(tup, axis, out, allocator=memoryPool())

	@staticmethod
	def split(ary, sections, axis):
		return backend().split(ary, sections, axis, allocator=memoryPool())

	@staticmethod
	def tile(ary, times, axis):
		return backend().tile(ary, times, axis, allocator=memoryPool())

	@staticmethod
	def timeKernel(func, args, kwargs=None, looplength=1000, log=True, logname=None, normalize=False, hotpass=True):
		return backend().timeKernel(func, args, kwargs, looplength, log, logname, normalize, hotpass)


PoolMode = B200Backend.PoolMode
BatchNormMode = B200Backend.BatchNormMode
SoftMaxMode = B200Backend.SoftMaxMode
ConvFwdAlgo = B200Backend.ConvFwdAlgo
ConvBwdDataAlgo = B200Backend.ConvBwdDataAlgo
ConvBwdFilterAlgo = B200Backend.ConvBwdFilterAlgo


class Dnn:
	"""reference: Backend/Dnn.py:169-288 (the wrap* closures of initBaseGPU / initCuda)"""

	@staticmethod
	def convNd(data, W, bias=None, stride=1, pad=0, dilation=1, groups=1, algo=ConvFwdAlgo.implicitGemm):
		return backend().dnn.convNd(
			data, W, bias.ravel() if bias is not None else None, stride, pad, dilation, groups, algo.value, None, memoryPool()
		)

	@staticmethod
	def convNdBackwardData(grad, W, data=None, stride=1, pad=0, dilation=1, groups=1, algo=ConvBwdDataAlgo.algo0):
		return backend().dnn.convNdBackwardData(
			grad, W, None, data, stride, pad, dilation, None, groups, algo.value, None, memoryPool()
		)

	@staticmethod
	def convNdBackwardParams(data, grad, W, bias=None, stride=1, pad=0, dilation=1, groups=1, wgrad=None, bgrad=None,
							 scale=1.0, momentum=0.0, algo=ConvBwdFilterAlgo.algo0):
		return backend().dnn.convNdBackwardParams(
			data, grad, W, stride, pad, dilation, groups, bias is not None, False, wgrad,
			bgrad.ravel() if bgrad is not None else None, scale, momentum, algo.value, memoryPool()
		)

	@staticmethod
	def deconvNd(data, W, bias=None, stride=1, pad=0, dilation=1, postpad=0, groups=1, algo=ConvBwdDataAlgo.algo0):
		return backend().dnn.convNdBackwardData(
			data, W, bias.ravel() if bias is not None else None, None, stride, pad, dilation, postpad, groups, algo.value,
			None, memoryPool()
		)

	@staticmethod
	def deconvNdBackwardData(grad, W, data=None, stride=1, pad=0, dilation=1, groups=1, algo=ConvFwdAlgo.implicitGemm):
		assert data is not None
		return backend().dnn.convNd(grad, W, None, stride, pad, dilation, groups, algo.value, None, memoryPool())

	@staticmethod
	def deconvNdBackwardParams(data, grad, W, bias=None, stride=1, pad=0, dilation=1, groups=1, wgrad=None, bgrad=None,
							   scale=1.0, momentum=0.0, algo=ConvBwdFilterAlgo.algo0):
		return backend().dnn.convNdBackwardParams(
			grad, data, W, stride, pad, dilation, groups, bias is not None, True, wgrad,
			bgrad.ravel() if bgrad is not None else None, scale, momentum, algo.value, memoryPool()
		)

	@staticmethod
	def convNdbenchmark(datashape, Wshape, stride=1, pad=0, dilation=1, groups=1, transpose=False):
		fwd, bwdData, bwdParam = backend().convNdbenchmark(datashape, Wshape, np.float32, stride, pad, dilation, groups)
		return fwd, bwdParam, bwdData

	@staticmethod
	def poolNd(data, size=2, stride=2, pad=0, mode=PoolMode.max, test=False):
		return backend().dnn.poolNd(data, size, stride, pad, mode.value, None, memoryPool()), None

	@staticmethod
	def poolNdBackward(indata, outdata, grad, workspace, size=2, stride=2, pad=0, mode=PoolMode.max):
		return backend().dnn.poolNdBackward(grad, indata, outdata, size, stride, pad, mode.value, None, memoryPool())

	@staticmethod
	def batchNormNd(data, scale, bias, mean, var, epsilon=1e-5, factor=1.0, test=False, mode=BatchNormMode.spatial, out=None):
		shape = scale.shape
		result = backend().dnn.batchNormNd(
			data, mean.ravel(), var.ravel(), scale.ravel(), bias.ravel(), epsilon, factor, test, mode.value, out=out,
			allocator=memoryPool()
		)
		if test:
			return result

		outdata, savemean, saveinvvar = result
		return outdata, savemean.reshape(shape), saveinvvar.reshape(shape)

	@staticmethod
	def batchNormNdBackward(data, grad, scale, savemean, saveinvvar, epsilon=1e-5, mode=BatchNormMode.spatial):
		shape = scale.shape
		ingrad, scalegrad, bgrad = backend().dnn.batchNormNdBackward(
			grad, data, scale.ravel(), savemean.ravel(), saveinvvar.ravel(), epsilon, mode.value, allocator=memoryPool()
		)
		return ingrad, scalegrad.reshape(shape), bgrad.reshape(shape)

	@staticmethod
	def softmaxNd(data, mode=SoftMaxMode.spatial):
		return backend().dnn.softmaxNd(data, mode.value, allocator=memoryPool())

	@staticmethod
	def softmaxNdBackward(outdata, grad):
		return backend().dnn.softmaxNdBackward(grad, outdata, allocator=memoryPool())

	@staticmethod
	def instanceNorm2d(data, scale, bias, epsilon=1e-5):
		return backend().instanceNorm2d(data, scale.ravel(), bias.ravel(), epsilon, allocator=memoryPool())

	@staticmethod
	def instanceNorm2dBackward(grad, data, extscale, savemean, saveinvvar, epsilon, affine=True):
		return backend().instanceNorm2dBackward(grad, data, extscale, savemean, saveinvvar, epsilon, affine, allocator=memoryPool())


class Rnn:
	"""reference: Backend/Dnn.py:299-333 (initRnnGPU)"""

	@staticmethod
	def createRnn(insize, hsize, layers, mode, direction, dropout, seed, batchsize):
		rnn, W, params = backend().createRnn(insize, hsize, np.float32, layers, mode=mode, direction=direction, dropout=dropout,
											 seed=seed, batchsize=0 if batchsize is None else batchsize)
		return rnn, W, {i: layer for i, layer in enumerate(params)}

	@staticmethod
	def acquireRnnParams(descRnn, w):
		return w, {i: layer for i, layer in enumerate(backend().acquireRnnParams(descRnn, w))}

	@staticmethod
	def updateRnnParams(descRnn, w, params):
		backend().updateRnnParams(descRnn, w, [params[layer] for layer in sorted(params.keys())])

	@staticmethod
	def forwardRnn(data, W, descRnn, test=False):
		return descRnn.forward(data, W, test=test, allocator=memoryPool())

	@staticmethod
	def backwardDataRnn(grad, outdata, W, reserve, descRnn):
		ingrad, _, _ = descRnn.backwardData(grad, outdata, W, reserve, allocator=memoryPool())
		return ingrad, reserve

	@staticmethod
	def backwardParamsRnn(data, outdata, _, reserve, descRnn):
		return descRnn.backwardParams(data, outdata, reserve, allocator=memoryPool())


class Blas:
	"""reference: Backend/Blas.py:43-70"""

	@staticmethod
	def toVectorAddVector(y, x, alpha=1.0):
		backend().toVectorAddVectorKer(y.dtype)(y, x, alpha)
		return y

	@staticmethod
	def addVectorToVector(x, y, out=None, alpha=1.0, beta=1.0):
		if out is None:
			out = GPUArray.empty(x.shape, x.dtype, allocator=memoryPool())
		else:
			assert out.shape == x.shape
		backend().addKer(out.dtype)(out, x, alpha, y, beta)
		return out

	@staticmethod
	def mulMatrixOnMatrix(A, B, out=None, transpA=False, transpB=False, alpha=1.0, beta=0.0):
		return backend().blas.gemm(A, B, out, transpA, transpB, alpha, beta, memoryPool())

	@staticmethod
	def sumOnMatrix(A, out=None, cols=True, alpha=1.0, beta=0.0):
		assert A.ndim == 2
		return backend().matmod.matsum(A, 0 if cols else 1, out, alpha, beta, memoryPool())


class MatVec:
	"""reference: Backend/Kernels/MatVec.py:30-48"""

	@staticmethod
	def addVecToMat(vec, mat, axis=0, out=None):
		return backend().matmod.addVecToMat(vec, mat, axis, out, memoryPool())

	@staticmethod
	def argmax(tensor, axis=0):
		return backend().matmod.argmax(tensor, axis, memoryPool())


class Memory:
	"""reference: Backend/Memory.py:43-66"""

	@staticmethod
	def depthConcat(data):
		return backend().dnn.depthConcat(data, allocator=memoryPool())

	@staticmethod
	def depthSplit(grad, indata):
		return backend().dnn.depthSplit(grad, indata, allocator=memoryPool())

	@staticmethod
	def moveaxis(data, src, dst):
		return backend().dnn.moveaxis(data, src, dst, allocator=memoryPool())

	@staticmethod
	def swapaxes(data, axis1, axis2):
		return backend().dnn.swapaxes(data, axis1, axis2, allocator=memoryPool())

	@staticmethod
	def transpose(data, axes):
		return backend().dnn.transpose(data, axes, allocator=memoryPool())


class Pool:
	"""reference: Backend/Kernels/Pool.py:38-56"""

	@staticmethod
	def maxpool2d(data, size, stride, pad):
		return backend().poolmod.maxpool2d(data, size, stride, pad, memoryPool())

	@staticmethod
	def maxpool2dBackward(grad, origshape, mask, size, stride, pad):
		return backend().poolmod.maxpool2dBackward(grad, origshape, mask, size, stride, pad, memoryPool())

	@staticmethod
	def maxunpool2d(data, origshape, mask):
		return backend().poolmod.maxunpool2d(data, origshape, mask, memoryPool())

	@staticmethod
	def maxunpool2dBackward(grad, poolshape, mask):
		return backend().poolmod.maxunpool2dBackward(grad, poolshape, mask, memoryPool())
