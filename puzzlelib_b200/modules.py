"""Host-side mirror of the reference's operator API for the hot path: `Module` plus the operators the benchmark nets
dispatch (ConvND / DeconvND / Linear / BatchNormND / InstanceNorm2D / Pool / Activation / SoftMax and the glue modules
Add / Replicate / Identity / Flatten) and the containers Sequential / Parallel.

Class names, constructor arguments, attribute names (`vars`, `attrs`, `data`, `grad`, `inData`, `train`, `calctype`),
the `__call__` / `backward(grad, updParamGrads, updGrad, scale, momentum)` protocol and the error behaviour
(`ModuleError` on shape / dtype violations) are the reference's (Modules/Module.py:39-147, ConvND.py, DeconvND.py,
Linear.py, BatchNormND.py, InstanceNorm2D.py, MaxPool2D.py, AvgPool2D.py, Activation.py, SoftMax.py, Add.py,
Replicate.py, Flatten.py, Identity.py; Containers/Sequential.py:186-234, Parallel.py:96-150), so the parity tests read
like the reference's own unit tests.  HDF5 save/load, blueprints and the graph container are out of scope
(SURVEY section 8f).  Every operator body is one or two calls into the `shim` tables, i.e. into libpzb200.so.
"""
import math
from enum import Enum

import numpy as np

from . import Config
from .shim import gpuarray, Dnn, Blas, MatVec, Memory, Pool, PoolMode, ConvFwdAlgo, ConvBwdDataAlgo, ConvBwdFilterAlgo, memoryPool, \
	backend, Rnn
from .driver import bfloat16


class ModuleError(Exception):
	pass


class ContainerError(ModuleError):
	pass


# ---------------------------------------------------------------------------------------------------------- Variable
class Variable:
	"""reference: Variable.py:5-56"""
	index = 0

	def __init__(self, data, name=None, withgrad=True, grad=None, updater=None, postUpdater=None):
		if name is None:
			self.name = str(type(self).index)
			type(self).index += 1
		else:
			self.name = name

		self.data = data
		self.updater = updater

		if updater is not None:
			return

		self.postUpdater = postUpdater
		self.grad = None

		if grad is not None:
			self.grad = grad
		elif withgrad and not Config.globalEvalMode:
			self.grad = gpuarray.zeros(self.data.shape, dtype=self.data.dtype)

		self.learnRate, self.momRate = 1.0, 1.0
		self.wc = 0.0

	@property
	def hasUpdater(self):
		return self.updater is not None

	@property
	def hasPostUpdater(self):
		return self.postUpdater is not None

	def set(self, variable):
		self.data.set(variable.data)
		if self.grad is not None:
			self.grad.set(variable.grad)


# ---------------------------------------------------------------------------------------------------------- Module
class InitScheme(str, Enum):
	none = "none"
	xavier = "xavier"
	xavierUniform = "xavier_uniform"
	xavierNormal = "xavier_normal"
	he = "he"
	gaussian = "gaussian"
	uniform = "uniform"


class Module:
	def __init__(self, name=None):
		self.name = name

		self.vars = {}
		self.attrs = {}

		self.gradUsesOutData = False
		self.movesData = False
		self.movesGrad = False

		self.grad = None
		self.inData = None
		self.data = None

		self.train = False if Config.globalEvalMode else True
		self.calctype = np.float32

	def setVar(self, name, var):
		setattr(self, name, var.data)
		self.vars[name] = var

	def getVar(self, name):
		return self.vars[name]

	def setAttr(self, name, attr):
		setattr(self, name, attr)
		self.attrs[name] = attr

	def hasAttr(self, name):
		return name in self.attrs

	def getVarTable(self, vartable=None, name=None, root=True):
		if root and name is None:
			name = self.name if self.name is not None else ""

		vartable = {} if vartable is None else vartable

		for paramName, var in self.vars.items():
			vartable.setdefault(var, []).append("%s%s" % (name, paramName))
		return vartable

	def __call__(self, data):
		if not Config.disableDtypeShapeChecks:
			self.checkDataShape(self.acquireShapesFrom(data))
			self.checkDataType(self.acquireDtypesFrom(data))

		self.data = None
		self.inData = data

		self.updateData(data)
		return self.data

	def backward(self, grad, updParamGrads=True, updGrad=True, scale=1.0, momentum=0.0):
		if not Config.disableDtypeShapeChecks:
			self.checkGradShape(self.acquireShapesFrom(grad))
			self.checkGradType(self.acquireDtypesFrom(grad))

		self.grad = None

		if updGrad:
			self.updateGrad(grad)

		if updParamGrads and self.train:
			self.accGradParams(grad, scale=scale, momentum=momentum)

	def updateData(self, data):
		raise NotImplementedError()

	def updateGrad(self, grad):
		raise NotImplementedError()

	def zeroGradParams(self):
		for var in self.vars.values():
			if var.hasUpdater:
				continue
			var.grad.fill(0)

	def accGradParams(self, grad, scale=1.0, momentum=0.0):
		pass

	def updateParams(self, learnRate):
		for var in self.vars.values():
			Blas.toVectorAddVector(var.data.ravel(), var.grad.ravel(), alpha=learnRate)

	def optimizeForShape(self, shape, memlimit=None):
		pass

	def trainMode(self):
		self.train = True
		self.reset()

	def evalMode(self):
		self.train = False
		self.reset()

	def calcMode(self, T):
		if T != np.float32:
			raise ModuleError("Unsupported dtype %s" % T)
		self.calctype = T

	def reset(self):
		self.inData, self.data, self.grad = None, None, None

	def checkDataShape(self, shape):
		pass

	def dataShapeFrom(self, shape):
		raise NotImplementedError()

	def checkDataType(self, dtype):
		self.genericCheckDataType(dtype)

	def checkGradShape(self, shape):
		pass

	def gradShapeFrom(self, shape):
		raise NotImplementedError()

	def checkGradType(self, dtype):
		self.genericCheckDataType(dtype)

	def genericCheckDataType(self, dtype):
		if isinstance(dtype, (tuple, list)):
			for d in dtype:
				self.genericCheckDataType(d)
		elif dtype != self.calctype:
			raise ModuleError("Expected dtype %s, got %s" % (self.calctype, dtype))

	def __str__(self):
		return "Module %s (name: %s)" % (self.__class__.__name__, self.name)

	def numOfParams(self):
		return sum(var.data.size for var in self.vars.values())

	def paramSize(self):
		return sum(var.data.nbytes for var in self.vars.values())

	@staticmethod
	def repeat(val, ntimes):
		return (val, ) * ntimes if isinstance(val, int) else tuple(val)

	@classmethod
	def acquireShapesFrom(cls, data):
		return [cls.acquireShapesFrom(d) for d in data] if isinstance(data, (tuple, list)) else data.shape

	@classmethod
	def acquireDtypesFrom(cls, data):
		return [cls.acquireDtypesFrom(d) for d in data] if isinstance(data, (tuple, list)) else data.dtype

	@staticmethod
	def createTensorWithScheme(scheme, shape, wscale, factorShape=None, factorTranspose=False, dtype=np.float32):
		"""reference: Modules/Module.py:405-453 (same draws from numpy's global RNG for the same seed)"""
		factorType = "in"
		if isinstance(scheme, (tuple, list)):
			if len(scheme) != 2:
				raise ValueError("Scheme tuple has %s length, expected 2" % len(scheme))
			scheme, factorType = scheme

		scheme = InitScheme(scheme) if scheme is not None else scheme
		outs, ins = Module.inferNeuronsNumber(shape if factorShape is None else factorShape, factorTranspose)

		try:
			factor = {"avg": (outs + ins) / 2, "in": ins, "out": outs}[factorType]
		except KeyError:
			raise NotImplementedError(factorType)

		if scheme == InitScheme.none:
			return None
		elif scheme == InitScheme.xavierUniform or scheme is None:
			nwscale = math.sqrt(3.0 / factor)
			return np.random.uniform(-nwscale, nwscale, shape).astype(dtype)
		elif scheme == InitScheme.xavierNormal or scheme == InitScheme.xavier:
			return np.random.normal(0, math.sqrt(1.0 / factor), shape).astype(dtype)
		elif scheme == InitScheme.he:
			return np.random.normal(0.0, math.sqrt(2.0 / factor), shape).astype(dtype)
		elif scheme == InitScheme.gaussian:
			return np.random.normal(0.0, wscale, shape).astype(dtype)
		elif scheme == InitScheme.uniform:
			return np.random.uniform(-wscale, wscale, shape).astype(dtype)
		raise NotImplementedError(scheme.value)

	@staticmethod
	def inferNeuronsNumber(shape, transpose):
		ndim = len(shape)
		if ndim == 1:
			return shape[0], shape[0]
		elif ndim == 2:
			neuronsIn, neuronsOut = shape
		else:
			outmaps, inmaps = shape[:2]
			field = int(np.prod(shape[2:]))
			neuronsOut, neuronsIn = outmaps * field, inmaps * field
		return (neuronsIn, neuronsOut) if transpose else (neuronsOut, neuronsIn)

	def _recastVars(self, T):
		# weights and their grads are stored in the compute type, no fp32 master copy (reference: ConvND.py:106-122)
		if self.calctype == T:
			return
		variables, self.vars = self.vars, {}
		for varName, var in variables.items():
			self.setVar(varName, Variable(
				var.data.astype(T), name=var.name, grad=var.grad.astype(T) if var.grad is not None else None
			))
		self.calctype = T


def _floatTypes():
	return {np.dtype(dtype) for dtype, _ in gpuarray.dtypesSupported()}


# ---------------------------------------------------------------------------------------------------------- conv
class ConvND(Module):
	def __init__(self, nd, inmaps, outmaps, size, stride=1, pad=0, dilation=1, wscale=1.0, useBias=True, name=None,
				 initscheme=None, empty=False, groups=1):
		super().__init__(name)

		self.stride = self.repeat(stride, nd)
		self.pad = self.repeat(pad, nd)
		self.dilation = self.repeat(dilation, nd)

		self.useBias = useBias
		self.groups = groups

		self.fwdAlgo, self.bwdFilterAlgo, self.bwdDataAlgo = ConvFwdAlgo.implicitGemm, ConvBwdFilterAlgo.algo0, ConvBwdDataAlgo.algo0

		if inmaps % groups != 0 or outmaps % groups != 0:
			raise ModuleError(
				"Number of input and output maps must be divisible by number of groups "
				"(%d inmaps, %d outmaps, %d groups)" % (inmaps, outmaps, groups)
			)

		inmaps //= groups
		self.W, self.b = None, None

		if empty:
			return

		Wshape = (outmaps, inmaps, *self.repeat(size, nd))
		W = self.createTensorWithScheme(initscheme, Wshape, wscale)
		self.setVar("W", Variable(gpuarray.empty(Wshape, dtype=self.calctype) if W is None else gpuarray.to_gpu(W)))

		if useBias:
			self.setVar("b", Variable(gpuarray.zeros((1, outmaps) + self.repeat(1, nd), dtype=self.calctype)))

	def optimizeForShape(self, shape, memlimit=None):
		Dnn.convNdbenchmark(shape, self.W.shape, self.stride, self.pad, self.dilation, self.groups, transpose=False)

	def updateData(self, data):
		self.data = Dnn.convNd(data, self.W, self.b, stride=self.stride, pad=self.pad, dilation=self.dilation,
							   groups=self.groups, algo=self.fwdAlgo)

	def updateGrad(self, grad):
		self.grad = Dnn.convNdBackwardData(grad, self.W, data=self.inData, stride=self.stride, pad=self.pad,
										   dilation=self.dilation, groups=self.groups, algo=self.bwdDataAlgo)

	def accGradParams(self, grad, scale=1.0, momentum=0.0):
		Dnn.convNdBackwardParams(
			self.inData, grad, self.W, self.b, stride=self.stride, pad=self.pad, dilation=self.dilation, groups=self.groups,
			wgrad=self.vars["W"].grad, bgrad=self.vars["b"].grad if self.b is not None else None, scale=scale,
			momentum=momentum, algo=self.bwdFilterAlgo
		)

	def calcMode(self, T):
		if np.dtype(T) not in _floatTypes():
			raise ModuleError("Unsupported dtype %s" % T)
		self._recastVars(T)


class Conv2D(ConvND):
	def __init__(self, inmaps, outmaps, size, stride=1, pad=0, dilation=1, wscale=1.0, useBias=True, name=None,
				 initscheme=None, empty=False, groups=1):
		super().__init__(2, inmaps, outmaps, size, stride, pad, dilation, wscale, useBias, name, initscheme, empty, groups)

	def checkDataShape(self, shape):
		if len(shape) != 4:
			raise ModuleError("Data must be 4d tensor")

		_, inmaps, inh, inw = shape
		_, _, fh, fw = self.W.shape
		hpad, wpad = self.pad
		hdilation, wdilation = self.dilation

		if inmaps != self.W.shape[1] * self.groups:
			raise ModuleError("Data has %d maps (expected: %d)" % (inmaps, self.W.shape[1] * self.groups))

		exth, extw = inh + 2 * hpad, inw + 2 * wpad
		extfh, extfw = hdilation * (fh - 1) + 1, wdilation * (fw - 1) + 1

		if exth < extfh:
			raise ModuleError("Data maps height is too small (got %d, expected at least %d)" % (exth, extfh))
		if extw < extfw:
			raise ModuleError("Data maps width is too small (got %d, expected at least %d)" % (extw, extfw))

	def dataShapeFrom(self, shape):
		batchsize, inmaps, inh, inw = shape
		outmaps, _, fh, fw = self.W.shape
		hpad, wpad = self.pad
		hdilation, wdilation = self.dilation
		hstride, wstride = self.stride

		outh = (inh + 2 * hpad - hdilation * (fh - 1) - 1) // hstride + 1
		outw = (inw + 2 * wpad - wdilation * (fw - 1) - 1) // wstride + 1
		return batchsize, outmaps, outh, outw

	def checkGradShape(self, shape):
		if len(shape) != 4:
			raise ModuleError("Grad must be 4d tensor")
		if shape[1] != self.W.shape[0]:
			raise ModuleError("Grad has %d maps (expected: %d)" % (shape[1], self.W.shape[0]))

	def gradShapeFrom(self, shape):
		batchsize, outmaps, outh, outw = shape
		_, inmaps, fh, fw = self.W.shape
		hpad, wpad = self.pad
		hdilation, wdilation = self.dilation
		hstride, wstride = self.stride

		inh = (outh - 1) * hstride + hdilation * (fh - 1) - 2 * hpad + 1
		inw = (outw - 1) * wstride + wdilation * (fw - 1) - 2 * wpad + 1
		return batchsize, inmaps * self.groups, inh, inw


class Conv3D(ConvND):
	"""reference: Modules/Conv3D.py (ConvND with nd = 3; the backend folds the filter depth into the channels, dnn3d.py)"""

	def __init__(self, inmaps, outmaps, size, stride=1, pad=0, dilation=1, wscale=1.0, useBias=True, name=None,
				 initscheme=None, empty=False, groups=1):
		super().__init__(3, inmaps, outmaps, size, stride, pad, dilation, wscale, useBias, name, initscheme, empty, groups)

	def checkDataShape(self, shape):
		if len(shape) != 5:
			raise ModuleError("Data must be 5d tensor")
		if shape[1] != self.W.shape[1] * self.groups:
			raise ModuleError("Data has %d maps (expected: %d)" % (shape[1], self.W.shape[1] * self.groups))
		for i, name in enumerate(("depth", "height", "width")):
			ext, extf = shape[2 + i] + 2 * self.pad[i], self.dilation[i] * (self.W.shape[2 + i] - 1) + 1
			if ext < extf:
				raise ModuleError("Data maps %s is too small (got %d, expected at least %d)" % (name, ext, extf))

	def dataShapeFrom(self, shape):
		return (shape[0], self.W.shape[0]) + tuple(
			(shape[2 + i] + 2 * self.pad[i] - self.dilation[i] * (self.W.shape[2 + i] - 1) - 1) // self.stride[i] + 1 for i in range(3))

	def checkGradShape(self, shape):
		if len(shape) != 5:
			raise ModuleError("Grad must be 5d tensor")
		if shape[1] != self.W.shape[0]:
			raise ModuleError("Grad has %d maps (expected: %d)" % (shape[1], self.W.shape[0]))

	def gradShapeFrom(self, shape):
		return (shape[0], self.W.shape[1] * self.groups) + tuple(
			(shape[2 + i] - 1) * self.stride[i] + self.dilation[i] * (self.W.shape[2 + i] - 1) - 2 * self.pad[i] + 1 for i in range(3))


class Conv1D(ConvND):
	"""1-d convolution run as a 2-d one with H = 1 (reference: Modules/Conv1D.py:14-35)"""

	def __init__(self, inmaps, outmaps, size, stride=1, pad=0, dilation=1, wscale=1.0, useBias=True, name=None,
				 initscheme=None, empty=False, groups=1):
		super().__init__(2, inmaps, outmaps, (1, size), (1, stride), (0, pad), (1, dilation), wscale, useBias, name,
						 initscheme, empty, groups)

	def updateData(self, data):
		data = data.reshape(*data.shape[:2], 1, *data.shape[2:])
		super().updateData(data)
		self.data = self.data.reshape(*self.data.shape[:2], *self.data.shape[3:])

	def updateGrad(self, grad):
		grad = grad.reshape(*grad.shape[:2], 1, *grad.shape[2:])
		data = self.inData
		self.inData = data.reshape(*data.shape[:2], 1, *data.shape[2:])
		super().updateGrad(grad)
		self.inData = data
		self.grad = self.grad.reshape(*self.grad.shape[:2], *self.grad.shape[3:])

	def accGradParams(self, grad, scale=1.0, momentum=0.0):
		grad = grad.reshape(*grad.shape[:2], 1, *grad.shape[2:])
		data = self.inData
		self.inData = data.reshape(*data.shape[:2], 1, *data.shape[2:])
		super().accGradParams(grad, scale, momentum)
		self.inData = data

	def checkDataShape(self, shape):
		if len(shape) != 3:
			raise ModuleError("Data must be 3d tensor")
		if shape[1] != self.W.shape[1] * self.groups:
			raise ModuleError("Data has %d maps (expected: %d)" % (shape[1], self.W.shape[1] * self.groups))

	def checkGradShape(self, shape):
		if len(shape) != 3:
			raise ModuleError("Grad must be 3d tensor")
		if shape[1] != self.W.shape[0]:
			raise ModuleError("Grad has %d maps (expected: %d)" % (shape[1], self.W.shape[0]))

	def dataShapeFrom(self, shape):
		batchsize, _, insize = shape
		outsize = (insize + 2 * self.pad[1] - self.dilation[1] * (self.W.shape[3] - 1) - 1) // self.stride[1] + 1
		return batchsize, self.W.shape[0], outsize

	def gradShapeFrom(self, shape):
		batchsize, _, outsize = shape
		insize = (outsize - 1) * self.stride[1] + self.dilation[1] * (self.W.shape[3] - 1) - 2 * self.pad[1] + 1
		return batchsize, self.W.shape[1] * self.groups, insize


class DeconvND(Module):
	"""reference: Modules/DeconvND.py:13-102 -- forward = conv backward-data (+bias), backward-data = conv forward"""

	def __init__(self, nd, inmaps, outmaps, size, stride=1, pad=0, dilation=1, postpad=0, wscale=1.0, useBias=True,
				 name=None, initscheme=None, empty=False, groups=1):
		super().__init__(name)

		self.stride = self.repeat(stride, nd)
		self.pad = self.repeat(pad, nd)
		self.dilation = self.repeat(dilation, nd)
		self.postpad = self.repeat(postpad, nd)

		self.useBias = useBias
		self.groups = groups

		self.fwdAlgo, self.bwdFilterAlgo, self.bwdDataAlgo = ConvBwdDataAlgo.algo0, ConvBwdFilterAlgo.algo0, ConvFwdAlgo.implicitGemm

		if inmaps % groups != 0 or outmaps % groups != 0:
			raise ModuleError(
				"Number of input and output maps must be divisible by number of groups "
				"(%d inmaps, %d outmaps, %d groups)" % (inmaps, outmaps, groups)
			)

		outmaps //= groups
		self.W, self.b = None, None

		if empty:
			return

		Wshape = (inmaps, outmaps, *self.repeat(size, nd))
		W = self.createTensorWithScheme(initscheme, Wshape, wscale)
		self.setVar("W", Variable(gpuarray.empty(Wshape, dtype=self.calctype) if W is None else gpuarray.to_gpu(W)))

		if useBias:
			# sic: `outmaps` was already divided by groups (reference quirk Q12, DeconvND.py:38,52-53)
			self.setVar("b", Variable(gpuarray.zeros((1, outmaps) + self.repeat(1, nd), dtype=self.calctype)))

	def updateData(self, data):
		self.data = Dnn.deconvNd(data, self.W, self.b, stride=self.stride, pad=self.pad, dilation=self.dilation,
								 postpad=self.postpad, groups=self.groups, algo=self.fwdAlgo)

	def updateGrad(self, grad):
		self.grad = Dnn.deconvNdBackwardData(grad, self.W, data=self.inData, stride=self.stride, pad=self.pad,
											 dilation=self.dilation, groups=self.groups, algo=self.bwdDataAlgo)

	def accGradParams(self, grad, scale=1.0, momentum=0.0):
		Dnn.deconvNdBackwardParams(
			self.inData, grad, self.W, self.b, stride=self.stride, pad=self.pad, dilation=self.dilation, groups=self.groups,
			wgrad=self.vars["W"].grad, bgrad=self.vars["b"].grad if self.b is not None else None, scale=scale,
			momentum=momentum, algo=self.bwdFilterAlgo
		)

	def calcMode(self, T):
		if np.dtype(T) not in _floatTypes():
			raise ModuleError("Unsupported dtype %s" % T)
		self._recastVars(T)


class Deconv2D(DeconvND):
	def __init__(self, inmaps, outmaps, size, stride=1, pad=0, dilation=1, postpad=0, wscale=1.0, useBias=True, name=None,
				 initscheme=None, empty=False, groups=1):
		super().__init__(2, inmaps, outmaps, size, stride, pad, dilation, postpad, wscale, useBias, name, initscheme, empty,
						 groups)

	def checkDataShape(self, shape):
		if len(shape) != 4:
			raise ModuleError("Data must be 4d tensor")
		if shape[1] != self.W.shape[0]:
			raise ModuleError("Data has %d maps (expected: %d)" % (shape[1], self.W.shape[0]))

	def dataShapeFrom(self, shape):
		batchsize, inmaps, inh, inw = shape
		_, outmaps, fh, fw = self.W.shape
		outh = (inh - 1) * self.stride[0] + self.dilation[0] * (fh - 1) - 2 * self.pad[0] + 1 + self.postpad[0]
		outw = (inw - 1) * self.stride[1] + self.dilation[1] * (fw - 1) - 2 * self.pad[1] + 1 + self.postpad[1]
		return batchsize, outmaps * self.groups, outh, outw

	def checkGradShape(self, shape):
		if len(shape) != 4:
			raise ModuleError("Grad must be 4d tensor")
		if shape[1] != self.W.shape[1] * self.groups:
			raise ModuleError("Grad has %d maps (expected: %d)" % (shape[1], self.W.shape[1] * self.groups))

	def gradShapeFrom(self, shape):
		batchsize, outmaps, outh, outw = shape
		inmaps, _, fh, fw = self.W.shape
		inh = (outh + 2 * self.pad[0] - self.dilation[0] * (fh - 1) - 1) // self.stride[0] + 1
		inw = (outw + 2 * self.pad[1] - self.dilation[1] * (fw - 1) - 1) // self.stride[1] + 1
		return batchsize, inmaps, inh, inw


# ---------------------------------------------------------------------------------------------------------- linear
class Linear(Module):
	def __init__(self, insize, outsize, wscale=1.0, useBias=True, initscheme=None, name=None, empty=False, transpose=False):
		super().__init__(name)

		self.transpose = transpose
		self.useBias = useBias
		self.W, self.b = None, None

		if empty:
			return

		Wshape, bshape = ((outsize, insize), (insize, )) if transpose else ((insize, outsize), (outsize, ))
		W = self.createTensorWithScheme(initscheme, Wshape, wscale, factorShape=Wshape)
		self.setVar("W", Variable(gpuarray.empty(Wshape, dtype=self.calctype) if W is None else gpuarray.to_gpu(W)))

		if useBias:
			self.setVar("b", Variable(gpuarray.zeros(bshape, dtype=self.calctype)))

	def updateData(self, data):
		if self.useBias:
			# bias add folded into the GEMM epilogue: one kernel instead of gemm + addVecToMat (Linear.py:36-40)
			self.data = backend().blas.gemmBias(data, self.W, self.b, transpB=self.transpose, allocator=memoryPool())
		else:
			self.data = Blas.mulMatrixOnMatrix(data, self.W, transpB=self.transpose)

	def updateGrad(self, grad):
		self.grad = Blas.mulMatrixOnMatrix(grad, self.W, transpB=not self.transpose)

	def accGradParams(self, grad, scale=1.0, momentum=0.0):
		if not self.transpose:
			Blas.mulMatrixOnMatrix(self.inData, grad, out=self.vars["W"].grad, transpA=True, alpha=scale, beta=momentum)
		else:
			Blas.mulMatrixOnMatrix(grad, self.inData, out=self.vars["W"].grad, transpA=True, alpha=scale, beta=momentum)

		if self.useBias:
			Blas.sumOnMatrix(grad, out=self.vars["b"].grad, alpha=scale, beta=momentum)

	def dataShapeFrom(self, shape):
		return (shape[0], self.W.shape[1]) if not self.transpose else (shape[0], self.W.shape[0])

	def checkDataShape(self, shape):
		if len(shape) != 2:
			raise ModuleError("Data must be 2d matrix")

		expected = self.W.shape[1] if self.transpose else self.W.shape[0]
		if shape[1] != expected:
			raise ModuleError("Expected %d data dimensions, %d were given" % (expected, shape[1]))

	def gradShapeFrom(self, shape):
		return (shape[0], self.W.shape[0]) if not self.transpose else (shape[0], self.W.shape[1])

	def checkGradShape(self, shape):
		if len(shape) != 2:
			raise ModuleError("Grad must be 2d matrix")

		expected = self.W.shape[0] if self.transpose else self.W.shape[1]
		if shape[1] != expected:
			raise ModuleError("Expected %d grad dimensions, %d were given" % (expected, shape[1]))

	def calcMode(self, T):
		if np.dtype(T) not in _floatTypes():
			raise ModuleError("Unsupported dtype %s" % T)
		self._recastVars(T)


# ---------------------------------------------------------------------------------------------------------- recurrent
class RNNMode(str, Enum):
	relu = "relu"
	tanh = "tanh"
	lstm = "lstm"
	gru = "gru"


class DirectionMode(str, Enum):
	uni = "uni"
	bi = "bi"


class WeightModifier(str, Enum):
	orthogonal = "orthogonal"
	identity = "identity"


class RNN(Module):
	"""reference: Modules/RNN.py:31-240 (same constructor, parameter initialisation draws, last-step selection and gradient
	scatter in Python; relu / tanh / lstm / gru, uni- and bidirectional)"""

	def __init__(self, insize, hsize, layers=1, mode="relu", direction="uni", dropout=0.0, getSequences=False, initscheme=None,
				 modifier="orthogonal", wscale=1.0, hintBatchSize=None, name=None):
		super().__init__(name)
		self.gradUsesOutData = True

		self.insize, self.hsize, self.layers = insize, hsize, layers
		self.mode, self.direction = RNNMode(mode), DirectionMode(direction)
		self.dropout, self.getSequences, self.hintBatchSize = dropout, getSequences, hintBatchSize

		bnd = backend()
		mode = {RNNMode.relu: bnd.RNNMode.relu, RNNMode.tanh: bnd.RNNMode.tanh, RNNMode.lstm: bnd.RNNMode.lstm,
				RNNMode.gru: bnd.RNNMode.gru}[self.mode]
		direction = {DirectionMode.uni: bnd.DirectionMode.uni, DirectionMode.bi: bnd.DirectionMode.bi}[self.direction]

		self.descRnn, W, params = Rnn.createRnn(insize, hsize, layers, mode, direction, dropout, seed=np.random.randint(1 << 31),
												batchsize=hintBatchSize)
		self.W = None
		self.setVar("W", Variable(W))
		self.params = params
		self.initParams(initscheme, wscale, modifier)
		self.reserve, self.fulldata, self.dw = None, None, None

	def initParams(self, initscheme, wscale, modifier):
		modifier = WeightModifier(modifier)
		for key in sorted(self.params.keys()):
			for paramName, param in sorted(self.params[key].items()):
				if paramName.startswith("b"):
					param.fill(0.0)
					continue
				if paramName.startswith("r"):
					if modifier == WeightModifier.orthogonal:
						a = np.random.normal(0.0, 1.0, param.shape)
						u, _, v = np.linalg.svd(a, full_matrices=False)
						W = u if u.shape == param.shape else v
						W = W[:param.shape[0], :param.shape[1]].astype(np.float32)
					elif modifier == WeightModifier.identity:
						W = np.identity(param.shape[0], dtype=np.float32)
					else:
						raise NotImplementedError(modifier)
				else:
					W = self.createTensorWithScheme(initscheme, param.shape, wscale)
					if W is None:
						continue
				param.set(W)
		self.updateDeviceMemory()

	def updateDeviceMemory(self):
		Rnn.updateRnnParams(self.descRnn, self.W, self.params)

	def setVar(self, name, var):
		if name == "W" and hasattr(self, "params"):
			_, self.params = Rnn.acquireRnnParams(self.descRnn, w=var.data)
		super().setVar(name, var)

	def updateData(self, data):
		if self.train:
			self.fulldata, self.reserve = Rnn.forwardRnn(data, self.W, self.descRnn)
		else:
			self.fulldata = Rnn.forwardRnn(data, self.W, self.descRnn, test=True)
		if self.direction == DirectionMode.uni:
			self.data = self.fulldata if self.getSequences else self.fulldata[-1]
		elif self.getSequences:
			self.data = self.fulldata
		else:
			# last step of the forward direction, first step of the backward one (Modules/RNN.py:131-138)
			sections = (self.hsize, self.hsize)
			self.data = [gpuarray.split(self.fulldata[-1], sections, axis=1)[0], gpuarray.split(self.fulldata[0], sections, axis=1)[1]]

	def updateGrad(self, grad):
		if self.getSequences:
			fullgrad = grad
		else:
			seqlen = self.fulldata.shape[0]
			if self.direction == DirectionMode.uni:
				fullgrad = gpuarray.empty((seqlen, ) + grad.shape, dtype=grad.dtype, allocator=memoryPool())
				if seqlen > 1:
					fullgrad[:seqlen - 1].fill(0.0)
				fullgrad[seqlen - 1].set(grad)
			else:
				fwdgrad, bwdgrad = grad
				batchsize = fwdgrad.shape[0]
				fullgrad = gpuarray.zeros((seqlen, batchsize, 2 * self.hsize), dtype=fwdgrad.dtype, allocator=memoryPool())
				fullgrad[0, :, self.hsize:].set(bwdgrad)
				fullgrad[seqlen - 1, :, :self.hsize].set(fwdgrad)
		self.grad, self.reserve = Rnn.backwardDataRnn(fullgrad, self.fulldata, self.W, self.reserve, self.descRnn)

	def accGradParams(self, grad, scale=1.0, momentum=0.0):
		self.dw = Rnn.backwardParamsRnn(self.inData, self.fulldata, self.W, self.reserve, self.descRnn)
		Blas.addVectorToVector(self.dw, self.getVar("W").grad, out=self.getVar("W").grad, alpha=scale, beta=momentum)

	def checkDataShape(self, shape):
		if len(shape) != 3:
			raise ModuleError("Data must be 3d tensor")
		if self.hintBatchSize is not None and shape[1] != self.hintBatchSize:
			raise ModuleError("Data batch size must be = %s (was given %s)" % (self.hintBatchSize, shape[1]))
		if shape[2] != self.insize:
			raise ModuleError("Data must have data size = %s (was given %s)" % (self.insize, shape[2]))

	def checkGradShape(self, shape):
		if self.getSequences:
			if len(shape) != 3:
				raise ModuleError("Grad must be 3d tensor")
		elif self.direction == DirectionMode.uni:
			if len(shape) != 2:
				raise ModuleError("Grad must be 2d matrix")
			if shape[-1] != self.hsize:
				raise ModuleError("Grad must have data size = %s (was given %s)" % (self.hsize, shape[-1]))
		else:
			fwdshape, bwdshape = shape
			if len(fwdshape) != 2 or len(bwdshape) != 2:
				raise ModuleError("Grads must be 2d matrices")
			if fwdshape[-1] != self.hsize or bwdshape[-1] != self.hsize:
				raise ModuleError("Grads must have data size = %s (was given %s and %s)" % (self.hsize, fwdshape[1], bwdshape[1]))

	def dataShapeFrom(self, shape):
		hsize = self.hsize if self.direction == DirectionMode.uni else 2 * self.hsize
		if self.getSequences:
			return shape[:2] + (hsize, )
		return (shape[1], hsize) if self.direction == DirectionMode.uni else [(shape[1], self.hsize), (shape[1], self.hsize)]

	def gradShapeFrom(self, shape):
		if self.getSequences:
			batchsize = shape[1]
		else:
			batchsize = shape[0] if self.direction == DirectionMode.uni else shape[0][0]
		return self.inData.shape[0], batchsize, self.insize

	def reset(self):
		super().reset()
		self.reserve, self.fulldata, self.dw = None, None, None

	def calcMode(self, T):
		if T != np.float32:
			raise ModuleError("Unsupported dtype %s" % T)
		self.calctype = T


# ---------------------------------------------------------------------------------------------------------- norm
class BatchNormND(Module):
	def __init__(self, nd, maps, epsilon=1e-5, initFactor=1.0, minFactor=0.1, sscale=0.01, affine=True, name=None,
				 empty=False, inplace=False):
		super().__init__(name)
		self.inplace = inplace

		self.maps = maps
		self.epsilon = epsilon
		self.initFactor = initFactor
		self.minFactor = minFactor
		self.numOfProps = 0
		self.affine = affine

		self.scale, self.bias, self.mean, self.var = None, None, None, None
		self.savemean, self.saveinvvar, self.scalegrad, self.biasgrad = None, None, None, None

		if empty:
			return

		shape = (1, maps) + self.repeat(1, nd)
		scale = np.random.normal(1.0, sscale if affine else 0.0, shape).astype(np.float32)

		# parameters and running statistics stay fp32 whatever the compute type is (CuDnnNorm.c:118-121)
		self.setVar("scale", Variable(gpuarray.to_gpu(scale)))
		self.setVar("bias", Variable(gpuarray.zeros(shape, dtype=np.float32)))

		self.setAttr("mean", gpuarray.zeros(shape, dtype=np.float32))
		self.setAttr("var", gpuarray.to_gpu(np.ones(shape, dtype=np.float32)))

	def updateData(self, data):
		if self.train:
			if self.inplace:
				raise ModuleError("%s: using inplace flag in train mode is prohibited" % self)

			self.numOfProps += 1
			factor = max(self.initFactor / self.numOfProps, self.minFactor)

			self.data, self.savemean, self.saveinvvar = Dnn.batchNormNd(
				data, self.scale, self.bias, self.mean, self.var, self.epsilon, factor, False
			)
		else:
			self.data = Dnn.batchNormNd(
				data, self.scale, self.bias, self.mean, self.var, self.epsilon, 0, True, out=data if self.inplace else None
			)

	def updateGrad(self, grad):
		tup = Dnn.batchNormNdBackward(self.inData, grad, self.scale, self.savemean, self.saveinvvar, self.epsilon)

		if self.affine:
			self.grad, self.scalegrad, self.biasgrad = tup
		else:
			self.grad, _, _ = tup

	def accGradParams(self, grad, scale=1.0, momentum=0.0):
		if self.affine:
			Blas.addVectorToVector(
				self.scalegrad.ravel(), self.vars["scale"].grad.ravel(), out=self.vars["scale"].grad.ravel(),
				alpha=scale, beta=momentum
			)
			Blas.addVectorToVector(
				self.biasgrad.ravel(), self.vars["bias"].grad.ravel(), out=self.vars["bias"].grad.ravel(),
				alpha=scale, beta=momentum
			)

	def dataShapeFrom(self, shape):
		return shape

	def gradShapeFrom(self, shape):
		return shape

	def reset(self):
		super().reset()
		self.savemean, self.saveinvvar = None, None
		if self.affine:
			self.scalegrad, self.biasgrad = None, None

	def calcMode(self, T):
		if np.dtype(T) not in _floatTypes():
			raise ModuleError("Unsupported dtype %s" % T)
		self.calctype = T


class BatchNorm2D(BatchNormND):
	def __init__(self, maps, epsilon=1e-5, initFactor=1.0, minFactor=0.1, sscale=0.01, affine=True, name=None, empty=False,
				 inplace=False):
		super().__init__(2, maps, epsilon, initFactor, minFactor, sscale, affine, name, empty, inplace)

	def checkDataShape(self, shape):
		if len(shape) != 4:
			raise ModuleError("Data must be 4d tensor")
		if shape[1] != self.maps:
			raise ModuleError("Data has %d maps (expected: %d)" % (shape[1], self.maps))

	def checkGradShape(self, shape):
		if len(shape) != 4:
			raise ModuleError("Grad must be 4d tensor")
		if shape[1] != self.maps:
			raise ModuleError("Grad has %d maps (expected: %d)" % (shape[1], self.maps))


class BatchNorm(BatchNormND):
	"""Batch norm over (batch, features) matrices, run as a (N, C, 1, 1) spatial BN (reference: Modules/BatchNorm.py)"""

	def __init__(self, size, epsilon=1e-5, initFactor=1.0, minFactor=0.1, sscale=0.01, affine=True, name=None, empty=False,
				 inplace=False):
		super().__init__(2, size, epsilon, initFactor, minFactor, sscale, affine, name, empty, inplace)

	def updateData(self, data):
		indata = data.reshape(*data.shape, 1, 1)
		super().updateData(indata)
		self.data = self.data.reshape(data.shape)

	def updateGrad(self, grad):
		data = self.inData
		self.inData = data.reshape(*data.shape, 1, 1)
		super().updateGrad(grad.reshape(*grad.shape, 1, 1))
		self.inData = data
		self.grad = self.grad.reshape(grad.shape)

	def checkDataShape(self, shape):
		if len(shape) != 2:
			raise ModuleError("Data must be 2d matrix")
		if shape[1] != self.maps:
			raise ModuleError("Data has %d features (expected: %d)" % (shape[1], self.maps))

	def checkGradShape(self, shape):
		if len(shape) != 2:
			raise ModuleError("Grad must be 2d matrix")
		if shape[1] != self.maps:
			raise ModuleError("Grad has %d features (expected: %d)" % (shape[1], self.maps))


class InstanceNorm2D(Module):
	"""reference: Modules/InstanceNorm2D.py:12-97"""

	def __init__(self, numOfMaps, epsilon=1e-5, affine=True, name=None):
		super().__init__(name)
		self.numOfMaps = numOfMaps
		self.epsilon = epsilon
		self.affine = affine

		shape = (1, numOfMaps, 1, 1)
		self.setVar("scale", Variable(gpuarray.to_gpu(np.ones(shape, dtype=np.float32))))
		self.setVar("bias", Variable(gpuarray.zeros(shape, dtype=np.float32)))

		self.savemean, self.saveinvvar, self.extscale, self.scalegrad, self.biasgrad = None, None, None, None, None

	def updateData(self, data):
		self.data, self.savemean, self.saveinvvar, self.extscale = Dnn.instanceNorm2d(data, self.scale, self.bias, self.epsilon)

	def updateGrad(self, grad):
		if self.affine:
			self.grad, self.scalegrad, self.biasgrad = Dnn.instanceNorm2dBackward(
				grad, self.inData, self.extscale, self.savemean, self.saveinvvar, self.epsilon, True
			)
		else:
			self.grad = Dnn.instanceNorm2dBackward(
				grad, self.inData, self.extscale, self.savemean, self.saveinvvar, self.epsilon, False
			)

	def accGradParams(self, grad, scale=1.0, momentum=0.0):
		if self.affine:
			Blas.addVectorToVector(self.scalegrad.ravel(), self.vars["scale"].grad.ravel(),
								   out=self.vars["scale"].grad.ravel(), alpha=scale, beta=momentum)
			Blas.addVectorToVector(self.biasgrad.ravel(), self.vars["bias"].grad.ravel(),
								   out=self.vars["bias"].grad.ravel(), alpha=scale, beta=momentum)

	def checkDataShape(self, shape):
		if len(shape) != 4:
			raise ModuleError("Data must be 4d tensor")
		if shape[1] != self.numOfMaps:
			raise ModuleError("Data has %d maps (expected: %d)" % (shape[1], self.numOfMaps))

	def dataShapeFrom(self, shape):
		return shape

	def gradShapeFrom(self, shape):
		return shape

	def reset(self):
		super().reset()
		self.savemean, self.saveinvvar, self.extscale = None, None, None
		if self.affine:
			self.scalegrad, self.biasgrad = None, None

	def calcMode(self, T):
		if np.dtype(T) not in _floatTypes():
			raise ModuleError("Unsupported dtype %s" % T)
		self.calctype = T


# ---------------------------------------------------------------------------------------------------------- pooling
class Pool2D(Module):
	def __init__(self, size=2, stride=2, pad=0, name=None):
		super().__init__(name)
		self.gradUsesOutData = True

		self.size = self.repeat(size, 2)
		self.stride = self.repeat(stride, 2)
		self.pad = self.repeat(pad, 2)
		self.workspace = None

	def dataShapeFrom(self, shape):
		batchsize, maps, inh, inw = shape
		outh = (inh + 2 * self.pad[0] - self.size[0]) // self.stride[0] + 1
		outw = (inw + 2 * self.pad[1] - self.size[1]) // self.stride[1] + 1
		return batchsize, maps, outh, outw

	def checkDataShape(self, shape):
		if len(shape) != 4:
			raise ModuleError("Data must be 4d tensor")

		_, _, inh, inw = shape
		if inh + 2 * self.pad[0] < self.size[0]:
			raise ModuleError("Data maps height is too small (got %d, expected at least %d)" % (inh + 2 * self.pad[0], self.size[0]))
		if inw + 2 * self.pad[1] < self.size[1]:
			raise ModuleError("Data maps width is too small (got %d, expected at least %d)" % (inw + 2 * self.pad[1], self.size[1]))

	def gradShapeFrom(self, shape):
		batchsize, maps, outh, outw = shape
		inh = (outh - 1) * self.stride[0] - 2 * self.pad[0] + self.size[0]
		inw = (outw - 1) * self.stride[1] - 2 * self.pad[1] + self.size[1]
		return batchsize, maps, inh, inw

	def checkGradShape(self, shape):
		if len(shape) != 4:
			raise ModuleError("Grad must be 4d tensor")

	def reset(self):
		super().reset()
		self.workspace = None

	def calcMode(self, T):
		if np.dtype(T) not in _floatTypes():
			raise ModuleError("Unsupported dtype %s" % T)
		self.calctype = T


class MaxPool2D(Pool2D):
	def __init__(self, size=2, stride=2, pad=0, useMask=False, name=None):
		super().__init__(size, stride, pad, name)
		self.useMask = useMask
		self.mask = None
		self.mode = PoolMode.max

	@property
	def withMask(self):
		return self.useMask

	@withMask.setter
	def withMask(self, val):
		self.useMask = val
		self.gradUsesOutData = False if val else True

	def updateData(self, data):
		if self.useMask:
			self.data, self.mask = Pool.maxpool2d(data, size=self.size, stride=self.stride, pad=self.pad)
		else:
			self.data, self.workspace = Dnn.poolNd(data, size=self.size, stride=self.stride, pad=self.pad, mode=self.mode,
												   test=not self.train)

	def updateGrad(self, grad):
		if self.useMask:
			self.grad = Pool.maxpool2dBackward(grad, self.inData.shape, self.mask, size=self.size, stride=self.stride, pad=self.pad)
		else:
			self.grad = Dnn.poolNdBackward(self.inData, self.data, grad, self.workspace, size=self.size, stride=self.stride,
										   pad=self.pad, mode=self.mode)

	def reset(self):
		super().reset()
		self.mask = None


class AvgPool2D(Pool2D):
	def __init__(self, size=2, stride=2, pad=0, includePad=True, name=None):
		super().__init__(size, stride, pad, name)
		self.mode = PoolMode.avgWithPad if includePad else PoolMode.avgNoPad

	def updateData(self, data):
		self.data, self.workspace = Dnn.poolNd(data, size=self.size, stride=self.stride, pad=self.pad, mode=self.mode,
											   test=not self.train)

	def updateGrad(self, grad):
		self.grad = Dnn.poolNdBackward(self.inData, self.data, grad, self.workspace, size=self.size, stride=self.stride,
									   pad=self.pad, mode=self.mode)


class Pool3D(Module):
	"""reference: Modules/Pool3D.py (the backend pools the slices, then the depth axis: dnn3d.py)"""

	def __init__(self, size=2, stride=2, pad=0, name=None):
		super().__init__(name)
		self.gradUsesOutData = True
		self.size, self.stride, self.pad = self.repeat(size, 3), self.repeat(stride, 3), self.repeat(pad, 3)
		self.workspace = None
		self.mode = PoolMode.max

	def updateData(self, data):
		self.data, self.workspace = Dnn.poolNd(data, size=self.size, stride=self.stride, pad=self.pad, mode=self.mode,
											   test=not self.train)

	def updateGrad(self, grad):
		self.grad = Dnn.poolNdBackward(self.inData, self.data, grad, self.workspace, size=self.size, stride=self.stride,
									   pad=self.pad, mode=self.mode)

	def dataShapeFrom(self, shape):
		return tuple(shape[:2]) + tuple((shape[2 + i] + 2 * self.pad[i] - self.size[i]) // self.stride[i] + 1 for i in range(3))

	def checkDataShape(self, shape):
		if len(shape) != 5:
			raise ModuleError("Data must be 5d tensor")
		for i in range(3):
			if shape[2 + i] + 2 * self.pad[i] < self.size[i]:
				raise ModuleError("Data maps are too small on dim #%d" % (i + 1))

	def gradShapeFrom(self, shape):
		return tuple(shape[:2]) + tuple((shape[2 + i] - 1) * self.stride[i] - 2 * self.pad[i] + self.size[i] for i in range(3))

	def checkGradShape(self, shape):
		if len(shape) != 5:
			raise ModuleError("Grad must be 5d tensor")

	def reset(self):
		super().reset()
		self.workspace = None


class MaxPool3D(Pool3D):
	pass


class AvgPool3D(Pool3D):
	def __init__(self, size=2, stride=2, pad=0, includePad=True, name=None):
		super().__init__(size, stride, pad, name)
		self.mode = PoolMode.avgWithPad if includePad else PoolMode.avgNoPad


class MaxUnpool2D(Module):
	"""reference: Modules/MaxUnpool2D.py -- scatters through the mask of a MaxPool2D(useMask=True)"""

	def __init__(self, maxpool2d, name=None):
		super().__init__(name)
		self.maxpool2d = maxpool2d
		self.maxpool2d.withMask = True

	def updateData(self, data):
		self.data = Pool.maxunpool2d(data, self.maxpool2d.inData.shape, self.maxpool2d.mask)

	def updateGrad(self, grad):
		self.grad = Pool.maxunpool2dBackward(grad, self.maxpool2d.data.shape, self.maxpool2d.mask)

	def dataShapeFrom(self, shape):
		return self.maxpool2d.gradShapeFrom(shape)

	def gradShapeFrom(self, shape):
		return self.maxpool2d.dataShapeFrom(shape)


# ---------------------------------------------------------------------------------------------------------- activations
class ActivationType(str, Enum):
	sigmoid = "sigmoid"
	tanh = "tanh"
	relu = "relu"
	leakyRelu = "leakyRelu"
	elu = "elu"
	softPlus = "softPlus"
	clip = "clip"


sigmoid = ActivationType.sigmoid
tanh = ActivationType.tanh
relu = ActivationType.relu
leakyRelu = ActivationType.leakyRelu
elu = ActivationType.elu
softPlus = ActivationType.softPlus
clip = ActivationType.clip


class Activation(Module):
	def __init__(self, activation, slc=None, inplace=False, name=None, args=()):
		super().__init__(name)

		self.gradUsesOutData = True
		self.inplace = inplace

		activation = ActivationType(activation)
		self.activation = activation
		self.slc = slc

		self.actArgs = args if len(args) > 0 else {
			ActivationType.leakyRelu: (0.01, ),
			ActivationType.elu: (1.0, ),
			ActivationType.clip: (0.0, 6.0)
		}.get(activation, ())

	def updateData(self, data):
		self.data = data if self.inplace else gpuarray.empty(data.shape, dtype=data.dtype, allocator=memoryPool())
		getattr(backend(), "%sKer" % self.activation.value)(data.dtype)(self.data, data, *self.actArgs, slice=self.slc)

	def updateGrad(self, grad):
		self.grad = grad if self.inplace else gpuarray.empty(grad.shape, dtype=grad.dtype, allocator=memoryPool())
		getattr(backend(), "%sDerKer" % self.activation.value)(grad.dtype)(self.grad, grad, self.data, *self.actArgs, slice=self.slc)

	def dataShapeFrom(self, shape):
		return shape

	def gradShapeFrom(self, shape):
		return shape

	def calcMode(self, T):
		if np.dtype(T) not in _floatTypes():
			raise ModuleError("Unsupported dtype %s" % T)
		self.calctype = T


class Gelu(Module):
	"""reference: Modules/Gelu.py -- the derivative kernel takes the INPUT (quirk Q5)"""

	def __init__(self, inplace=False, name=None):
		super().__init__(name)
		if inplace:
			raise ModuleError("%s: inplace gelu cannot compute its gradient" % self)

	def updateData(self, data):
		self.data = gpuarray.empty(data.shape, dtype=data.dtype, allocator=memoryPool())
		backend().geluKer(data.dtype)(self.data, data)

	def updateGrad(self, grad):
		self.grad = gpuarray.empty(grad.shape, dtype=grad.dtype, allocator=memoryPool())
		backend().geluDerKer(grad.dtype)(self.grad, grad, self.inData)

	def dataShapeFrom(self, shape):
		return shape

	def gradShapeFrom(self, shape):
		return shape


class SoftMax(Module):
	def __init__(self, name=None):
		super().__init__(name)
		self.gradUsesOutData = True

	def updateData(self, data):
		shape = data.shape
		ndim = max(0, 4 - len(shape))

		data = data.reshape(shape + tuple(1 for _ in range(ndim)))
		self.data = Dnn.softmaxNd(data).reshape(shape)

	def updateGrad(self, grad):
		shape = grad.shape
		ndim = max(0, 4 - len(shape))

		grad = grad.reshape(shape + tuple(1 for _ in range(ndim)))
		data = self.data.reshape(shape + tuple(1 for _ in range(ndim)))
		self.grad = Dnn.softmaxNdBackward(data, grad).reshape(shape)

	def dataShapeFrom(self, shape):
		return shape

	def gradShapeFrom(self, shape):
		return shape

	def calcMode(self, T):
		if np.dtype(T) not in _floatTypes():
			raise ModuleError("Unsupported dtype %s" % T)
		self.calctype = T


# ---------------------------------------------------------------------------------------------------------- glue
def _sumInto(arrays):
	"""Sum of k same-shaped arrays into a fresh pool array: Add.updateData / Replicate.updateGrad (Add.py:15-23)."""
	first = arrays[0]
	out = gpuarray.empty(first.shape, dtype=first.dtype, allocator=memoryPool())

	if Config.fuseAdd and len(arrays) >= 2:
		backend().add2Ker(first.dtype)(out, arrays[0], arrays[1])
		rest = arrays[2:]
	else:
		out.fill(0)
		rest = arrays

	for ary in rest:
		Blas.toVectorAddVector(out.ravel(), ary.ravel())
	return out


class Add(Module):
	def __init__(self, name=None):
		super().__init__(name)
		self.movesGrad = True

	def updateData(self, data):
		self.data = _sumInto(data)

	def updateGrad(self, grad):
		self.grad = [grad] * len(self.inData)

	def checkDataShape(self, shapes):
		for shape in shapes:
			if shape != shapes[0]:
				raise ModuleError("Shape %s is not equal to initial shape %s" % (shape, shapes[0]))

	def dataShapeFrom(self, shape):
		return shape[0]

	def gradShapeFrom(self, shape):
		return [shape] * len(self.inData)

	def calcMode(self, T):
		self.calctype = T


class Replicate(Module):
	def __init__(self, times, name=None):
		super().__init__(name)
		self.movesData = True
		self.times = times

	def updateData(self, data):
		self.data = [data] * self.times

	def updateGrad(self, grad):
		self.grad = _sumInto(grad)

	def dataShapeFrom(self, shape):
		return [shape] * self.times

	def gradShapeFrom(self, shape):
		return shape[0]

	def calcMode(self, T):
		self.calctype = T


class Identity(Module):
	def __init__(self, name=None):
		super().__init__(name)
		self.movesData = True
		self.movesGrad = True

	def updateData(self, data):
		self.data = data

	def updateGrad(self, grad):
		self.grad = grad

	def dataShapeFrom(self, shape):
		return shape

	def gradShapeFrom(self, shape):
		return shape

	def calcMode(self, T):
		self.calctype = T


class Flatten(Module):
	def __init__(self, name=None):
		super().__init__(name)
		self.movesData = True
		self.movesGrad = True
		self.inshape = None

	def updateData(self, data):
		self.inshape = data.shape
		self.data = data.reshape(data.shape[0], int(np.prod(data.shape[1:])))

	def updateGrad(self, grad):
		self.grad = grad.reshape(self.inshape)

	def dataShapeFrom(self, shape):
		return shape[0], int(np.prod(shape[1:]))

	def gradShapeFrom(self, shape):
		return (shape[0], ) + self.inshape[1:]

	def calcMode(self, T):
		self.calctype = T


class Dropout(Module):
	"""reference: Modules/Dropout.py:12-95 (train: one random word per element, keep where word < partition and rescale by 1/(1-p);
	eval: identity).  `slicing` is not supported."""

	def __init__(self, p=0.5, rng=None, slicing=None, inplace=False, name=None):
		super().__init__(name)
		if slicing is not None:
			raise NotImplementedError("slicing")
		self.p = p
		self.partition, self.rands = None, None
		self.rng = backend().globalRng if rng is None else rng
		self.inplace = inplace

	def updateData(self, data):
		if not self.train:
			self.data = data
			return

		self.data = data if self.inplace else gpuarray.empty(data.shape, dtype=data.dtype, allocator=memoryPool())
		parttype = np.uint32 if data.dtype == np.float32 else np.uint16
		nwords = (data.nbytes + 3) // 4
		words = gpuarray.empty((nwords, ), dtype=np.uint32, allocator=memoryPool())
		self.rng.fillInteger(words)
		self.rands = words.view(parttype)

		p = 1.0 - self.p
		self.partition = int(p * np.iinfo(parttype).max)
		backend().dropoutKer(data.dtype)(self.data, data, self.rands, self.partition, np.float32(p))

	def updateGrad(self, grad):
		if not self.train:
			self.grad = grad
			return
		self.grad = grad if self.inplace else gpuarray.empty(grad.shape, dtype=grad.dtype, allocator=memoryPool())
		backend().dropoutKer(grad.dtype)(self.grad, grad, self.rands, self.partition, 1.0 - self.p)

	def dataShapeFrom(self, shape):
		return shape

	def gradShapeFrom(self, shape):
		return shape

	def reset(self):
		super().reset()
		self.rands = None

	def calcMode(self, T):
		if np.dtype(T) not in _floatTypes():
			raise ModuleError("Unsupported dtype %s" % T)
		self.calctype = T


class Transpose(Module):
	"""reference: Modules/Transpose.py:7-45"""

	def __init__(self, axes=None, name=None):
		super().__init__(name)
		self.axes = axes
		if axes is None:
			self.invaxes = None
		else:
			self.invaxes = [0] * len(axes)
			for i, axis in enumerate(axes):
				self.invaxes[axis] = i

	def updateData(self, data):
		self.data = Memory.transpose(data, self.axes)

	def updateGrad(self, grad):
		self.grad = Memory.transpose(grad, self.invaxes)

	def checkDataShape(self, shape):
		if self.axes is not None and len(shape) != len(self.axes):
			raise ModuleError("Data dimension needs to be %d, (data has %d)" % (len(self.axes), len(shape)))

	checkGradShape = checkDataShape

	def dataShapeFrom(self, shape):
		return tuple(shape[axis] for axis in (self.axes if self.axes is not None else reversed(range(len(shape)))))

	def gradShapeFrom(self, shape):
		return tuple(shape[axis] for axis in (self.invaxes if self.invaxes is not None else reversed(range(len(shape)))))

	def calcMode(self, T):
		self.calctype = T


class MoveAxis(Module):
	"""reference: Modules/MoveAxis.py:7-60"""

	def __init__(self, src, dst, name=None):
		super().__init__(name)
		if src == dst:
			raise ModuleError("Trivial axis move is treated as error")
		self.src, self.dst = src, dst

	def updateData(self, data):
		self.data = Memory.moveaxis(data, self.src, self.dst)

	def updateGrad(self, grad):
		self.grad = Memory.moveaxis(grad, self.dst, self.src)

	def checkDataShape(self, shape):
		ln = max(self.src, self.dst)
		if len(shape) - 1 < ln:
			raise ModuleError("Data dimension needs to be at least %d, (data has %d)" % (ln + 1, len(shape)))

	checkGradShape = checkDataShape

	@staticmethod
	def _moved(shape, src, dst):
		axes = [a for a in range(len(shape)) if a != src]
		axes.insert(dst, src)
		return tuple(shape[a] for a in axes)

	def dataShapeFrom(self, shape):
		return self._moved(shape, self.src, self.dst)

	def gradShapeFrom(self, shape):
		return self._moved(shape, self.dst, self.src)

	def calcMode(self, T):
		self.calctype = T


class SwapAxes(Module):
	"""reference: Modules/SwapAxes.py:7-55"""

	def __init__(self, axis1, axis2, name=None):
		super().__init__(name)
		if axis1 == axis2:
			raise ModuleError("Trivial axes swap is treated as error")
		self.axis1, self.axis2 = (axis2, axis1) if axis1 > axis2 else (axis1, axis2)

	def updateData(self, data):
		self.data = Memory.swapaxes(data, self.axis1, self.axis2)

	def updateGrad(self, grad):
		self.grad = Memory.swapaxes(grad, self.axis1, self.axis2)

	def checkDataShape(self, shape):
		if len(shape) - 1 < self.axis2:
			raise ModuleError("Data dimension needs to be at least %d, (data has %d)" % (self.axis2 + 1, len(shape)))

	checkGradShape = checkDataShape

	def dataShapeFrom(self, shape):
		shape = list(shape)
		shape[self.axis1], shape[self.axis2] = shape[self.axis2], shape[self.axis1]
		return tuple(shape)

	gradShapeFrom = dataShapeFrom

	def calcMode(self, T):
		self.calctype = T


class DepthConcat(Module):
	"""reference: Modules/DepthConcat.py:7-70 (Inception-style join of branches with different plane sizes)"""

	def __init__(self, name=None):
		super().__init__(name)
		self.movesData = True

	def updateData(self, data):
		self.data = Memory.depthConcat(data)

	def updateGrad(self, grad):
		self.grad = Memory.depthSplit(grad, self.inData)

	def checkDataShape(self, shapes):
		if not isinstance(shapes, list):
			raise ModuleError("Data must be list of tensors")
		for shape in shapes:
			if len(shape) != 4:
				raise ModuleError("Data must consist of 4d tensors")
			if shape[0] != shapes[0][0]:
				raise ModuleError("Inconsistency in batch size")

	def dataShapeFrom(self, shapes):
		depth, h, w = 0, 0, 0
		for shape in shapes:
			depth += shape[1]
			h, w = max(h, shape[2]), max(w, shape[3])
		return shapes[0][0], depth, h, w

	def checkGradShape(self, shape):
		if len(shape) != 4:
			raise ModuleError("Grad must be 4d tensor")
		gradshape = self.dataShapeFrom([data.shape for data in self.inData])
		if shape != gradshape:
			raise ModuleError("Bad grad shape (%s given, %s expected)" % (shape, gradshape))

	def gradShapeFrom(self, shape):
		return [data.shape for data in self.inData]

	def calcMode(self, T):
		self.calctype = T


class Cast(Module):
	"""reference: Modules/Cast.py -- dtype conversion between two compute types (fp32 <-> fp16 / bf16)"""

	def __init__(self, intype, outtype, name=None):
		super().__init__(name)
		self.intype, self.outtype = np.dtype(intype), np.dtype(outtype)
		self.calctype = self.intype

	def updateData(self, data):
		self.data = data.astype(self.outtype)

	def updateGrad(self, grad):
		self.grad = grad.astype(self.intype)

	def checkGradType(self, dtype):
		if dtype != self.outtype:
			raise ModuleError("Expected dtype %s, got %s" % (self.outtype, dtype))

	def dataShapeFrom(self, shape):
		return shape

	def gradShapeFrom(self, shape):
		return shape

	def calcMode(self, T):
		pass


# ---------------------------------------------------------------------------------------------------------- containers
class Container(Module):
	def __init__(self, name=None):
		super().__init__(name)
		self.modules = {}

	def append(self, mod, acquire=True):
		mod.name = str(len(self.modules)) if mod.name is None else mod.name

		if mod.name in self.modules:
			if acquire:
				mod.name = str(len(self.modules))
			else:
				raise ContainerError("Module with name '%s' is already in container" % mod.name)

		self.modules[mod.name] = mod
		return self

	def getByName(self, name):
		if name in self.modules:
			return self.modules[name]

		for m in self.modules.values():
			if isinstance(m, Container):
				mod = m.getByName(name)
				if mod is not None:
					return mod
		return None

	def __getitem__(self, name):
		mod = self.getByName(name)
		if mod is None:
			raise ContainerError("%s: Module %s not found" % (self, name))
		return mod

	def getVarTable(self, vartable=None, name=None, root=True):
		# dotted names "<child>.<grandchild>.<param>" rooted at this container (reference: Container.py:96-103)
		name = "" if root else name
		vartable = {} if vartable is None else vartable

		for mod in self.modules.values():
			mod.getVarTable(vartable, "%s%s." % (name, mod.name), root=False)
		return vartable

	def setVar(self, name, var):
		sep = name.find(".")
		if sep == -1:
			raise ContainerError("Cannot find dot-delimiter in variable name: %s" % name)
		self.modules[name[:sep]].setVar(name[sep + 1:], var)

	def getVar(self, name):
		sep = name.find(".")
		if sep == -1:
			raise ContainerError("Cannot find dot-delimiter in variable name: %s" % name)
		return self.modules[name[:sep]].getVar(name[sep + 1:])

	def leaves(self):
		for mod in self.modules.values():
			if isinstance(mod, Container):
				yield from mod.leaves()
			else:
				yield mod

	def zeroGradParams(self):
		for mod in self.modules.values():
			mod.zeroGradParams()

	def updateParams(self, learnRate):
		for mod in self.modules.values():
			mod.updateParams(learnRate)

	def trainMode(self):
		super().trainMode()
		for mod in self.modules.values():
			mod.trainMode()

	def evalMode(self):
		super().evalMode()
		for mod in self.modules.values():
			mod.evalMode()

	def calcMode(self, T):
		for mod in self.modules.values():
			try:
				mod.calcMode(T)
			except Exception as e:
				self.handleError(mod, e)

	def reset(self):
		super().reset()
		for mod in self.modules.values():
			mod.reset()

	def numOfParams(self):
		return sum(mod.numOfParams() for mod in self.modules.values())

	def paramSize(self):
		return sum(mod.paramSize() for mod in self.modules.values())

	def handleError(self, mod, e):
		msg = str(e)
		msg = ":\n" + msg if len(msg) > 0 else ""
		raise ModuleError("%s:\nError in module %s (type %s)%s" % (self, mod, type(e), msg)) from e


class Sequential(Container):
	def __init__(self, name=None):
		super().__init__(name)
		self.graph = []

	def append(self, mod, acquire=True):
		super().append(mod, acquire)
		self.graph.append(mod)
		return self

	def extend(self, container, acquire=True):
		for mod in container.graph:
			self.append(mod, acquire)
		return self

	def pop(self):
		mod = self.graph.pop()
		self.modules.pop(mod.name)
		return mod

	def __getitem__(self, item):
		if isinstance(item, str):
			return super().__getitem__(item)
		elif isinstance(item, int):
			return self.graph[item]
		elif isinstance(item, slice):
			assert item.step == 1 or item.step is None
			seq = Sequential()
			seq.extend_list(self.graph[item.start:item.stop])
			return seq
		raise NotImplementedError(type(item).__name__)

	def extend_list(self, mods):
		for mod in mods:
			self.append(mod)
		return self

	def updateData(self, data):
		for i, mod in enumerate(self.graph):
			try:
				mod(data)
			except ModuleError as e:
				raise ModuleError("%s:\nData error in module %d (%s):\n%s" % (self, i, mod, e))
			data = mod.data

		self.data = data if len(self.graph) == 0 else self.graph[-1].data

	def dataShapeFrom(self, shape):
		for mod in self.graph:
			shape = mod.dataShapeFrom(shape)
		return shape

	def backward(self, grad, updParamGrads=True, updGrad=True, scale=1.0, momentum=1.0):
		# NB the reference never forwards updGrad=False to the first module (always-true guard), so the
		# first layer's data gradient IS computed; kept for parity of work and results (Sequential.py:212-221)
		for i, mod in enumerate(reversed(self.graph)):
			try:
				mod.backward(grad, updParamGrads=updParamGrads, scale=scale, momentum=momentum)
			except ModuleError as e:
				raise ModuleError("%s:\nGrad error in module %d (%s):\n%s" % (self, len(self.graph) - 1 - i, mod, e))
			grad = mod.grad

		self.grad = grad if len(self.graph) == 0 else self.graph[0].grad

	def gradShapeFrom(self, shape):
		for mod in reversed(self.graph):
			shape = mod.gradShapeFrom(shape)
		return shape

	def updateGrad(self, grad):
		assert False

	def checkDataType(self, dtype):
		pass

	def checkGradType(self, dtype):
		pass


class Parallel(Container):
	def __init__(self, name=None):
		super().__init__(name)
		self.graph = []

	def append(self, mod, acquire=True):
		super().append(mod, acquire)
		self.graph.append(mod)
		return self

	def __getitem__(self, item):
		if isinstance(item, str):
			return super().__getitem__(item)
		elif isinstance(item, int):
			return self.graph[item]
		raise NotImplementedError(type(item).__name__)

	def updateData(self, data):
		assert len(data) == len(self.graph)

		self.data = []
		for i, mod in enumerate(self.graph):
			try:
				mod(data[i])
			except ModuleError as e:
				raise ModuleError("%s:\nData error in module %d (%s):\n%s" % (self, i, mod, e))
			self.data.append(mod.data)

	def dataShapeFrom(self, shapes):
		return [mod.dataShapeFrom(shapes[i]) for i, mod in enumerate(self.graph)]

	def backward(self, grad, updParamGrads=True, updGrad=True, scale=1.0, momentum=1.0):
		assert len(grad) == len(self.graph)

		self.grad = []
		for i, mod in enumerate(self.graph):
			try:
				mod.backward(grad[i], updParamGrads=updParamGrads, updGrad=updGrad, scale=scale, momentum=momentum)
			except ModuleError as e:
				raise ModuleError("%s:\nGrad error in module %d (%s):\n%s" % (self, i, mod, e))
			self.grad.append(mod.grad)

	def gradShapeFrom(self, shapes):
		return [mod.gradShapeFrom(shapes[i]) for i, mod in enumerate(self.graph)]

	def updateGrad(self, grad):
		assert False

	def checkDataType(self, dtype):
		pass

	def checkGradType(self, dtype):
		pass
