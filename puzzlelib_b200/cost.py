"""The cost either side of the hot path: `Cost` bookkeeping and `CrossEntropy` (reference: Cost/Cost.py:10-118,
Cost/CrossEntropy.py:13-80).  The gradient is ascent-direction, `(target - pred) / N`, the error is kept on the
device and only read back when the caller asks for it."""
import numpy as np

from .shim import gpuarray, backend, memoryPool


class CostError(Exception):
	pass


class Cost:
	def __init__(self):
		self.accumErr = gpuarray.empty((), dtype=np.float32)
		self.devErr = gpuarray.empty((), dtype=np.float32)

		self.error, self.valError, self.grad = None, None, None
		self.batchsize, self.numOfSamples = None, None

		self.dirty = True
		self.resetAccumulator()

	def resetAccumulator(self):
		self.resetDeviceAccumulator()
		self.batchsize, self.numOfSamples = 0, 0

	def updateState(self, samples):
		self.batchsize = samples
		self.numOfSamples += samples

	def resetDeviceAccumulator(self):
		self.accumErr.fill(0.0)

	def getError(self):
		if self.dirty:
			self.error = self.devErr.get() / self.batchsize
			self.dirty = False
		return self.error

	def getMeanError(self):
		return self.accumErr.get() / self.numOfSamples

	def getValError(self):
		return self.valError

	def __call__(self, pred, target, queryError=True):
		if isinstance(target, gpuarray.GPUArray) and isinstance(pred, gpuarray.GPUArray):
			assert pred.shape[0] == target.shape[0]

		self.checkDataShape(pred, target)
		self.reset()

		self.grad = self.calcGrad(pred, target)
		self.calcError(pred, target)
		self.dirty = True

		self.updateState(self.getBatchsize(pred))

		if queryError:
			self.error = self.getError()
			return self.error, self.grad
		return self.grad

	def calcError(self, pred, target):
		raise NotImplementedError()

	def calcGrad(self, pred, target):
		raise NotImplementedError()

	def validate(self, pred, target):
		if isinstance(target, gpuarray.GPUArray) and isinstance(pred, gpuarray.GPUArray):
			assert pred.shape[0] == target.shape[0]

		self.checkValDataShape(pred, target)
		self.valError = self.calcVal(pred, target)
		return self.valError

	def calcVal(self, pred, target):
		raise NotImplementedError()

	def reset(self):
		self.error, self.valError, self.grad = None, None, None

	def checkDataShape(self, pred, target):
		pass

	def checkValDataShape(self, pred, target):
		pass

	def getBatchsize(self, pred):
		return pred.shape[0]


class CrossEntropy(Cost):
	def __init__(self, maxlabels=None, weights=None):
		super().__init__()
		self.maxlabels = maxlabels
		self.mostProb = None

		if isinstance(weights, np.ndarray):
			weights = gpuarray.to_gpu(weights)
		self.weights = weights

	def calcGrad(self, scores, labels):
		self.devErr, grad = backend().costmod.crossEntropy(scores, labels, self.weights, self.devErr, memoryPool())
		return grad

	def calcError(self, scores, labels):
		backend().toVectorAddVectorKer(np.float32)(self.accumErr, self.devErr, 1.0)

	def calcVal(self, scores, labels):
		bnd = backend()
		if scores.ndim == 2:
			self.mostProb = bnd.matmod.argmax(scores, axis=1, allocator=memoryPool())
		else:
			flat = scores.reshape(*scores.shape[:2], int(np.prod(scores.shape[2:])))
			self.mostProb = bnd.matmod.argmax(flat, axis=1, allocator=memoryPool()).reshape(labels.shape)

		calcAccuracy = bnd.getAccuracyKernel("calcAccuracy")
		return calcAccuracy(self.mostProb, labels, allocator=memoryPool()).get() / np.prod(labels.shape)

	def reset(self):
		super().reset()
		self.mostProb = None

	def checkDataShape(self, scores, labels):
		assert scores.ndim > 1 and labels.ndim == scores.ndim - 1
		assert labels.dtype == np.int32

		if scores.ndim > 2:
			assert scores.shape[2:] == labels.shape[1:]
		if self.maxlabels:
			assert scores.shape[1] == self.maxlabels
		if self.weights is not None:
			assert self.weights.shape[0] == scores.shape[1]

	def checkValDataShape(self, scores, labels):
		assert scores.ndim > 1 and labels.ndim == scores.ndim - 1
		assert labels.dtype == np.int32

		if scores.ndim > 2:
			assert scores.shape[2:] == labels.shape[1:]
		if self.maxlabels:
			assert scores.shape[1] == self.maxlabels
