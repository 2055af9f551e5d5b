"""Optimizers on the hot path: `Optimizer` with the reference's local / GLOBAL-state modes, SGD, MomentumSGD, NesterovSGD, Adam.

Global-state mode is the data-parallel hook: all parameters (and all gradients) of one dtype are re-homed into ONE flat
`SharedArray`, module variables become views of it, and `update()` runs `nodeinfo.sumTensor("grad", flatGrad)` followed by
a single update kernel over the flat buffer (reference: Optimizers/Optimizer.py:48-63,66-111,146-170; SGD.py:11-18;
MomentumSGD.py:12-27).  With an NCCL `NodeInfo` and MomentumSGD the all-reduce and the update are issued as one fused
C call (pz_nccl_allreduce_sgd_momentum): the 1/P of the mean is folded into the update kernel, saving a full pass over
the gradient buffer.
"""
import math
from collections import OrderedDict

import numpy as np

from .shim import gpuarray, backend
from .backend import SharedArray
from .modules import Variable
from .driver import lib, check, dtypeCode


class Optimizer:
	def __init__(self, nodeinfo=None):
		self.t = 0
		self.learnRate = 0.0
		self.attrs = {"t", "learnRate"}

		self.module = None
		self.states = {}
		self.hooks = []

		self.shParams, self.shGrads = {}, {}
		self.globalState = False
		self.globalVar = OrderedDict()

		self.customVars = []
		self.nodeinfo = nodeinfo

	def setAttr(self, name, attr):
		setattr(self, name, attr)
		self.attrs.add(name)

	def getAttrDict(self):
		return {name: getattr(self, name) for name in self.attrs}

	def addHook(self, hook):
		self.hooks.append(hook)

	def setupOn(self, mod, useGlobalState=False):
		if self.nodeinfo is not None:
			assert useGlobalState

		self.module = mod
		vartable = self.module.getVarTable()

		if useGlobalState:
			self.globalState = True
			self.setupGlobalState(vartable)
		else:
			self.setupLocalStates(vartable)

		if self.nodeinfo is not None:
			assert len(self.customVars) == 0

	def setupGlobalState(self, vartable):
		variables = sorted(((names, var) for var, names in vartable.items()), key=lambda elem: elem[0][0])

		for names, var in variables:
			if var.hasUpdater:
				assert self.nodeinfo is None
				self.customVars.append(names[0])
				continue

			dtype = var.data.dtype
			shParams = self.shParams.setdefault(dtype, SharedArray(dtype))
			shGrads = self.shGrads.setdefault(dtype, SharedArray(dtype))

			shParams.register(var.data.shape, dtype, names[0])
			shGrads.register(var.grad.shape, dtype, names[0])

		for shParams, shGrads in zip(self.shParams.values(), self.shGrads.values()):
			shParams.build()
			shGrads.build()
			# alignment gaps between blocks must hold finite values: they ride through the all-reduce and the update
			shParams.ary.fill(0)
			shGrads.ary.fill(0)
			self.globalVar[shParams.dtype] = Variable(shParams.ary, grad=shGrads.ary)

		for names, var in variables:
			if var.hasUpdater:
				continue

			dtype = var.data.dtype
			data, grad = self.shParams[dtype][names[0]], self.shGrads[dtype][names[0]]
			data.set(var.data)
			grad.set(var.grad)

			for name in names:
				self.module.setVar(name, Variable(data, grad=grad))

		for dtype, globalVar in self.globalVar.items():
			if self.nodeinfo is not None:
				self.nodeinfo.broadcastBuffer("data", globalVar.data.gpudata)
			self.states[dtype] = self.setupState(globalVar)

	def setupLocalStates(self, vartable):
		for var, names in vartable.items():
			if var.hasUpdater:
				self.customVars.append(names[0])
				continue
			self.states[names[0]] = self.setupState(var)

	def zeroGradParams(self):
		if self.globalState:
			for globalVar in self.globalVar.values():
				globalVar.grad.fill(0)
		else:
			for name in self.states:
				var = self.module.getVar(name)
				if not var.hasUpdater:
					var.grad.fill(0)

	def setupState(self, var):
		return {}

	def update(self, useStreams=False, sync=True):
		self.t += 1

		if self.globalState:
			self.updateGlobalState()
		else:
			self.updateLocalStates()

		for name in self.customVars:
			var = self.module.getVar(name)
			var.update(self.learnRate)

	def updateGlobalState(self):
		for dtype, globalVar in self.globalVar.items():
			state = self.states[dtype]

			for hook in self.hooks:
				hook(globalVar, state)

			if self.nodeinfo is not None:
				self.nodeinfo.sumTensor("grad", globalVar.grad)

			if globalVar.learnRate > 0.0:
				self.updateVar(globalVar, state)

	def updateLocalStates(self):
		for name, state in self.states.items():
			var = self.module.getVar(name)
			assert var.grad is not None and var.data.shape == var.grad.shape

			for hook in self.hooks:
				hook(var, state, None)

			if var.learnRate > 0.0:
				self.updateVar(var, state)

	def updateVar(self, var, state, stream=None):
		raise NotImplementedError()


class SGD(Optimizer):
	def __init__(self, learnRate=1e-3, nodeinfo=None):
		super().__init__(nodeinfo)
		self.setAttr("learnRate", learnRate)

	def updateVar(self, var, state, stream=None):
		backend().toVectorAddVectorKer(var.data.dtype)(var.data, var.grad, self.learnRate * var.learnRate)


class NesterovSGD(SGD):
	"""reference: Optimizers/NesterovSGD.py:12-27"""

	def __init__(self, learnRate=1e-3, momRate=0.9, nodeinfo=None):
		super().__init__(learnRate, nodeinfo)
		self.momRate = None
		self.setAttr("momRate", momRate)

	def setupState(self, var):
		return {"mom": gpuarray.zeros(var.data.shape, dtype=var.data.dtype)}

	def updateVar(self, var, state, stream=None):
		backend().nesterovMomSGDKer(var.data.dtype)(
			var.data, var.grad, state["mom"], self.learnRate * var.learnRate, self.momRate * var.momRate
		)


class Adam(Optimizer):
	"""reference: Optimizers/Adam.py:14-45 (bias correction folded into the step size, fp32 moments)"""

	def __init__(self, alpha=1e-3, beta1=0.9, beta2=0.999, epsilon=1e-8, nodeinfo=None):
		super().__init__(nodeinfo)
		self.alpha, self.beta1, self.beta2, self.epsilon = None, None, None, None
		self.setAttr("alpha", alpha)
		self.setAttr("beta1", beta1)
		self.setAttr("beta2", beta2)
		self.setAttr("epsilon", epsilon)

	def setupState(self, var):
		return {"mg": gpuarray.zeros(var.data.shape, dtype=np.float32), "ms": gpuarray.zeros(var.data.shape, dtype=np.float32)}

	def updateVar(self, var, state, stream=None):
		fix1, fix2 = 1.0 - self.beta1 ** self.t, 1.0 - self.beta2 ** self.t
		self.learnRate = self.alpha * math.sqrt(fix2) / fix1
		backend().adamKer(var.data.dtype)(
			var.data, var.grad, state["mg"], state["ms"], self.learnRate * var.learnRate, 1.0 - self.beta1, 1.0 - self.beta2,
			self.epsilon
		)


class MomentumSGD(SGD):
	def __init__(self, learnRate=1e-3, momRate=0.9, nodeinfo=None):
		super().__init__(learnRate, nodeinfo)
		self.momRate = None
		self.setAttr("momRate", momRate)

	def setupState(self, var):
		return {"mom": gpuarray.zeros(var.data.shape, dtype=var.data.dtype)}

	def updateVar(self, var, state, stream=None):
		backend().classicMomSGDKer(var.data.dtype)(
			var.data, var.grad, state["mom"], self.learnRate * var.learnRate, self.momRate * var.momRate
		)

	def updateGlobalState(self):
		fused = getattr(self.nodeinfo, "sumTensorAndMomentumSGD", None)
		if fused is None or self.hooks:
			return super().updateGlobalState()

		# all-reduce + mean + momentum update of the flat buffer as one call (Optimizer.py:166-170 + MomentumSGD.py:24-27)
		for dtype, globalVar in self.globalVar.items():
			if globalVar.learnRate > 0.0:
				fused(globalVar.data, globalVar.grad, self.states[dtype]["mom"], self.learnRate * globalVar.learnRate,
					  self.momRate * globalVar.momRate)
			else:
				self.nodeinfo.sumTensor("grad", globalVar.grad)
