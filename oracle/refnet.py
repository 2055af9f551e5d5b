"""CPU oracle, network level: numpy layers with the reference's forward / backward protocol, assembled into the
benchmark nets (ResNet-50, VGG-16, LeNet-like) exactly as the reference builds them.

TEST INFRASTRUCTURE ONLY (see oracle/ops.py).  Used (1) by tests to check a whole GPU forward+backward pass layer by
layer against the per-op oracle on identical weights and inputs, and (2) by bench.py as the timed CPU baseline
(`cpu_baseline.kind = "port"`: the reference's own numpy CPU backend cannot run any conv / pool / batch-norm backward
pass, SURVEY F5, so the CPU arm is this restatement, float32, numpy BLAS on all host cores).

Protocol mirrored from the reference: Module.__call__ / backward(grad, scale, momentum) with parameter gradients
ACCUMULATED as grad = scale * new + momentum * grad (Sequential.backward default momentum = 1.0,
Containers/Sequential.py:212-234); topology from Models/Nets/ResNet.py:23-121, VGG.py:15-111, LeNet.py:13-33.
"""
import string

import numpy as np

from . import ops


class Layer:
	params = ()

	def __init__(self, name=None):
		self.name = name
		self.x = self.y = self.dx = None

	def forward(self, x, dtype):
		raise NotImplementedError

	def backward(self, g, dtype, scale=1.0, momentum=1.0):
		raise NotImplementedError

	def leaves(self):
		yield self


class Conv(Layer):
	params = ("W", "b")

	def __init__(self, inmaps, outmaps, size, stride=1, pad=0, bias=True, name=None):
		super().__init__(name)
		self.stride, self.pad = stride, pad
		self.W = np.zeros((outmaps, inmaps, size, size), np.float32)
		self.b = np.zeros((outmaps, ), np.float32) if bias else None
		self.dW = np.zeros_like(self.W)
		self.db = np.zeros_like(self.b) if bias else None

	def forward(self, x, dtype):
		self.x = x
		self.y = ops.conv2d(x, self.W, self.b, self.stride, self.pad, dtype=dtype)
		return self.y

	def backward(self, g, dtype, scale=1.0, momentum=1.0):
		self.dx = ops.conv2d_bwd_data(g, self.W, self.x.shape, stride=self.stride, pad=self.pad, dtype=dtype)
		res = ops.conv2d_bwd_params(self.x, g, self.W.shape, self.stride, self.pad, withbias=self.b is not None,
									wgrad=self.dW, bgrad=self.db, scale=scale, momentum=momentum, dtype=dtype)
		if self.b is not None:
			self.dW, self.db = res
		else:
			self.dW = res
		return self.dx


class BatchNorm(Layer):
	params = ("scale", "bias")

	def __init__(self, maps, epsilon=1e-5, name=None):
		super().__init__(name)
		self.epsilon = epsilon
		self.scale, self.bias = np.ones(maps, np.float32), np.zeros(maps, np.float32)
		self.mean, self.var = np.zeros(maps, np.float32), np.ones(maps, np.float32)
		self.dscale, self.dbias = np.zeros(maps, np.float32), np.zeros(maps, np.float32)
		self.numOfProps = 0

	def forward(self, x, dtype):
		self.x = x
		self.numOfProps += 1
		factor = max(1.0 / self.numOfProps, 0.1)
		self.y, self.savemean, self.saveinvvar, self.mean, self.var = ops.batchnorm_train(
			x, self.scale, self.bias, self.mean, self.var, self.epsilon, factor, dtype=dtype
		)
		return self.y

	def backward(self, g, dtype, scale=1.0, momentum=1.0):
		self.dx, dscale, dbias = ops.batchnorm_bwd(self.x, g, self.scale, self.savemean, self.saveinvvar, dtype=dtype)
		self.dscale = scale * dscale + momentum * self.dscale
		self.dbias = scale * dbias + momentum * self.dbias
		return self.dx


class Relu(Layer):
	def forward(self, x, dtype):
		self.y = ops.activation("relu", x, dtype=dtype)
		return self.y

	def backward(self, g, dtype, scale=1.0, momentum=1.0):
		self.dx = ops.activation_bwd("relu", g, self.y, dtype=dtype)
		return self.dx


class MaxPool(Layer):
	def __init__(self, size=2, stride=2, pad=0, name=None):
		super().__init__(name)
		self.size, self.stride, self.pad = size, stride, pad

	def forward(self, x, dtype):
		self.x = x
		self.y = ops.pool2d(x, self.size, self.stride, self.pad, "max", dtype=dtype)
		return self.y

	def backward(self, g, dtype, scale=1.0, momentum=1.0):
		self.dx = ops.pool2d_bwd(self.x, self.y, g, self.size, self.stride, self.pad, "max", dtype=dtype)
		return self.dx


class AvgPool(MaxPool):
	def forward(self, x, dtype):
		self.x = x
		self.y = ops.pool2d(x, self.size, self.stride, self.pad, "avgWithPad", dtype=dtype)
		return self.y

	def backward(self, g, dtype, scale=1.0, momentum=1.0):
		self.dx = ops.pool2d_bwd(self.x, self.y, g, self.size, self.stride, self.pad, "avgWithPad", dtype=dtype)
		return self.dx


class Flatten(Layer):
	def forward(self, x, dtype):
		self.inshape = x.shape
		self.y = x.reshape(x.shape[0], -1)
		return self.y

	def backward(self, g, dtype, scale=1.0, momentum=1.0):
		self.dx = g.reshape(self.inshape)
		return self.dx


class Linear(Layer):
	params = ("W", "b")

	def __init__(self, insize, outsize, name=None):
		super().__init__(name)
		self.W, self.b = np.zeros((insize, outsize), np.float32), np.zeros(outsize, np.float32)
		self.dW, self.db = np.zeros_like(self.W), np.zeros_like(self.b)

	def forward(self, x, dtype):
		self.x = x
		self.y = ops.gemm(x, self.W, dtype=dtype) + self.b.astype(dtype)
		return self.y

	def backward(self, g, dtype, scale=1.0, momentum=1.0):
		self.dx = ops.gemm(g, self.W, transpB=True, dtype=dtype)
		self.dW = ops.gemm(self.x, g, out=self.dW, transpA=True, alpha=scale, beta=momentum, dtype=dtype)
		self.db = ops.matsum(g, 0, out=self.db, alpha=scale, beta=momentum, dtype=dtype)
		return self.dx


class SoftMax(Layer):
	def forward(self, x, dtype):
		self.y = ops.softmax(x.reshape(x.shape + (1, 1)), dtype=dtype).reshape(x.shape)
		return self.y

	def backward(self, g, dtype, scale=1.0, momentum=1.0):
		shape = g.shape + (1, 1)
		self.dx = ops.softmax_bwd(self.y.reshape(shape), g.reshape(shape), dtype=dtype).reshape(g.shape)
		return self.dx


class Seq(Layer):
	def __init__(self, layers=(), name=None):
		super().__init__(name)
		self.layers = list(layers)

	def append(self, layer):
		self.layers.append(layer)
		return self

	def forward(self, x, dtype):
		for layer in self.layers:
			x = layer.forward(x, dtype)
		self.y = x
		return x

	def backward(self, g, dtype, scale=1.0, momentum=1.0):
		for layer in reversed(self.layers):
			g = layer.backward(g, dtype, scale, momentum)
		self.dx = g
		return g

	def leaves(self):
		for layer in self.layers:
			yield from layer.leaves()


class Residual(Layer):
	"""Replicate(2) -> Parallel(branch, shortcut) -> Add (reference: ResNet.py:54-59); the trailing ReLU is a
	separate leaf like in the reference graph"""

	def __init__(self, branch, shortcut, name=None):
		super().__init__(name)
		self.branch, self.shortcut = branch, shortcut

	def forward(self, x, dtype):
		a = self.branch.forward(x, dtype)
		b = self.shortcut.forward(x, dtype) if self.shortcut is not None else x
		self.y = (0.0 + a) + b
		return self.y

	def backward(self, g, dtype, scale=1.0, momentum=1.0):
		ga = self.branch.backward(g, dtype, scale, momentum)
		gb = self.shortcut.backward(g, dtype, scale, momentum) if self.shortcut is not None else g
		self.dx = (0.0 + ga) + gb
		return self.dx

	def leaves(self):
		yield from self.branch.leaves()
		if self.shortcut is not None:
			yield from self.shortcut.leaves()


def _mini(inmaps, outmaps, size, stride, pad, blockname, mininame, act):
	layers = [Conv(inmaps, outmaps, size, stride, pad, bias=False, name="res%s_branch%s" % (blockname, mininame)),
			  BatchNorm(outmaps, name="bn%s_branch%s" % (blockname, mininame))]
	if act:
		layers.append(Relu(name="res%s_branch%s_relu" % (blockname, mininame)))
	return layers


def _block(net, inmaps, hmaps, stride, blockname, convShortcut):
	branch = Seq(_mini(inmaps, hmaps, 1, stride, 0, blockname, "2a", True) + _mini(hmaps, hmaps, 3, 1, 1, blockname, "2b", True) +
				 _mini(hmaps, 4 * hmaps, 1, 1, 0, blockname, "2c", False))
	shortcut = Seq(_mini(inmaps, 4 * hmaps, 1, stride, 0, blockname, "1", False)) if convShortcut else None
	net.append(Residual(branch, shortcut, name="res%s" % blockname))
	net.append(Relu(name="res%s_relu" % blockname))


def resnet50():
	net = Seq(name="ResNet-50")
	net.append(Conv(3, 64, 7, 2, 3, bias=False, name="conv1")).append(BatchNorm(64, name="bn_conv1")).append(Relu(name="conv1_relu"))
	net.append(MaxPool(3, 2, name="pool1"))

	_block(net, 64, 64, 1, "2a", True)
	_block(net, 256, 64, 1, "2b", False)
	_block(net, 256, 64, 1, "2c", False)
	_block(net, 256, 128, 2, "3a", True)
	for alpha in string.ascii_lowercase[1:4]:
		_block(net, 512, 128, 1, "3%s" % alpha, False)
	_block(net, 512, 256, 2, "4a", True)
	for alpha in string.ascii_lowercase[1:6]:
		_block(net, 1024, 256, 1, "4%s" % alpha, False)
	_block(net, 1024, 512, 2, "5a", True)
	_block(net, 2048, 512, 1, "5b", False)
	_block(net, 2048, 512, 1, "5c", False)

	net.append(AvgPool(7, 1, name="pool5")).append(Flatten()).append(Linear(2048, 1000, name="fc1000")).append(SoftMax())
	return net


def vgg16():
	cfg = ((64, 64), (128, 128), (256, 256, 256), (512, 512, 512), (512, 512, 512))
	net, inmaps = Seq(name="VGG_ILSVRC_16_layers"), 3
	for stage, widths in enumerate(cfg, start=1):
		for idx, outmaps in enumerate(widths, start=1):
			net.append(Conv(inmaps, outmaps, 3, 1, 1, name="conv%d_%d" % (stage, idx))).append(Relu(name="relu%d_%d" % (stage, idx)))
			inmaps = outmaps
		net.append(MaxPool(2, 2, name="pool%d" % stage))
	net.append(Flatten())
	net.append(Linear(512 * 7 * 7, 4096, name="fc6")).append(Relu(name="relu6"))
	net.append(Linear(4096, 4096, name="fc7")).append(Relu(name="relu7"))
	net.append(Linear(4096, 1000, name="fc8")).append(SoftMax())
	return net


def lenet():
	net = Seq(name="lenet-5-like")
	net.append(Conv(1, 16, 3)).append(MaxPool()).append(Relu())
	net.append(Conv(16, 32, 4)).append(MaxPool()).append(Relu())
	net.append(Flatten()).append(Linear(800, 1024)).append(Relu()).append(Linear(1024, 10))
	return net


def init_he(net, seed=1234):
	"""He-normal weights, BN scale ~ N(1, 0.01): the initialisation the benchmark uses for both arms"""
	rng = np.random.RandomState(seed)
	for layer in net.leaves():
		if isinstance(layer, Conv):
			fan = layer.W.shape[1] * layer.W.shape[2] * layer.W.shape[3]
			layer.W = rng.normal(0.0, np.sqrt(2.0 / fan), layer.W.shape).astype(np.float32)
			if layer.b is not None:
				layer.b = rng.normal(0.0, 0.01, layer.b.shape).astype(np.float32)
		elif isinstance(layer, Linear):
			layer.W = rng.normal(0.0, np.sqrt(2.0 / layer.W.shape[0]), layer.W.shape).astype(np.float32)
			layer.b = rng.normal(0.0, 0.01, layer.b.shape).astype(np.float32)
		elif isinstance(layer, BatchNorm):
			layer.scale = rng.normal(1.0, 0.01, layer.scale.shape).astype(np.float32)
			layer.bias = rng.normal(0.0, 0.01, layer.bias.shape).astype(np.float32)
	return net
