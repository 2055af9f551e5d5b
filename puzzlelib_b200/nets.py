"""Workload definitions of the benchmark configurations, built from the module mirror with the reference's layer
names and hyper-parameters (reference: Models/Nets/ResNet.py:23-121, VGG.py:15-111, LeNet.py:13-33).

Only the topology matters here -- weights are random-initialised (there are no checkpoints to load), so `modelpath`
is not supported.
"""
import string

import numpy as np

from .modules import Sequential, Parallel, Conv2D, BatchNorm2D, Activation, relu, Identity, Replicate, Add, MaxPool2D, \
	AvgPool2D, Flatten, Linear, SoftMax


def _resMini(inmaps, outmaps, size, stride, pad, blockname, mininame, addAct, actInplace, bnInplace, initscheme):
	mods = [
		Conv2D(inmaps, outmaps, size, stride=stride, pad=pad, useBias=False, initscheme=initscheme,
			   name="res%s_branch%s" % (blockname, mininame)),
		BatchNorm2D(outmaps, name="bn%s_branch%s" % (blockname, mininame), inplace=bnInplace)
	]
	if addAct:
		mods.append(Activation(relu, inplace=actInplace, name="res%s_branch%s_relu" % (blockname, mininame)))
	return mods


def _resBlock(net, inmaps, hmaps, stride, blockname, convShortcut, actInplace, bnInplace, initscheme):
	# Caffe-style bottleneck: the stride sits on the first 1x1 and on the projection shortcut (ResNet.py:38-61)
	branch = Sequential()
	for mod in _resMini(inmaps, hmaps, 1, stride, 0, blockname, "2a", True, actInplace, bnInplace, initscheme) + \
			   _resMini(hmaps, hmaps, 3, 1, 1, blockname, "2b", True, actInplace, bnInplace, initscheme) + \
			   _resMini(hmaps, 4 * hmaps, 1, 1, 0, blockname, "2c", False, actInplace, bnInplace, initscheme):
		branch.append(mod)

	shortcut = Sequential()
	if convShortcut:
		for mod in _resMini(inmaps, 4 * hmaps, 1, stride, 0, blockname, "1", False, actInplace, bnInplace, initscheme):
			shortcut.append(mod)
	else:
		shortcut.append(Identity())

	net.append(Replicate(2))
	net.append(Parallel().append(branch).append(shortcut))
	net.append(Add())
	net.append(Activation(relu, inplace=actInplace))


def loadResNet(modelpath, layers, actInplace=False, bnInplace=False, initscheme="none", name=None):
	assert modelpath is None, "checkpoint loading is out of scope"

	if layers == "50":
		name = "ResNet-50" if name is None else name
		level3names = ["3%s" % alpha for alpha in string.ascii_lowercase[1:4]]
		level4names = ["4%s" % alpha for alpha in string.ascii_lowercase[1:6]]
	elif layers == "101":
		name = "ResNet-101" if name is None else name
		level3names = ["3b%s" % num for num in range(1, 4)]
		level4names = ["4b%s" % num for num in range(1, 23)]
	elif layers == "152":
		name = "ResNet-152" if name is None else name
		level3names = ["3b%s" % num for num in range(1, 8)]
		level4names = ["4b%s" % num for num in range(1, 36)]
	else:
		raise ValueError("Unsupported ResNet layers mode")

	net = Sequential(name=name)

	net.append(Conv2D(3, 64, 7, stride=2, pad=3, name="conv1", initscheme=initscheme, useBias=False))
	net.append(BatchNorm2D(64, name="bn_conv1", inplace=bnInplace))
	net.append(Activation(relu, inplace=actInplace, name="conv1_relu"))
	net.append(MaxPool2D(3, 2, name="pool1"))

	args = (actInplace, bnInplace, initscheme)
	_resBlock(net, 64, 64, 1, "2a", True, *args)
	_resBlock(net, 256, 64, 1, "2b", False, *args)
	_resBlock(net, 256, 64, 1, "2c", False, *args)

	_resBlock(net, 256, 128, 2, "3a", True, *args)
	for blockname in level3names:
		_resBlock(net, 512, 128, 1, blockname, False, *args)

	_resBlock(net, 512, 256, 2, "4a", True, *args)
	for blockname in level4names:
		_resBlock(net, 1024, 256, 1, blockname, False, *args)

	_resBlock(net, 1024, 512, 2, "5a", True, *args)
	_resBlock(net, 2048, 512, 1, "5b", False, *args)
	_resBlock(net, 2048, 512, 1, "5c", False, *args)

	net.append(AvgPool2D(7, 1))
	net.append(Flatten())
	net.append(Linear(2048, 1000, initscheme=initscheme, name="fc1000"))
	net.append(SoftMax())
	return net


_VGG_CFG = {
	"11": ((64, ), (128, ), (256, 256), (512, 512), (512, 512)),
	"16": ((64, 64), (128, 128), (256, 256, 256), (512, 512, 512), (512, 512, 512)),
	"19": ((64, 64), (128, 128), (256, 256, 256, 256), (512, 512, 512, 512), (512, 512, 512, 512))
}


def loadVGG(modelpath, layers, poolmode="max", initscheme="none", withLinear=True, actInplace=False, name=None):
	assert modelpath is None, "checkpoint loading is out of scope"

	if poolmode not in ("avg", "max"):
		raise ValueError("Unsupported pool mode")
	if layers not in _VGG_CFG:
		raise ValueError("Unsupported VGG layers mode")

	pool = AvgPool2D if poolmode == "avg" else MaxPool2D
	net = Sequential(name="VGG_ILSVRC_%s_layers" % layers if name is None else name)

	inmaps = 3
	for stage, widths in enumerate(_VGG_CFG[layers], start=1):
		for idx, outmaps in enumerate(widths, start=1):
			net.append(Conv2D(inmaps, outmaps, 3, pad=1, initscheme=initscheme, name="conv%d_%d" % (stage, idx)))
			net.append(Activation(relu, inplace=actInplace, name="relu%d_%d" % (stage, idx)))
			inmaps = outmaps
		net.append(pool(2, 2, name="pool%d" % stage))

	if withLinear:
		net.append(Flatten())
		insize = int(np.prod(net.dataShapeFrom((1, 3, 224, 224))))

		net.append(Linear(insize, 4096, initscheme=initscheme, name="fc6"))
		net.append(Activation(relu, inplace=actInplace, name="relu6"))
		net.append(Linear(4096, 4096, initscheme=initscheme, name="fc7"))
		net.append(Activation(relu, inplace=actInplace, name="relu7"))
		net.append(Linear(4096, 1000, initscheme=initscheme, name="fc8"))
		net.append(SoftMax())
	return net


def loadLeNet(modelpath, initscheme="none", name="lenet-5-like"):
	assert modelpath is None, "checkpoint loading is out of scope"
	net = Sequential(name=name)

	net.append(Conv2D(1, 16, 3, initscheme=initscheme))
	net.append(MaxPool2D())
	net.append(Activation(relu))

	net.append(Conv2D(16, 32, 4, initscheme=initscheme))
	net.append(MaxPool2D())
	net.append(Activation(relu))

	net.append(Flatten())
	net.append(Linear(32 * 5 * 5, 1024, initscheme=initscheme))
	net.append(Activation(relu))

	net.append(Linear(1024, 10, initscheme=initscheme))
	return net
