"""Generate tests/golden/ref_cpu_*.npz from the REFERENCE ITSELF.

Runs only in the build container (needs /root/reference).  It imports the unmodified reference package with
`Config.backend = cpu` (its numpy + gcc-JIT CPU backend -- the only backend of the reference that can execute without
a GPU), drives the reference's own Modules on seeded inputs and stores inputs + outputs.  What that backend can do
(SURVEY F5): Conv2D / MaxPool2D / AvgPool2D(pad=0) / BatchNorm2D(eval) forward, Linear forward + backward +
accGradParams, and every Activation forward + backward.  h5py (absent in this image) is stubbed at import.

usage: python tools/gen_golden.py            (writes tests/golden/ref_cpu_ops.npz)
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def importReference():
	root = "/tmp/pzref"
	os.makedirs(root, exist_ok=True)
	link = os.path.join(root, "PuzzleLib")
	if not os.path.exists(link):
		os.symlink(REF, link)
	sys.path.insert(0, root)

	h5py = types.ModuleType("h5py")
	h5py.h5p = types.ModuleType("h5py.h5p")
	h5py.h5f = types.ModuleType("h5py.h5f")
	h5py.File = object
	sys.modules["h5py"], sys.modules["h5py.h5p"], sys.modules["h5py.h5f"] = h5py, h5py.h5p, h5py.h5f

	from PuzzleLib import Config
	Config.backend = Config.Backend.cpu
	Config.showWarnings = False
	return Config


def main():
	importReference()
	from PuzzleLib.Backend import gpuarray
	from PuzzleLib.Modules.Conv2D import Conv2D
	from PuzzleLib.Modules.MaxPool2D import MaxPool2D
	from PuzzleLib.Modules.AvgPool2D import AvgPool2D
	from PuzzleLib.Modules.BatchNorm2D import BatchNorm2D
	from PuzzleLib.Modules.Linear import Linear
	from PuzzleLib.Modules.Activation import Activation

	rng = np.random.RandomState(20261017)
	gold = {}

	# ---- Conv2D forward: (N, C, H, W, K, size, stride, pad, dilation, bias)
	convcases = [(2, 3, 9, 9, 4, 3, 1, 1, 1, True), (2, 4, 12, 10, 6, 3, 2, 1, 1, False), (1, 2, 11, 11, 3, 3, 1, 2, 2, True),
				 (3, 5, 8, 8, 7, 1, 1, 0, 1, False), (2, 3, 16, 16, 8, 7, 2, 3, 1, False)]
	for i, (N, C, H, W, K, size, stride, pad, dil, bias) in enumerate(convcases):
		mod = Conv2D(C, K, size, stride=stride, pad=pad, dilation=dil, useBias=bias)
		w = rng.randn(*mod.W.shape).astype(np.float32)
		mod.W.set(w)
		if bias:
			b = rng.randn(*mod.b.shape).astype(np.float32)
			mod.b.set(b)
			gold["conv%d_b" % i] = b
		x = rng.randn(N, C, H, W).astype(np.float32)
		y = mod(gpuarray.to_gpu(x)).get()
		gold["conv%d_cfg" % i] = np.array([size, stride, pad, dil, int(bias)])
		gold["conv%d_x" % i], gold["conv%d_w" % i], gold["conv%d_y" % i] = x, w, y

	# ---- pooling forward: (N, C, H, W, size, stride, pad, kind)
	poolcases = [(2, 3, 8, 8, 2, 2, 0, "max"), (2, 2, 9, 9, 3, 2, 0, "max"), (2, 2, 8, 8, 3, 2, 1, "max"), (2, 3, 8, 8, 2, 2, 0, "avg"),
				 (1, 2, 7, 7, 7, 1, 0, "avg")]
	for i, (N, C, H, W, size, stride, pad, kind) in enumerate(poolcases):
		mod = (MaxPool2D if kind == "max" else AvgPool2D)(size, stride, pad)
		x = rng.randn(N, C, H, W).astype(np.float32)
		y = mod(gpuarray.to_gpu(x)).get()
		gold["pool%d_cfg" % i] = np.array([size, stride, pad, 0 if kind == "max" else 1])
		gold["pool%d_x" % i], gold["pool%d_y" % i] = x, y

	# ---- BatchNorm2D inference
	for i, (N, C, H, W) in enumerate([(4, 5, 3, 3), (2, 8, 6, 5)]):
		mod = BatchNorm2D(C)
		mod.evalMode()
		scale, bias = rng.randn(1, C, 1, 1).astype(np.float32), rng.randn(1, C, 1, 1).astype(np.float32)
		mean, var = rng.randn(1, C, 1, 1).astype(np.float32), (1.0 + rng.randn(1, C, 1, 1) ** 2).astype(np.float32)
		mod.scale.set(scale); mod.bias.set(bias); mod.mean.set(mean); mod.var.set(var)
		x = rng.randn(N, C, H, W).astype(np.float32)
		y = mod(gpuarray.to_gpu(x)).get()
		for key, val in (("x", x), ("scale", scale), ("bias", bias), ("mean", mean), ("var", var), ("y", y)):
			gold["bn%d_%s" % (i, key)] = val

	# ---- Linear forward / backward / accGradParams
	for i, (B, I, O, transpose) in enumerate([(5, 7, 3, False), (4, 6, 9, True)]):
		# a transposed Linear sizes its bias by `insize` (reference quirk, Linear.py:27) -> only usable without bias
		mod = Linear(I, O, transpose=transpose, useBias=not transpose)
		w = rng.randn(*mod.W.shape).astype(np.float32)
		mod.W.set(w)
		b = rng.randn(O).astype(np.float32) if not transpose else np.zeros(O, dtype=np.float32)
		if not transpose:
			mod.b.set(b)
		x, g = rng.randn(B, I).astype(np.float32), rng.randn(B, O).astype(np.float32)
		y = mod(gpuarray.to_gpu(x)).get()
		mod.backward(gpuarray.to_gpu(g))
		for key, val in (("w", w), ("b", b), ("x", x), ("g", g), ("y", y), ("dx", mod.grad.get()),
						 ("dw", mod.vars["W"].grad.get()),
						 ("db", mod.vars["b"].grad.get() if not transpose else np.zeros(O, dtype=np.float32))):
			gold["lin%d_%s" % (i, key)] = val
		gold["lin%d_transpose" % i] = np.array([int(transpose)])

	# ---- activations forward / backward (fp32 math in gcc-compiled C, CPU/Kernels/ElementWise.py)
	x = (rng.randn(4, 33) * 2.0).astype(np.float32)
	g = rng.randn(4, 33).astype(np.float32)
	gold["act_x"], gold["act_g"] = x, g
	for kind in ("sigmoid", "tanh", "relu", "leakyRelu", "elu", "softPlus", "clip"):
		mod = Activation(kind)
		y = mod(gpuarray.to_gpu(x)).get()
		mod.backward(gpuarray.to_gpu(g))
		gold["act_%s_y" % kind], gold["act_%s_dx" % kind] = y, mod.grad.get()

	os.makedirs(OUT, exist_ok=True)
	path = os.path.join(OUT, "ref_cpu_ops.npz")
	np.savez_compressed(path, **gold)
	print("wrote %s: %d arrays, %d bytes" % (path, len(gold), os.path.getsize(path)))


if __name__ == "__main__":
	main()
