"""The reference's OWN numpy CPU backend (Config.backend = cpu) on the forward pass it can run: LeNet N=64 and ResNet-50 up to
fc1000 in evalMode (its CPU backend has no conv / pool / batch-norm backward and no softmax -- SURVEY F5).
Runs the unmodified tree under baseline/_ref in its own process (the backend choice is import-time global).  One JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--batch", type=int, default=8)
	ap.add_argument("--reps", type=int, default=3)
	args = ap.parse_args()

	ref = os.path.join(ROOT, "baseline", "_ref")
	sys.path.insert(0, ref)
	sys.path.append(os.path.join(ref, "stubs"))
	import numpy as np
	from PuzzleLib import Config
	Config.backend = Config.Backend.cpu
	Config.showWarnings = False
	from PuzzleLib.Backend import gpuarray
	from PuzzleLib.Models.Nets.LeNet import loadLeNet
	from PuzzleLib.Models.Nets.ResNet import loadResNet

	np.random.seed(1234)

	def median(fn):
		fn()                                  # first call: gcc JIT of CPU/Kernels/ElementWise.py
		times = []
		for _ in range(args.reps):
			t0 = time.perf_counter()
			fn()
			times.append(time.perf_counter() - t0)
		return sorted(times)[len(times) // 2]

	lenet = loadLeNet(None, initscheme=None)
	lenet.evalMode()
	xl = gpuarray.to_gpu(np.random.randn(64, 1, 28, 28).astype(np.float32))
	tl = median(lambda: lenet(xl))

	resnet = loadResNet(None, layers="50", initscheme=None)
	resnet.evalMode()
	xr = gpuarray.to_gpu(np.random.randn(args.batch, 3, 224, 224).astype(np.float32))

	def forward():
		data = xr
		for mod in resnet.graph[:-1]:         # up to fc1000: no softmax on this backend
			data = mod(data)
		return data

	tr = median(forward)
	print(json.dumps({
		"backend": "PuzzleLib numpy CPU backend (Config.Backend.cpu), unmodified, forward only",
		"cores": os.cpu_count(),
		"lenet_n64_forward_ms": tl * 1e3,
		"resnet50_forward_images_per_s": args.batch / tr,
		"resnet50_forward_batch": args.batch, "resnet50_forward_s": tr,
	}), flush=True)


if __name__ == "__main__":
	main()
