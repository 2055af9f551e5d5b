"""Time one training step of a reference model through the reference's OWN Python layers with either backend underneath:

    --impl ref    PuzzleLib's cuDNN 9 / cuBLAS 12 / NVRTC backend (baseline/_ref, built by baseline/build_ref.py)
    --impl b200   this repository behind the Cuda/Backend.py seam (eager: no CUDA graph, the reference has none)

Step = optimizer.zeroGradParams + net(data) + net.backward(grad) + MomentumSGD.update + net.reset -- the same step
bench.py times.  Timed with device synchronisation on both sides (the reference runs everything on the legacy stream
and is host-driven, so host time IS part of its step).  Prints one JSON line.
"""
import argparse, json, os, sys, time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument("--impl", choices=("ref", "b200"), required=True)
	ap.add_argument("--model", default="resnet50", choices=("resnet50", "vgg16", "lenet"))
	ap.add_argument("--batch", type=int, default=64)
	ap.add_argument("--steps", type=int, default=20)
	ap.add_argument("--warmup", type=int, default=10)
	ap.add_argument("--dtype", default="f32", choices=("f32", "f16"))
	ap.add_argument("--forward-only", action="store_true")
	args = ap.parse_args()

	refroot = os.path.join(ROOT, "baseline", "_ref")
	if args.impl == "b200":
		from puzzlelib_b200 import seam
		seam.install(refroot)
	else:
		sys.path.insert(0, refroot)
		sys.path.append(os.path.join(refroot, "stubs"))

	from PuzzleLib import Config
	Config.showWarnings = False
	from PuzzleLib.Backend import gpuarray
	from PuzzleLib.Optimizers.MomentumSGD import MomentumSGD

	np.random.seed(1234)
	if args.model == "resnet50":
		from PuzzleLib.Models.Nets.ResNet import loadResNet
		net, shape, nout = loadResNet(None, "50", initscheme="he"), (args.batch, 3, 224, 224), 1000
	elif args.model == "vgg16":
		from PuzzleLib.Models.Nets.VGG import loadVGG
		net, shape, nout = loadVGG(None, "16", initscheme="he"), (args.batch, 3, 224, 224), 1000
	else:
		from PuzzleLib.Models.Nets.LeNet import loadLeNet
		net, shape, nout = loadLeNet(None, initscheme="he"), (args.batch, 1, 28, 28), 10

	dtype = np.float32 if args.dtype == "f32" else np.float16
	if args.dtype != "f32":
		net.calcMode(dtype)

	optimizer = MomentumSGD(learnRate=1e-3, momRate=0.9)
	optimizer.setupOn(net, useGlobalState=True)

	rng = np.random.RandomState(1234)
	data = gpuarray.to_gpu(rng.randn(*shape).astype(dtype))
	grad = gpuarray.to_gpu((rng.randn(args.batch, nout) * 1e-3).astype(dtype))
	sync = gpuarray.backend.Driver.Device.synchronize if hasattr(gpuarray.backend.Driver, "Device") else None

	def step():
		if args.forward_only:
			net(data)
			net.reset()
			return
		optimizer.zeroGradParams()
		net(data)
		net.backward(grad)
		optimizer.update()
		net.reset()

	for _ in range(args.warmup):
		step()
	sync()
	t0 = time.perf_counter()
	for _ in range(args.steps):
		step()
	host = time.perf_counter() - t0
	sync()
	dt = (time.perf_counter() - t0) / args.steps

	print(json.dumps({
		"impl": args.impl, "backend": type(gpuarray.backend).__name__, "device": gpuarray.getDeviceName(), "model": args.model,
		"dtype": args.dtype, "batch": args.batch, "steps": args.steps, "warmup": args.warmup, "forward_only": args.forward_only,
		"ms_per_step": dt * 1e3, "images_per_s": args.batch / dt, "host_enqueue_ms_per_step": host / args.steps * 1e3,
		"api": "reference Modules / Containers / Optimizers, eager, one rank",
	}), flush=True)


if __name__ == "__main__":
	main()
