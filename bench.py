"""bench.py -- ResNet-50 fp32 forward+backward images/s on synthetic Nx3x224x224 (BASELINE.json configs[1]).

  python bench.py --gpus 1 --steps K --warmup W                      this repo's CUDA path (one JSON line)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   data parallel, one rank per GPU
  python bench.py --impl reference ...                               the CPU arm (oracle port on the host cores)

The model, the containers, the optimizer and their whole Python call path are the REFERENCE's own
(`PuzzleLib.Models.Nets.ResNet.loadResNet`, `Containers.Sequential / Parallel`, `Optimizers.MomentumSGD`, `Backend/*.py`),
imported unmodified from baseline/_ref with this repository behind the `Cuda/Backend.py` seam (puzzlelib_b200/seam.py).

One "step" = zeroGradParams + net(data) + net.backward(grad) + gradient sync (mean over ranks, NCCL) + momentum-SGD
update, i.e. one Trainer.handleBatch of the reference with the cost replaced by a fixed synthetic output gradient
(SURVEY 8d C2/C4).  `value` times K steps with the input batch already resident in HBM; `e2e` repeats the measurement
with the batch copied from pinned host memory every step and the softmax output read back to the host every step.
At N=1 the line also carries `reference_gpu`: the same step through the reference's OWN cuDNN / cuBLAS / NVRTC backend
on the same GPU (tools/bench_ref_cuda.py --impl ref, a separate process: the backend is an import-time global).
torch is used only as the rendezvous (gloo) between ranks; all device work goes through libpzb200.so.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)

METRIC = "ResNet-50 fp32 fwd+bwd images/sec"
BATCH = 64
FLOP_PER_IMAGE = 23.01e9          # 3 x (490.53 GFLOP conv + 0.262 fc) / 64


def loadPeaks():
	path = os.path.join(ROOT, "MEASURED_PEAKS.json")
	if os.path.exists(path):
		with open(path) as f:
			peaks = json.load(f)
		out = {"hbm": float(peaks["hbm_gbs"]), "bf16": float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])), "src": "measured"}
	else:
		out = {"hbm": 6650.0, "bf16": 1400.0, "src": "fallback"}
	# the TF32 figure is not in MEASURED_PEAKS.json: tools/measure_tf32_peak.py (cuBLAS fp32 GEMM with TF32 math on a B200 of this pool)
	tf32 = os.path.join(ROOT, "profiles", "r02_tf32_peak.json")
	if os.path.exists(tf32):
		with open(tf32) as f:
			out["tf32"] = float(json.load(f)["tf32_tflops_sustained"])
	return out


# ---------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
	FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
			 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

	def __init__(self, device):
		self.device, self.samples, self.proc, self.thread = device, [], None, None

	def start(self):
		try:
			self.proc = subprocess.Popen(
				["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
				stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True
			)
		except OSError:
			return
		self.thread = threading.Thread(target=self._read, daemon=True)
		self.thread.start()

	def _read(self):
		for line in self.proc.stdout:
			parts = [p.strip() for p in line.split(",")]
			if len(parts) >= 7:
				self.samples.append(parts)

	def stop(self):
		if self.proc is None:
			return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
		self.proc.terminate()
		try:
			self.proc.wait(timeout=5)
		except Exception:
			self.proc.kill()

		sm, smmax, reasons = [], [], set()
		for parts in self.samples:
			try:
				sm.append(float(parts[0]))
				smmax.append(float(parts[1]))
			except ValueError:
				continue
			for name, flag in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
				if flag.lower().startswith("active"):
					reasons.add(name)
		return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smmax) if smmax else None,
				"reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- CPU arm
def cpuThreads():
	try:
		from threadpoolctl import threadpool_info
		return max([info.get("num_threads", 1) for info in threadpool_info()] or [os.cpu_count() or 1])
	except Exception:
		return os.cpu_count() or 1


def cpuStep(net, x, gy):
	from oracle import refnet      # the CPU baseline IS the oracle port (cpu_baseline.kind = "port")
	for layer in net.leaves():
		for name in ("dW", "db", "dscale", "dbias"):
			if getattr(layer, name, None) is not None:
				setattr(layer, name, np.zeros_like(getattr(layer, name)))
	net.forward(x, np.float32)
	net.backward(gy, np.float32, 1.0, 1.0)


def cpuBaseline(images, steps, warmup):
	from oracle import refnet
	net = refnet.init_he(refnet.resnet50(), seed=1234)
	rng = np.random.RandomState(1234)
	x = rng.randn(images, 3, 224, 224).astype(np.float32)
	gy = (rng.randn(images, 1000) * 1e-3).astype(np.float32)

	for _ in range(warmup):
		cpuStep(net, x, gy)
	t0 = time.perf_counter()
	for _ in range(steps):
		cpuStep(net, x, gy)
	dt = (time.perf_counter() - t0) / steps
	return images / dt, dt


def referenceArm(args):
	rank = int(os.environ.get("RANK", "0"))
	if rank != 0:
		return

	# bounded sample: calibrate the per-step image count so that (steps + warmup) steps stay within ~150 s
	ips, dt = cpuBaseline(1, 1, 1)
	budget = 150.0 / max(1, args.steps + args.warmup)
	images = int(max(1, min(BATCH, budget * ips * 1.5)))
	ips, dt = cpuBaseline(images, args.steps, args.warmup)
	cores = cpuThreads()

	line = {
		"impl": "reference", "metric": METRIC, "value": ips, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
		"warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
		"dtype": "f32", "data": "synthetic",
		"config": {"workload": "ResNet-50 fp32 fwd+bwd, synthetic %dx3x224x224 (bounded sample of the 64-image batch)" % images,
				   "images_per_step": images, "parallelism": "cpu"},
		"cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
						 "sample": "%d images/step, %d steps: numpy float32 restatement of the reference ops (oracle/refnet.py); "
								   "the reference's own numpy CPU backend has no conv/pool/batch-norm backward (SURVEY F5)" % (images, args.steps)},
		"e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
		"gpu_launches": 0,
	}
	print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------- GPU arm
def referenceGpu(args):
	"""the reference's own CUDA backend (cuDNN / cuBLAS / NVRTC) on the same box, same model / batch / step"""
	cmd = [sys.executable, os.path.join(ROOT, "tools", "bench_ref_cuda.py"), "--impl", "ref", "--model", args.model, "--batch", str(BATCH),
		   "--steps", str(max(3, min(10, args.steps))), "--warmup", "3", "--dtype", "f32" if args.dtype == "f32" else "f16"]
	try:
		out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
		line = [l for l in out.stdout.splitlines() if l.startswith("{")]
		if not line:
			return {"unavailable": (out.stderr.strip().splitlines() or ["no output"])[-1][:200]}
		res = json.loads(line[-1])
		return {"value": res["images_per_s"], "unit": "images/s", "ms_per_step": res["ms_per_step"], "steps": res["steps"],
				"backend": "PuzzleLib Cuda backend (cuDNN %s, cuBLAS, NVRTC) built by baseline/build_ref.py; eager, host-driven" % cudnnVersion(),
				"dtype": res["dtype"]}
	except Exception as e:      # noqa: BLE001
		return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}


def referenceCpuForward():
	"""the reference's own numpy CPU backend on what it can run: the FORWARD pass (LeNet N=64 = BASELINE.json configs[0]; ResNet-50 up to
	fc1000) -- a second, clearly labelled CPU number next to the fwd+bwd oracle port of `cpu_baseline`"""
	cmd = [sys.executable, os.path.join(ROOT, "tools", "bench_ref_cpu.py"), "--batch", "8", "--reps", "3"]
	try:
		out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
		line = [l for l in out.stdout.splitlines() if l.startswith("{")]
		if not line:
			return {"unavailable": (out.stderr.strip().splitlines() or ["no output"])[-1][:200]}
		return json.loads(line[-1])
	except Exception as e:      # noqa: BLE001
		return {"unavailable": "%s: %s" % (type(e).__name__, str(e)[:200])}


def cudnnVersion():
	try:
		import ctypes
		return str(ctypes.CDLL("libcudnn.so.9").cudnnGetVersion())
	except Exception:      # noqa: BLE001
		return "9"


def calcMode16(net, dt):
	"""net.calcMode for a 16-bit type.  float16 is the reference's own path (Containers/Container.py:216-223); bfloat16 is this
	backend's extension through seam.calcMode, which also covers modules whose calcMode only checks a
	{float16, float32} whitelist (Modules/BatchNormND.py:102-110)."""
	from puzzlelib_b200 import seam
	if np.dtype(dt) == np.float16:
		net.calcMode(np.float16)
	else:
		seam.calcMode(net, dt)


def gpuArm(args):
	from puzzlelib_b200 import seam, driver
	from puzzlelib_b200.grid import nodeFromEnvironment

	node = nodeFromEnvironment()
	if node.gridsize != args.gpus:
		raise SystemExit("--gpus %d does not match WORLD_SIZE %d (launch with torch.distributed.run)" % (args.gpus, node.gridsize))

	seam.install(deviceIdx=node.device)            # the reference tree (baseline/_ref) over this backend; Config.deviceIdx = our GPU
	driver.Device(node.device).set()
	node.attach()

	from PuzzleLib import Config
	Config.showWarnings = False
	from PuzzleLib.Backend import gpuarray
	from PuzzleLib.Models.Nets.ResNet import loadResNet
	from PuzzleLib.Optimizers.MomentumSGD import MomentumSGD

	np.random.seed(1234)                       # same initial weights on every rank (and broadcast from rank 0 anyway)
	# the headline is ResNet-50 fp32 (BASELINE.json configs[1]); --model / --dtype / --batch time the other configs for profiles/
	global BATCH, FLOP_PER_IMAGE, METRIC
	if args.batch:
		BATCH = args.batch
	if args.model == "vgg16":
		from PuzzleLib.Models.Nets.VGG import loadVGG
		net = loadVGG(None, "16", initscheme="he")
		FLOP_PER_IMAGE = 92.8e9
	else:
		net = loadResNet(None, "50", initscheme="he")
	dt = np.dtype(np.float32)
	if args.dtype != "f32":
		dt = driver.bfloat16 if args.dtype == "bf16" else np.dtype(np.float16)
		calcMode16(net, dt)
	if args.model != "resnet50" or args.dtype != "f32":
		METRIC = "%s %s fwd+bwd images/sec" % ({"resnet50": "ResNet-50", "vgg16": "VGG-16"}[args.model], args.dtype)
	optimizer = MomentumSGD(learnRate=1e-3, momRate=0.9, nodeinfo=node if node.gridsize > 1 else None)
	optimizer.setupOn(net, useGlobalState=True)

	rng = np.random.RandomState(1234 + node.index)      # every rank draws its own shard of the global batch
	pinned = driver.PinnedBuffer((BATCH, 3, 224, 224), dt)
	pinned.array[...] = rng.randn(BATCH, 3, 224, 224).astype(dt)
	data = gpuarray.to_gpu(pinned.array)
	grad = gpuarray.to_gpu((rng.randn(BATCH, 1000) * 1e-3).astype(dt))
	hostOut = driver.PinnedBuffer((BATCH, 1000), dt)

	def step(e2e=False, asyncCopy=False):
		if e2e:                                                      # H2D of the batch from pinned host memory
			driver.check(driver.lib.pz_memcpy_h2d(data.ptr, pinned.ptr, data.nbytes, None, 1 if asyncCopy else 0))
		optimizer.zeroGradParams()
		out = net(data)
		net.backward(grad)
		optimizer.update()                                           # N > 1: nodeinfo.sumTensor (NCCL mean) + the SGD kernel
		if e2e:                                                      # D2H of the step's result (the softmax output)
			driver.check(driver.lib.pz_memcpy_d2h(hostOut.ptr, out.ptr, out.nbytes, None, 1 if asyncCopy else 0))
		net.reset()                                                  # like Handler.handle: activations go back to the pool

	# e2e with the input pipeline a trainer runs: while step i computes on one device buffer, the batch of step i+1 travels from
	# pinned host memory into the other on a copy stream (every step still pays for one H2D of a full batch and one D2H of its
	# result inside the timed region; they overlap the kernels instead of preceding them).  Two steps = one unit (buffers A, B).
	dataB = gpuarray.empty(data.shape, dtype=dt)
	copyStream = driver.Stream()
	forkEv, joinEv = driver.Event(timing=False), driver.Event(timing=False)

	def stepPrefetch(cur, nxt):
		forkEv.record()                                              # on the stream the step runs on
		copyStream.waitEvent(forkEv)
		driver.check(driver.lib.pz_memcpy_h2d(nxt.ptr, pinned.ptr, nxt.nbytes, copyStream.handle, 1))
		optimizer.zeroGradParams()
		out = net(cur)
		net.backward(grad)
		optimizer.update()
		driver.check(driver.lib.pz_memcpy_d2h(hostOut.ptr, out.ptr, out.nbytes, None, 1))
		joinEv.record(copyStream)
		driver.check(driver.lib.pz_stream_wait_event(driver.currentStream.handle if driver.currentStream is not None else None, joinEv.handle))
		net.reset()

	def stepPair():
		stepPrefetch(data, dataB)
		stepPrefetch(dataB, data)

	hostMs = [0.0]

	def timed(nsteps, e2e=False):
		node.barrier()
		driver.Device.synchronize()
		start, end = driver.Event(), driver.Event()
		launches = driver.launchCount()
		start.record()
		t0 = time.perf_counter()
		for _ in range(nsteps):
			step(e2e)
		hostMs[0] = (time.perf_counter() - t0) * 1e3 / nsteps      # host time to ENQUEUE one step (the GPU runs behind)
		end.record()
		end.synchronize()
		driver.Device.synchronize()
		ms = start.timeTill(end)
		launches = driver.launchCount() - launches
		node.barrier()
		if node.gridsize > 1:
			ms = node.rendezvous.maxValue(ms)                        # device time, max over ranks
		return ms, launches

	def timedGraph(graph, nsteps):
		node.barrier()
		graph.synchronize()
		driver.Device.synchronize()
		start, end = driver.Event(), driver.Event()
		start.record(graph.stream)
		for _ in range(nsteps):
			graph.launch()
		end.record(graph.stream)
		end.synchronize()
		graph.synchronize()
		ms = start.timeTill(end)
		node.barrier()
		if node.gridsize > 1:
			ms = node.rendezvous.maxValue(ms)
		return ms

	if args.profile_run:
		# under ncu (tools/gpu_profile.sh): a few eager steps, nothing else -- every launch of the last step is a row of the launch list
		for _ in range(args.warmup + args.steps):
			step()
		driver.Device.synchronize()
		node.close()
		return

	warmup = max(10, args.warmup)                                # >= 10: the batch-norm running-average factor reaches its floor (0.1)
	for _ in range(warmup):
		step()

	# ---- eager: every operator call goes through the Python module API (Module.__call__ -> Backend -> ctypes -> libpzb200.so)
	eagerMs, launches = timed(args.steps)
	hostEnqueueMs = hostMs[0]
	eagerE2eMs, _ = timed(args.steps, e2e=True)
	launchesPerStep = launches / args.steps

	# ---- graph: the same step (same module calls, same kernels, same buffers) captured once into a CUDA graph and replayed
	graphNote, sampler, clocks = None, None, None
	graphs = []
	try:
		if args.no_graph:
			raise RuntimeError("disabled by --no-graph")
		graph = driver.StepGraph(lambda: step(False), warmup=2)
		graphE2e = driver.StepGraph(stepPair, warmup=2)             # two pipelined steps per replay
		graphs += [graph, graphE2e]
		timedGraph(graph, 3)
		sampler = ClockSampler(node.device) if node.index == 0 else None
		if sampler:
			sampler.start()
		ms = timedGraph(graph, args.steps)
		clocks = sampler.stop() if sampler else None
		pairs = max(1, args.steps // 2)
		msE2e = timedGraph(graphE2e, pairs) * args.steps / (2.0 * pairs)
		launches = int(round(launchesPerStep * args.steps))
		api = "StepGraph replay of the module-API step (driver.StepGraph: capture once, one launch per step)"
	except Exception as e:                                           # noqa: BLE001 -- report eager numbers instead
		graphNote = "graph capture unavailable: %s" % (str(e).splitlines()[0] if str(e) else type(e).__name__)
		driver.setDefaultStream(None)
		sampler = ClockSampler(node.device) if node.index == 0 else None
		if sampler:
			sampler.start()
		ms, launches = timed(args.steps)
		clocks = sampler.stop() if sampler else None
		msE2e = eagerE2eMs
		api = "eager module API"

	for g in graphs:                       # graphs hold captured NCCL work: release them before the communicator goes away
		g.destroy()
	graphs.clear()

	# roofline pass: the same K steps with CUDA events around every launch of each kernel family
	driver.profileEnable(True)
	msProf, _ = timed(args.steps)
	driver.profileEnable(False)
	families = {name: driver.profileCollect(name) for name in driver.PROF_FAMILIES}

	# data parallelism keeps the replicas identical: after all those steps every rank must hold the same parameters, bit for bit
	# (same initial broadcast, same averaged gradients, same update kernel)
	checksumEqual = None
	if node.gridsize > 1:
		flat = [gv.data.get() for gv in optimizer.globalVar.values()]
		chk = float(sum(np.frombuffer(a.tobytes(), np.uint32).astype(np.uint64).sum() % (1 << 52) for a in flat))
		checksumEqual = node.rendezvous.maxValue(chk) == -node.rendezvous.maxValue(-chk)
		if not checksumEqual:
			raise SystemExit("data-parallel replicas diverged: parameter checksums differ across ranks")

	if node.index != 0:
		node.close()
		return

	peaks = loadPeaks()
	images = BATCH * node.gridsize * args.steps
	value = images / (ms * 1e-3)

	total = sum(f["ms"] for f in families.values()) or 1.0
	top = max(families, key=lambda name: families[name]["ms"])
	fam = families[top]
	if top == "gemm":     # tcgen05 launches above the machine ridge (3x3 / 7x7 convolutions, large GEMMs)
		achieved = fam["flops"] / (fam["ms"] * 1e-3) / 1e12
		if args.dtype == "f32" and "tf32" in peaks:
			peak = peaks["tf32"]
			note = "tf32 products: cuBLAS fp32 GEMM with TF32 math, 8192^3, sustained, measured on a B200 of this pool (profiles/r02_tf32_peak.json)"
		elif args.dtype == "f32":
			peak = peaks["bf16"] / 2.0
			note = "tf32 products: 0.5 x the %s bf16 sustained GEMM peak (no tf32 figure in MEASURED_PEAKS.json)" % peaks["src"]
		else:
			peak = peaks["bf16"]
			note = "%s bf16 sustained GEMM peak (16-bit products, f32 accumulation)" % peaks["src"]
		roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
					"peak_note": note}
	else:
		achieved = fam["bytes"] / (fam["ms"] * 1e-3) / 1e9
		roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s", "frac": achieved / peaks["hbm"],
					"traffic": None, "peak_note": "%s STREAM-style copy bandwidth" % peaks["src"]}
	# DRAM traffic per launch of that family from the committed ncu capture of this command (profiles/, tools/summarize_launches.py)
	try:
		with open(os.path.join(ROOT, "profiles", "r02_family_traffic.json")) as f:
			captured = json.load(f)["families"]
		key = top if top in captured else ("gemm" if top.startswith("gemm") else top)
		roofline["traffic"] = captured[key]["dram_bytes_per_launch"]
		roofline["traffic_note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the family's launches in " \
								   "profiles/r02_launches.md (ncu, cold cache); algorithmic bytes per launch: %.0f" % (
									   fam["bytes"] / max(1, fam["launches"]))
	except Exception:      # noqa: BLE001 -- no capture committed
		pass
	roofline.update({
		"kernel": {"gemm": "umma_gemm_kernel, launches with arithmetic intensity above the ridge (tcgen05 implicit-GEMM 3x3 / 7x7 conv, GEMM)",
				   "gemm_hbm": "umma_gemm_kernel, launches below the ridge (1x1 convolutions: HBM-bound even at full efficiency)",
				   "bn_fwd": "bn_fwd_cluster_kernel (bn_stats_kernel + bn_apply_kernel for planes beyond the stash)",
				   "bn_bwd": "bn_bwd_cluster_kernel (bn_bwd_stats_kernel + bn_bwd_apply_kernel for planes beyond the stash)",
				   "eltwise": "ew_kernel", "pool": "pool kernels", "other": "other"}[top],
		"launches_per_step": fam["launches"] / args.steps, "avg_launch_us": fam["ms"] * 1e3 / max(1, fam["launches"]),
		"share_of_profiled_kernel_time": fam["ms"] / total, "profiled_ms_per_step": msProf / args.steps,
		"families_ms_per_step": {name: f["ms"] / args.steps for name, f in families.items()},
		"families_achieved": {
			name: ({"TFLOP/s": f["flops"] / (f["ms"] * 1e-3) / 1e12, "GB/s": f["bytes"] / (f["ms"] * 1e-3) / 1e9} if name.startswith("gemm")
				   else {"GB/s": f["bytes"] / (f["ms"] * 1e-3) / 1e9})
			for name, f in families.items() if f["ms"] > 0
		},
	})

	line = {
		"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": node.gridsize, "steps": args.steps, "warmup": warmup,
		"ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
		"dtype": "f32 (tensor-core contractions: tf32 products, f32 accumulation -- what cuDNN/cuBLAS TENSOR_OP_MATH give the reference here)"
				 if args.dtype == "f32" else "%s storage, f32 accumulation" % args.dtype,
		"data": "synthetic",
		"config": {
			"workload": "%s %s fwd+bwd, synthetic %dx3x224x224 per GPU%s" % (
				{"resnet50": "ResNet-50", "vgg16": "VGG-16"}[args.model], {"f32": "fp32"}.get(args.dtype, args.dtype), BATCH,
				" (BASELINE.json configs[1])" if args.model == "resnet50" and args.dtype == "f32" and BATCH == 64 else ""),
			"batch_per_gpu": BATCH,
			"global_batch": BATCH * node.gridsize, "parallelism": "dp%d" % node.gridsize,
			"step": "zeroGradParams + forward + backward (incl. conv1 dgrad) + grad mean over ranks + momentum-SGD update",
			"l2": "no explicit flush: one step streams ~20 GB of activations through the 126 MB L2, every kernel's inputs exceed L2 between reuses",
			"model_flops_per_image": FLOP_PER_IMAGE, "achieved_model_tflops_per_gpu": value / node.gridsize * FLOP_PER_IMAGE / 1e12,
			"api": api,
		},
		"eager": {"value": images / (eagerMs * 1e-3), "ms_per_step": eagerMs / args.steps, "e2e_value": images / (eagerE2eMs * 1e-3),
				  "host_enqueue_ms_per_step": hostEnqueueMs, "note": "same step driven op by op through the Python module API"},
		"clocks": clocks,
		"e2e": {"value": images / (msE2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": int(data.nbytes) * node.gridsize,
				"d2h_bytes_per_step": BATCH * 1000 * dt.itemsize * node.gridsize,
				"pipeline": "double-buffered: the H2D of step i+1's batch (pinned memory, copy stream) overlaps step i; the D2H of every "
							"step's output is in stream order" if api.startswith("StepGraph") else "H2D, step, D2H in stream order"},
		"gpu_launches": launches,
		"roofline": roofline,
	}

	if node.gridsize == 1 and not args.no_cpu:
		ips, dt = cpuBaseline(args.cpu_images, 1, 1)
		line["cpu_baseline"] = {
			"value": ips, "unit": "images/s", "cores": cpuThreads(), "kind": "port",
			"sample": "%d images, 1 warm-up + 1 timed fwd+bwd step of the numpy float32 oracle port (oracle/refnet.py), %.1f s; the "
					  "reference's own numpy CPU backend cannot run conv/pool/batch-norm backward (SURVEY F5)" % (args.cpu_images, dt)
		}

	if node.gridsize > 1:
		line["config"]["gradient_sync"] = "NCCL mean of ~24 MB gradient buckets on a communication stream, overlapped with the backward pass" \
			if node.sync is not None else "one NCCL all-reduce after the backward pass"
		line["config"]["param_checksum_equal_across_ranks"] = checksumEqual

	if node.gridsize == 1 and not args.no_ref_gpu:
		line["reference_gpu"] = referenceGpu(args)
	if node.gridsize == 1 and not args.no_cpu and args.model == "resnet50":
		line["cpu_reference_forward"] = referenceCpuForward()

	if graphNote:
		line["config"]["graph"] = graphNote
	print(json.dumps(line), flush=True)
	node.close()


def main():
	parser = argparse.ArgumentParser()
	parser.add_argument("--gpus", type=int, default=1)
	parser.add_argument("--steps", type=int, default=20)
	parser.add_argument("--warmup", type=int, default=5)
	parser.add_argument("--impl", default="ours", choices=["ours", "reference"])
	parser.add_argument("--cpu-images", type=int, default=8, help="bounded CPU-baseline sample (images per step)")
	parser.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
	parser.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-CUDA-backend leg (reference_gpu)")
	parser.add_argument("--profile-run", action="store_true", help="eager steps only, no timing / JSON: the command ncu wraps")
	parser.add_argument("--no-graph", action="store_true", help="time the eager module API only (no CUDA-graph replay)")
	parser.add_argument("--model", default="resnet50", choices=["resnet50", "vgg16"], help="side measurements; the headline is resnet50")
	parser.add_argument("--dtype", default="f32", choices=["f32", "bf16", "f16"], help="storage type (side measurements)")
	parser.add_argument("--batch", type=int, default=0, help="per-GPU batch (default 64)")
	args = parser.parse_args()

	if os.environ.get("PZ_BENCH_WATCHDOG"):      # debugging aid: dump every thread's stack and exit if the run takes too long
		import faulthandler
		faulthandler.dump_traceback_later(int(os.environ["PZ_BENCH_WATCHDOG"]), exit=True)

	if args.impl == "reference":
		referenceArm(args)
	else:
		gpuArm(args)


if __name__ == "__main__":
	main()
