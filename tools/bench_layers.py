"""Per-layer timing of the conv engine on the ResNet-50 (N=64) shapes: fprop / dgrad / wgrad with random data.
Prints ms, TFLOP/s (algorithmic) and GB/s (compulsory bytes) per pass.  Run on the GPU box."""
import sys
import numpy as np
sys.path.insert(0, ".")
from puzzlelib_b200.backend import getBackend


def backend():
	return getBackend(0, 2)
from puzzlelib_b200 import driver

# C, H, K, R, stride, pad, count   
LAYERS = [(3, 224, 64, 7, 2, 3, 1), (64, 55, 64, 1, 1, 0, 1), (64, 55, 64, 3, 1, 1, 3), (64, 55, 256, 1, 1, 0, 4), (256, 55, 64, 1, 1, 0, 2),
		  (256, 55, 128, 1, 2, 0, 1), (256, 55, 512, 1, 2, 0, 1), (128, 28, 128, 3, 1, 1, 4), (128, 28, 512, 1, 1, 0, 4),
		  (512, 28, 128, 1, 1, 0, 3), (512, 28, 256, 1, 2, 0, 1), (512, 28, 1024, 1, 2, 0, 1), (256, 14, 256, 3, 1, 1, 6),
		  (256, 14, 1024, 1, 1, 0, 6), (1024, 14, 256, 1, 1, 0, 5), (1024, 14, 512, 1, 2, 0, 1), (1024, 14, 2048, 1, 2, 0, 1),
		  (512, 7, 512, 3, 1, 1, 3), (512, 7, 2048, 1, 1, 0, 3), (2048, 7, 512, 1, 1, 0, 2)]


def timeit(fn, reps=5):
	fn()
	e0, e1 = driver.Event(), driver.Event()
	e0.record()
	for _ in range(reps):
		fn()
	e1.record()
	e1.synchronize()
	return e0.timeTill(e1) / reps


def timeline(launches):
	"""per-role phase breakdown of the tcgen05 engine (instrumented library only: PZB200_LIB=puzzlelib_b200/libpzb200_timeline.so)"""
	import ctypes
	buf = (ctypes.c_ulonglong * 32)()
	if driver.lib.pz_debug_timeline(ctypes.cast(buf, ctypes.c_void_p)) != 0:
		return ""
	v = [float(x) for x in buf]
	sms, ghz = 148.0, 1.965
	def us(x, warps):      # clock sum -> microseconds per warp per launch (upper bound on the grid: 148 CTAs)
		return x / (warps * sms * launches) / (ghz * 1e3)
	prod = ["setup", "issue", "wait_empty", "wait_data", "store", "fence+arrive", "advance"]
	mma = ["decode", "wait_accempty", "wait_full", "mma+commit"]
	epi = ["wait_accfull", "drain+store"]
	out = ["    producers (us/warp): " + "  ".join("%s %.1f" % (n, us(v[i], 16)) for i, n in enumerate(prod)),
		   "    mma thread (us):     " + "  ".join("%s %.1f" % (n, us(v[8 + i], 1)) for i, n in enumerate(mma)),
		   "    epilogue (us/warp):  " + "  ".join("%s %.1f" % (n, us(v[12 + i], 8)) for i, n in enumerate(epi))]
	return "\n".join(out)


def main():
	N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
	only = sys.argv[2] if len(sys.argv) > 2 else None
	bnd = backend()
	rng = np.random.RandomState(0)
	total = {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}
	print("%-34s %9s %9s %9s   (ms | TFLOP/s | GB/s)" % ("layer", "fprop", "dgrad", "wgrad"))
	for idx, (C, H, K, R, s, p, count) in enumerate(LAYERS):
		if only is not None and str(idx) != only:
			continue
		P = (H + 2 * p - R) // s + 1
		x = bnd.GPUArray.toGpu(rng.randn(N, C, H, H).astype(np.float32))
		w = bnd.GPUArray.toGpu(rng.randn(K, C, R, R).astype(np.float32))
		dy = bnd.GPUArray.toGpu(rng.randn(N, K, P, P).astype(np.float32))
		y = bnd.GPUArray.empty((N, K, P, P), np.float32)
		dx = bnd.GPUArray.empty((N, C, H, H), np.float32)
		dw = bnd.GPUArray.zeros((K, C, R, R), np.float32)
		flops = 2.0 * N * K * P * P * C * R * R
		nbytes = 4.0 * (x.size + w.size + y.size)
		passes = {
			"fprop": lambda: bnd.dnn.convNd(x, w, None, s, p, 1, 1, out=y),
			"dgrad": lambda: bnd.dnn.convNdBackwardData(dy, w, None, x, s, p, 1, None, 1, out=dx, allocator=bnd.memoryPool),
			"wgrad": lambda: bnd.dnn.convNdBackwardParams(x, dy, w, s, p, 1, 1, False, False, dw, None, 1.0, 1.0),
		}
		t, tl = {}, {}
		for name, fn in passes.items():
			timeline(1)
			t[name] = timeit(fn)
			tl[name] = timeline(6)
		cells = []
		for name in ("fprop", "dgrad", "wgrad"):
			total[name] += t[name] * count
			cells.append("%6.3f|%5.1f|%5.0f" % (t[name], flops / t[name] / 1e9, nbytes / t[name] / 1e6))
		print("%2d %4dx%-3d^2 -> %4d %dx%d s%d  x%d   %s" % (idx, C, H, K, R, R, s, count, "  ".join(cells)))
		for name in ("fprop", "dgrad", "wgrad"):
			if tl[name]:
				print("  %s\n%s" % (name, tl[name]))
	print("sum over the net (ms): fprop %.2f dgrad %.2f wgrad %.2f total %.2f" % (total["fprop"], total["dgrad"], total["wgrad"], sum(total.values())))


if __name__ == "__main__":
	main()
