"""1x1 / stride-1 fprop and dgrad over 16-byte aligned planes: error of the engine's result against a float64 contraction of the
same inputs, for the PZ_TMA_FPROP level of this process (0 = producer gather, 1 / 2 = MN-major operand through the copy engine,
rounded to tf32 in place after it lands)."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
from puzzlelib_b200.backend import getBackend

CASES = [(128, 28, 512, 4), (512, 28, 128, 4), (256, 14, 1024, 8), (1024, 14, 256, 8), (64, 12, 96, 3), (32, 6, 40, 5), (96, 32, 32, 2)]


def main():
	bnd = getBackend(0, 2)
	rng = np.random.RandomState(2)
	print("PZ_TMA_FPROP=%s PZ_DEBUG_SKIP=%s" % (os.environ.get("PZ_TMA_FPROP", "(unset)"), os.environ.get("PZ_DEBUG_SKIP", "(unset)")))
	worst = 0.0
	for C, H, K, N in CASES:
		x = rng.randn(N, C, H, H).astype(np.float32)
		w = (rng.randn(K, C, 1, 1) / np.sqrt(C)).astype(np.float32)
		b = rng.randn(K).astype(np.float32)
		dy = rng.randn(N, K, H, H).astype(np.float32)
		gx, gw, gdy = bnd.GPUArray.toGpu(x), bnd.GPUArray.toGpu(w), bnd.GPUArray.toGpu(dy)
		y = bnd.dnn.convNd(gx, gw, bnd.GPUArray.toGpu(b), 1, 0, 1, 1).get().astype(np.float64)
		dx = bnd.dnn.convNdBackwardData(gdy, gw, None, gx, 1, 0, 1, None, 1, allocator=bnd.memoryPool).get().astype(np.float64)
		w64 = w[:, :, 0, 0].astype(np.float64)
		wy = np.einsum("kc,nchw->nkhw", w64, x.astype(np.float64)) + b.reshape(1, K, 1, 1)
		wdx = np.einsum("kc,nkhw->nchw", w64, dy.astype(np.float64))
		ey = np.abs(y - wy).max() / np.abs(wy).max()
		ex = np.abs(dx - wdx).max() / np.abs(wdx).max()
		worst = max(worst, ey, ex)
		print("C=%4d H=%2d K=%4d N=%d  fprop max rel err %.2e  dgrad %.2e" % (C, H, K, N, ey, ex))
	print("worst %.2e %s" % (worst, "OK" if worst < 1e-3 else "FAIL"))


if __name__ == "__main__":
	main()
