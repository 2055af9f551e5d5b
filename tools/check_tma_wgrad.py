"""wgrad of 1x1 / 3x3 filters over 16-byte aligned planes: error of the engine's result against a float64 contraction of the same
inputs, for the PZ_TMA_WGRAD level of this process (0 = producer gather with round-to-nearest tf32, 1 / 2 = copy-engine operands,
whose landed fp32 tiles the producers round in place to the same values).  `slope-1` is the systematic shrink an un-rounded
(truncated) operand would show: -3.5e-4 per operand.  Run on the GPU box once per level: check_tma_wgrad.py [f32|f16] [case]."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
from puzzlelib_b200.backend import getBackend

CASES = [(128, 28, 512, 1, 0, 8), (512, 28, 128, 1, 0, 8), (256, 14, 1024, 1, 0, 16), (1024, 14, 256, 1, 0, 16), (128, 28, 128, 3, 1, 8),
		 (256, 14, 256, 3, 1, 8), (128, 12, 256, 1, 0, 5), (256, 8, 128, 1, 0, 3), (64, 24, 128, 3, 1, 4), (64, 16, 64, 5, 2, 3),
		 (192, 20, 128, 3, 1, 2)]


def main():
	bnd = getBackend(0, 2)
	rng = np.random.RandomState(1)
	print("PZ_TMA_WGRAD=%s" % os.environ.get("PZ_TMA_WGRAD", "(unset)"))
	worst = 0.0
	dt = {"f32": np.float32, "f16": np.float16}[sys.argv[1] if len(sys.argv) > 1 else "f32"]
	tol = 1e-3 if dt == np.float32 else 2e-3
	only = int(sys.argv[2]) if len(sys.argv) > 2 else None
	for idx, (C, H, K, R, pad, N) in enumerate(CASES):
		if only is not None and idx != only:
			continue
		x = rng.randn(N, C, H, H).astype(dt)
		dy = (rng.randn(N, K, H, H) / 8).astype(dt)
		w = bnd.GPUArray.toGpu(np.zeros((K, C, R, R), dt))
		dw = bnd.GPUArray.zeros((K, C, R, R), dt)
		bnd.dnn.convNdBackwardParams(bnd.GPUArray.toGpu(x), bnd.GPUArray.toGpu(dy), w, 1, pad, 1, 1, False, False, dw, None, 1.0, 0.0)
		got = dw.get().astype(np.float64)
		xp = np.pad(x.astype(np.float64), ((0, 0), (0, 0), (pad, pad), (pad, pad)))
		want = np.empty((K, C, R, R))
		for r in range(R):
			for s in range(R):
				want[:, :, r, s] = np.einsum("nkpq,ncpq->kc", dy.astype(np.float64), xp[:, :, r:r + H, s:s + H])
		err = np.abs(got - want).max() / np.abs(want).max()
		# systematic shrink (truncation bias): regression slope of got on want
		slope = float((got * want).sum() / (want * want).sum())
		worst = max(worst, err)
		print("C=%4d H=%2d K=%4d R=%d N=%2d  max rel err %.2e  slope-1 %+.2e" % (C, H, K, R, N, err, slope - 1.0))
	print("worst %.2e %s" % (worst, "OK" if worst < tol else "FAIL"))


if __name__ == "__main__":
	main()
