"""Per-op timing of the bandwidth-bound kernels on the ResNet-50 (N=64) shapes: BN fwd/bwd, ReLU fwd/bwd, Add, max-pool.
Prints ms and GB/s over the algorithmic bytes.  usage: python tools/bench_ops.py [N] [only-substring]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from puzzlelib_b200.backend import getBackend


def backend():
	return getBackend(0, 2)
from puzzlelib_b200 import driver

# (C, H) of the batch-norm inputs of ResNet-50 with their multiplicity
BN_SHAPES = [(64, 112, 1), (64, 55, 6), (256, 55, 4), (128, 28, 8), (512, 28, 5), (256, 14, 12), (1024, 14, 7), (512, 7, 6), (2048, 7, 4)]


def timeit(fn, reps=20):
	fn()
	e0, e1 = driver.Event(), driver.Event()
	e0.record()
	for _ in range(reps):
		fn()
	e1.record()
	e1.synchronize()
	return e0.timeTill(e1) / reps


def main():
	N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
	only = sys.argv[2] if len(sys.argv) > 2 else ""
	bnd = backend()
	f32 = np.float32
	total = {}

	def report(name, shape, count, ms, nbytes):
		total[name] = total.get(name, 0.0) + ms * count
		print("%-10s %-22s x%-2d %8.3f ms %8.0f GB/s" % (name, "x".join(map(str, shape)), count, ms, nbytes / ms / 1e6))

	for C, H, count in BN_SHAPES:
		shape = (N, C, H, H)
		tag = "%dx%d" % (C, H)
		if only and only not in "bn relu add " + tag:
			continue
		bnonly = only == "bn"
		x = bnd.GPUArray.toGpu(np.random.randn(*shape).astype(f32))
		dy = bnd.GPUArray.toGpu(np.random.randn(*shape).astype(f32))
		out = bnd.GPUArray.empty(shape, f32)
		mean, var = bnd.GPUArray.zeros((C, ), f32), bnd.GPUArray.zeros((C, ), f32)
		scale, bias = bnd.GPUArray.toGpu(np.ones(C, f32)), bnd.GPUArray.zeros((C, ), f32)
		E = float(x.size) * 4
		res = bnd.dnn.batchNormNd(x, mean, var, scale, bias, 1e-5, 1.0, False, 1, out=out)
		sm, siv = res[1], res[2]
		report("bn_fwd", shape, count, timeit(lambda: bnd.dnn.batchNormNd(x, mean, var, scale, bias, 1e-5, 1.0, False, 1, out=out)), 2 * E)
		report("bn_bwd", shape, count, timeit(lambda: bnd.dnn.batchNormNdBackward(dy, x, scale, sm, siv, 1e-5, 1, out=out)), 3 * E)
		if bnonly:
			continue
		relu, reluDer, add = bnd.reluKer(f32), bnd.reluDerKer(f32), bnd.addKer(f32)
		report("relu_fwd", shape, count, timeit(lambda: relu(out, x)), 2 * E)
		report("relu_bwd", shape, count, timeit(lambda: reluDer(out, dy, x)), 3 * E)
		report("add", shape, count, timeit(lambda: add(out, x, 1.0, dy, 1.0)), 3 * E)

	if not only or only in "pool":
		x = bnd.GPUArray.toGpu(np.maximum(np.random.randn(N, 64, 112, 112), 0).astype(f32))
		y = bnd.dnn.poolNd(x, 3, 2, 0, 0)
		dy = bnd.GPUArray.toGpu(np.random.randn(*y.shape).astype(f32))
		dx = bnd.GPUArray.empty(x.shape, f32)
		report("pool_fwd", x.shape, 1, timeit(lambda: bnd.dnn.poolNd(x, 3, 2, 0, 0, out=y)), 4.0 * (x.size + y.size))
		report("pool_bwd", x.shape, 1, timeit(lambda: bnd.dnn.poolNdBackward(dy, x, y, 3, 2, 0, 0, out=dx)), 4.0 * (2 * x.size + 2 * y.size))
	print("totals over the net (ms):", ", ".join("%s %.2f" % kv for kv in total.items()))


if __name__ == "__main__":
	main()
