import sys, numpy as np
sys.path.insert(0, '/root/repo')
from puzzlelib_b200.backend import getBackend
bnd = getBackend(0, 2)
G = lambda a: bnd.GPUArray.toGpu(np.ascontiguousarray(a))
rng = np.random.RandomState(11)
names = ("out", "sm", "siv", "mean", "var", "dx", "ds", "db")
for shape in ((64, 16, 55, 55), (32, 64, 7, 7), (6, 10, 28, 28)):
	C = shape[1]
	x, dy = rng.randn(*shape).astype(np.float32), rng.randn(*shape).astype(np.float32)
	scale, bias = rng.randn(C).astype(np.float32), rng.randn(C).astype(np.float32)
	runs = []
	for _ in range(3):
		mean, var = G(np.zeros(C, np.float32)), G(np.ones(C, np.float32))
		out, sm, siv = bnd.dnn.batchNormNd(G(x), mean, var, G(scale), G(bias), 1e-5, 0.5, False)
		dx, ds, db = bnd.dnn.batchNormNdBackward(G(dy), G(x), G(scale), sm, siv, 1e-5)
		runs.append([a.get() for a in (out, sm, siv, mean, var, dx, ds, db)])
	for k in (1, 2):
		for name, a, b in zip(names, runs[0], runs[k]):
			if not np.array_equal(a, b):
				d = np.abs(a.astype(np.float64) - b)
				print(shape, "run", k, name, "differs: n=%d max=%.3e nan=%d" % ((a != b).sum(), np.nanmax(d), np.isnan(a).sum() + np.isnan(b).sum()), np.argwhere(a != b)[:5].tolist())
	data = G(x)
	mean, var = G(np.zeros(C, np.float32)), G(np.ones(C, np.float32))
	res, _, _ = bnd.dnn.batchNormNd(data, mean, var, G(scale), G(bias), 1e-5, 0.5, False, out=data)
	got = data.get()
	if not np.array_equal(got, runs[0][0]):
		d = np.abs(got.astype(np.float64) - runs[0][0])
		print(shape, "in-place differs: n=%d max=%.3e" % ((got != runs[0][0]).sum(), d.max()), np.argwhere(got != runs[0][0])[:5].tolist())
print("done")
