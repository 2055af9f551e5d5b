"""Generate tests/golden/ref_cpu_train.npz from the REFERENCE ITSELF (build container only): the optimizer update kernels and the
dropout kernels of the reference's gcc-JIT CPU backend (CPU/Kernels/ElementWise.py:156-285) on seeded inputs, plus the
reference's own Adam / NesterovSGD / MomentumSGD optimizer objects stepping a variable (Optimizers/*.py).

usage: python tools/gen_golden_train.py
"""
import os

import numpy as np

from gen_golden import importReference, OUT


def main():
	importReference()
	from PuzzleLib.Backend import gpuarray
	from PuzzleLib.Backend.Kernels.ElementWise import adamKer, classicMomSGDKer, nesterovMomSGDKer, dropoutKer, dropout2dKer

	rng = np.random.RandomState(20261018)
	gold = {}
	shape = (11, 13)
	f32 = np.float32

	w, dw, mom = (rng.randn(*shape).astype(f32) for _ in range(3))
	ms, mg = (1.0 + rng.randn(*shape) ** 2).astype(f32), rng.randn(*shape).astype(f32)
	gold["w"], gold["dw"], gold["mom"], gold["mg"], gold["ms"] = w, dw, mom, mg, ms

	for name, ker in (("classic", classicMomSGDKer), ("nesterov", nesterovMomSGDKer)):
		gw, gmom = gpuarray.to_gpu(w.copy()), gpuarray.to_gpu(mom.copy())
		ker(np.dtype(f32))(gw, gpuarray.to_gpu(dw), gmom, f32(0.01), f32(0.9))
		gold["%s_w" % name], gold["%s_mom" % name] = gw.get(), gmom.get()

	gw, gmg, gms = gpuarray.to_gpu(w.copy()), gpuarray.to_gpu(mg.copy()), gpuarray.to_gpu(ms.copy())
	adamKer(np.dtype(f32))(gw, gpuarray.to_gpu(dw), gmg, gms, f32(0.0316), f32(0.1), f32(0.001), f32(1e-8))
	gold["adam_w"], gold["adam_mg"], gold["adam_ms"] = gw.get(), gmg.get(), gms.get()
	gold["adam_args"] = np.array([0.0316, 0.1, 0.001, 1e-8], dtype=f32)

	x = rng.randn(4, 6, 5, 7).astype(f32)
	words = rng.randint(0, np.iinfo(np.uint32).max, size=x.size, dtype=np.int64).astype(np.uint32)
	v, p = np.uint32(int(0.7 * np.iinfo(np.uint32).max)), f32(0.7)
	out = gpuarray.empty(x.shape, dtype=f32)
	dropoutKer(np.dtype(f32))(out, gpuarray.to_gpu(x), gpuarray.to_gpu(words), v, p)
	gold["drop_x"], gold["drop_words"], gold["drop_v"], gold["drop_p"], gold["drop_y"] = x, words, np.array([v]), np.array([p]), out.get()
	mapwords = words[:x.shape[0] * x.shape[1]].copy()
	dropout2dKer(np.dtype(f32))(out, gpuarray.to_gpu(x), gpuarray.to_gpu(mapwords), v, p, np.int32(x.shape[2] * x.shape[3]))
	gold["drop2d_words"], gold["drop2d_y"] = mapwords, out.get()

	# the optimizer OBJECTS: three steps on one variable with a fixed gradient sequence
	from PuzzleLib.Variable import Variable
	from PuzzleLib.Optimizers.Adam import Adam
	from PuzzleLib.Optimizers.NesterovSGD import NesterovSGD
	from PuzzleLib.Optimizers.MomentumSGD import MomentumSGD

	class Holder:
		def __init__(self, var):
			self.var = var

		def getVarTable(self):
			return {self.var: ["w"]}

		def getVar(self, name):
			return self.var

	grads = [rng.randn(*shape).astype(f32) for _ in range(3)]
	gold["opt_grads"] = np.stack(grads)
	for name, make in (("adam", lambda: Adam(alpha=1e-2)), ("nesterov", lambda: NesterovSGD(learnRate=1e-1, momRate=0.9)),
					   ("momentum", lambda: MomentumSGD(learnRate=1e-1, momRate=0.9))):
		var = Variable(gpuarray.to_gpu(w.copy()))
		opt = make()
		opt.setupOn(Holder(var))
		for g in grads:
			var.grad.set(g)
			opt.update()
		gold["opt_%s_w" % name] = var.data.get()

	path = os.path.join(OUT, "ref_cpu_train.npz")
	np.savez_compressed(path, **gold)
	print("wrote %s: %d arrays, %d bytes" % (path, len(gold), os.path.getsize(path)))


if __name__ == "__main__":
	main()
