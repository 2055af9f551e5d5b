"""Seeded operator cases driven through the REFERENCE's own function table (`PuzzleLib.Backend.{gpuarray,Dnn,Blas,Memory}`,
`PuzzleLib.Backend.Kernels.*`) -- the calls `Modules/*` make.

The same table runs twice:
* `tools/gen_golden_cuda.py` runs it on a B200 with the reference's OWN cuDNN 9 / cuBLAS 12 / NVRTC backend
  (baseline/_ref, built by baseline/build_ref.py) and stores every output in tests/golden/ref_cuda_ops.npz;
* `tests/test_gpu_parity_cuda.py` runs it with this repository's backend behind the seam and compares
  (bit-exact for masks / argmax / integer work, the tolerances of BASELINE.json's north_star for floating point);
  `tests/test_oracle.py` holds the numpy oracle to the same file on the CPU.

Every case is a function `case(B, rng) -> dict[name -> numpy array]`; `B` is the namespace built by `bind()`.
Inputs are stored too (prefix `in_`), so the CPU oracle can be checked without re-deriving them.
"""
import types

import numpy as np


def bind():
	"""The reference's late-bound function table (whatever backend PuzzleLib was configured with)."""
	from PuzzleLib.Backend import gpuarray, Dnn, Blas, Memory
	from PuzzleLib.Backend.Kernels import ElementWise, MatVec, Pool, Costs, PRelu, Pad, Embedder, Upsample
	return types.SimpleNamespace(gpuarray=gpuarray, Dnn=Dnn, Blas=Blas, Memory=Memory, ElementWise=ElementWise, MatVec=MatVec,
								 Pool=Pool, Costs=Costs, PRelu=PRelu, Pad=Pad, Embedder=Embedder, Upsample=Upsample)


def _g(B, ary):
	return B.gpuarray.to_gpu(np.ascontiguousarray(ary))


def _seq(v, nd):
	return (v, ) * nd if isinstance(v, int) else tuple(v)


# ---------------------------------------------------------------------------------------------------------- convolution
CONV_GEOMETRIES = {
	# name: (N, C, spatial, K, filter, stride, pad, dilation, groups, bias)
	"conv3x3": (2, 3, (9, 9), 4, (3, 3), 1, 1, 1, 1, True),
	"conv3x3s2": (2, 4, (11, 11), 6, (3, 3), 2, 1, 1, 1, True),
	"convdil2": (1, 2, (12, 12), 3, (3, 3), 1, 2, 2, 1, False),
	"convgroups": (2, 4, (8, 8), 6, (3, 3), 1, 1, 1, 2, True),
	"conv1x1s2": (2, 8, (14, 14), 16, (1, 1), 2, 0, 1, 1, False),
	"conv7x7s2": (1, 3, (32, 32), 8, (7, 7), 2, 3, 1, 1, False),
	"convrect": (2, 3, (7, 10), 5, (2, 3), (1, 2), (0, 1), 1, 1, True),
	"conv1x1wide": (3, 64, (7, 7), 32, (1, 1), 1, 0, 1, 1, False),
	"conv3d": (1, 2, (5, 6, 6), 3, (2, 3, 3), (1, 2, 1), (1, 0, 1), 1, 1, True),
}


def convCase(name, dtype):
	N, C, spatial, K, fsize, stride, pad, dilation, groups, withbias = CONV_GEOMETRIES[name]

	def case(B, rng):
		Dnn = B.Dnn
		x = rng.randn(N, C, *spatial).astype(dtype)
		W = (rng.randn(K, C // groups, *fsize) / np.sqrt(C // groups * np.prod(fsize))).astype(dtype)
		bshape = (1, K) + (1, ) * len(spatial)
		b = rng.randn(*bshape).astype(dtype) if withbias else None

		data, gW = _g(B, x), _g(B, W)
		gb = _g(B, b) if withbias else None

		y = Dnn.convNd(data, gW, gb, stride, pad, dilation, groups, Dnn.ConvFwdAlgo.implicitGemm)
		dy = rng.randn(*y.shape).astype(dtype)
		grad = _g(B, dy)

		dx = Dnn.convNdBackwardData(grad, gW, data, stride, pad, dilation, groups, Dnn.ConvBwdDataAlgo.algo0)

		# overwrite (momentum 0) into NaN-poisoned buffers, then accumulate (scale 0.5, momentum 0.9) into known ones
		wgrad = _g(B, np.full(W.shape, np.nan, dtype))
		bgrad = _g(B, np.full(bshape, np.nan, dtype)) if withbias else None
		Dnn.convNdBackwardParams(data, grad, gW, gb, stride, pad, dilation, groups, wgrad, bgrad, 1.0, 0.0,
								 Dnn.ConvBwdFilterAlgo.algo0)

		w0 = rng.randn(*W.shape).astype(dtype)
		b0 = rng.randn(*bshape).astype(dtype)
		wacc = _g(B, w0)
		bacc = _g(B, b0) if withbias else None
		Dnn.convNdBackwardParams(data, grad, gW, gb, stride, pad, dilation, groups, wacc, bacc, 0.5, 0.9,
								 Dnn.ConvBwdFilterAlgo.algo0)

		out = {"in_x": x, "in_W": W, "in_dy": dy, "in_w0": w0, "y": y.get(), "dx": dx.get(), "wgrad": wgrad.get(), "wacc": wacc.get()}
		if withbias:
			out.update({"in_b": b, "in_b0": b0, "bgrad": bgrad.get(), "bacc": bacc.get()})
		return out

	return case


DECONV_GEOMETRIES = {
	# name: (N, inmaps, spatial, outmaps, filter, stride, pad, dilation, postpad, groups, bias)
	"deconv3x3s2": (2, 4, (5, 5), 3, (3, 3), 2, 1, 1, 1, 1, True),
	"deconv2x2s2": (2, 6, (4, 6), 4, (2, 2), 2, 0, 1, 0, 1, False),
	"deconv4x4s2p1": (1, 8, (7, 7), 5, (4, 4), 2, 1, 1, 0, 1, True),
}


def deconvCase(name, dtype):
	N, C, spatial, K, fsize, stride, pad, dilation, postpad, groups, withbias = DECONV_GEOMETRIES[name]

	def case(B, rng):
		Dnn = B.Dnn
		x = rng.randn(N, C, *spatial).astype(dtype)
		W = (rng.randn(C, K // groups, *fsize) / np.sqrt(C * np.prod(fsize))).astype(dtype)
		bshape = (1, K) + (1, ) * len(spatial)
		b = rng.randn(*bshape).astype(dtype) if withbias else None

		data, gW = _g(B, x), _g(B, W)
		gb = _g(B, b) if withbias else None

		y = Dnn.deconvNd(data, gW, gb, stride, pad, dilation, postpad, groups, Dnn.ConvBwdDataAlgo.algo0)
		dy = rng.randn(*y.shape).astype(dtype)
		grad = _g(B, dy)

		dx = Dnn.deconvNdBackwardData(grad, gW, data, stride, pad, dilation, groups, Dnn.ConvFwdAlgo.implicitGemm)

		wgrad = _g(B, np.zeros(W.shape, dtype))
		bgrad = _g(B, np.zeros(bshape, dtype)) if withbias else None
		Dnn.deconvNdBackwardParams(data, grad, gW, gb, stride, pad, dilation, groups, wgrad, bgrad, 1.0, 0.0,
								   Dnn.ConvBwdFilterAlgo.algo0)

		out = {"in_x": x, "in_W": W, "in_dy": dy, "y": y.get(), "dx": dx.get(), "wgrad": wgrad.get()}
		if withbias:
			out.update({"in_b": b, "bgrad": bgrad.get()})
		return out

	return case


# ---------------------------------------------------------------------------------------------------------- batch norm
def batchNormCase(shape, dtype, mode="spatial", factor=0.3):
	def case(B, rng):
		Dnn = B.Dnn
		bnmode = getattr(Dnn.BatchNormMode, mode)
		C = shape[1] if mode == "spatial" else int(np.prod(shape[1:]))
		pshape = (1, shape[1]) + (1, ) * (len(shape) - 2) if mode == "spatial" else (1, ) + tuple(shape[1:])

		# off-centre inputs with a different spread per channel: catches E[x^2] - E[x]^2 style statistics
		x = (rng.randn(*shape) * (0.5 + rng.rand(*pshape)) + 3.0 * rng.randn(*pshape)).astype(dtype)
		scale = (1.0 + 0.3 * rng.randn(*pshape)).astype(np.float32)
		bias = rng.randn(*pshape).astype(np.float32)
		mean0 = rng.randn(*pshape).astype(np.float32)
		var0 = (0.5 + rng.rand(*pshape)).astype(np.float32)
		dy = rng.randn(*shape).astype(dtype)

		data, grad = _g(B, x), _g(B, dy)
		gscale, gbias, mean, var = _g(B, scale), _g(B, bias), _g(B, mean0), _g(B, var0)

		y, savemean, saveinvvar = Dnn.batchNormNd(data, gscale, gbias, mean, var, 1e-5, factor, False, bnmode)
		dx, dscale, dbias = Dnn.batchNormNdBackward(data, grad, gscale, savemean, saveinvvar, 1e-5, bnmode)

		# inference with the updated running statistics
		yinfer = Dnn.batchNormNd(data, gscale, gbias, mean, var, 1e-5, 1.0, True, bnmode)

		assert C == scale.size
		return {
			"in_x": x, "in_scale": scale, "in_bias": bias, "in_mean0": mean0, "in_var0": var0, "in_dy": dy,
			"y": y.get(), "savemean": savemean.get(), "saveinvvar": saveinvvar.get(), "runmean": mean.get(), "runvar": var.get(),
			"dx": dx.get(), "dscale": dscale.get(), "dbias": dbias.get(), "yinfer": yinfer.get()
		}

	return case


def instanceNormCase(shape, dtype):
	def case(B, rng):
		Dnn = B.Dnn
		x = (rng.randn(*shape) + 2.0 * rng.randn(shape[0], shape[1], 1, 1)).astype(dtype)
		scale = (1.0 + 0.3 * rng.randn(1, shape[1], 1, 1)).astype(np.float32)
		bias = rng.randn(1, shape[1], 1, 1).astype(np.float32)
		dy = rng.randn(*shape).astype(dtype)

		data, grad, gscale, gbias = _g(B, x), _g(B, dy), _g(B, scale), _g(B, bias)
		y, savemean, saveinvvar, extscale = Dnn.instanceNorm2d(data, gscale, gbias, 1e-5)
		dx, dscale, dbias = Dnn.instanceNorm2dBackward(grad, data, extscale, savemean, saveinvvar, 1e-5, True)
		return {"in_x": x, "in_scale": scale, "in_bias": bias, "in_dy": dy, "y": y.get(), "savemean": savemean.get(),
				"saveinvvar": saveinvvar.get(), "dx": dx.get(), "dscale": dscale.get(), "dbias": dbias.get()}

	return case


# ---------------------------------------------------------------------------------------------------------- pooling
def poolCase(shape, size, stride, pad, mode, dtype, ties):
	def case(B, rng):
		Dnn = B.Dnn
		if ties:
			# few distinct values: most windows hold several equal maxima (ReLU outputs look like this)
			x = rng.randint(0, 3, size=shape).astype(dtype)
		else:
			x = rng.randn(*shape).astype(dtype)

		data = _g(B, x)
		y, workspace = Dnn.poolNd(data, size, stride, pad, getattr(Dnn.PoolMode, mode), False)
		dy = rng.randn(*y.shape).astype(dtype)
		grad = _g(B, dy)
		dx = Dnn.poolNdBackward(data, y, grad, workspace, size, stride, pad, getattr(Dnn.PoolMode, mode))
		return {"in_x": x, "in_dy": dy, "y": y.get(), "dx": dx.get()}

	return case


def maxpoolMaskCase(shape, size, stride, pad, ties):
	def case(B, rng):
		x = (rng.randint(0, 3, size=shape) if ties else rng.randn(*shape)).astype(np.float32)
		data = _g(B, x)
		y, mask = B.Pool.maxpool2d(data, _seq(size, 2), _seq(stride, 2), _seq(pad, 2))
		dy = rng.randn(*y.shape).astype(np.float32)
		grad = _g(B, dy)
		dx = B.Pool.maxpool2dBackward(grad, data.shape, mask, _seq(size, 2), _seq(stride, 2), _seq(pad, 2))
		up = B.Pool.maxunpool2d(y, data.shape, mask)
		dup = B.Pool.maxunpool2dBackward(_g(B, x), y.shape, mask)
		return {"in_x": x, "in_dy": dy, "y": y.get(), "mask": mask.get(), "dx": dx.get(), "unpool": up.get(), "unpoolgrad": dup.get()}

	return case


# ---------------------------------------------------------------------------------------------------------- softmax, lrn
def softmaxCase(shape, dtype):
	def case(B, rng):
		x = (3.0 * rng.randn(*shape)).astype(dtype)
		dy = rng.randn(*shape).astype(dtype)
		y = B.Dnn.softmaxNd(_g(B, x))
		dx = B.Dnn.softmaxNdBackward(y, _g(B, dy))
		return {"in_x": x, "in_dy": dy, "y": y.get(), "dx": dx.get()}

	return case


def lrnCase(shape, kind, N, dtype):
	def case(B, rng):
		Dnn = B.Dnn
		x = rng.randn(*shape).astype(dtype)
		dy = rng.randn(*shape).astype(dtype)
		data, grad = _g(B, x), _g(B, dy)
		alpha, beta, K = 1e-2, 0.75, 2.0
		if kind == "cross":
			y, ws = Dnn.crossMapLRN(data, N, alpha, beta, K, False)
			dx = Dnn.crossMapLRNBackward(data, y, grad, ws, N, alpha, beta, K)
		else:
			y, ws = Dnn.mapLRN(data, None, N, alpha, beta, K, False)
			dx = Dnn.mapLRNBackward(data, y, grad, None, ws, N, alpha, beta, K)
		return {"in_x": x, "in_dy": dy, "y": y.get(), "dx": dx.get()}

	return case


# ---------------------------------------------------------------------------------------------------------- blas
def gemmCase(M, N, K, dtype):
	def case(B, rng):
		Blas = B.Blas
		A = rng.randn(M, K).astype(dtype)
		Bm = rng.randn(K, N).astype(dtype)
		C0 = rng.randn(M, N).astype(dtype)
		gA, gB = _g(B, A), _g(B, Bm)
		out = {"in_A": A, "in_B": Bm, "in_C0": C0}
		out["nn"] = Blas.mulMatrixOnMatrix(gA, gB).get()
		out["tn"] = Blas.mulMatrixOnMatrix(_g(B, A.T), gB, transpA=True).get()
		out["nt"] = Blas.mulMatrixOnMatrix(gA, _g(B, Bm.T), transpB=True).get()
		acc = _g(B, C0)
		Blas.mulMatrixOnMatrix(gA, gB, out=acc, alpha=0.5, beta=0.9)
		out["acc"] = acc.get()
		out["colsum"] = Blas.sumOnMatrix(gA).get()
		out["rowsum"] = Blas.sumOnMatrix(gA, cols=False).get()
		return out

	return case


def matvecCase(dtype):
	def case(B, rng):
		A = rng.randn(37, 120).astype(dtype)
		u, v, w = rng.randn(120).astype(dtype), rng.randn(37).astype(dtype), rng.randn(30).astype(dtype)
		gA = _g(B, A)
		ties = rng.randint(0, 4, size=(19, 50)).astype(dtype)
		return {
			"in_A": A, "in_u": u, "in_v": v, "in_w": w, "in_ties": ties,
			"addrow": B.MatVec.addVecToMat(_g(B, u), gA, 1, None).get(),
			"addcol": B.MatVec.addVecToMat(_g(B, v), gA, 0, None).get(),
			"addtile": B.MatVec.addVecToMat(_g(B, w), gA, 1, None).get(),
			"argmax1": B.MatVec.argmax(gA, 1).get(), "argmax0": B.MatVec.argmax(gA, 0).get(),
			"argmaxties": B.MatVec.argmax(_g(B, ties), 1).get(),
		}

	return case


# ---------------------------------------------------------------------------------------------------------- elementwise
ACTIVATIONS = {
	"sigmoid": (), "tanh": (), "relu": (), "leakyRelu": (0.01, ), "elu": (1.0, ), "softPlus": (), "clip": (0.0, 6.0), "gelu": ()
}


def activationCase(dtype):
	def case(B, rng):
		E = B.ElementWise
		x = (2.5 * rng.randn(1031)).astype(dtype)
		x[:8] = [0.0, -0.0, 6.0, -6.0, 1e-3, -1e-3, 20.0, -20.0]
		dy = rng.randn(1031).astype(dtype)
		out = {"in_x": x, "in_dy": dy}
		for name, args in ACTIVATIONS.items():
			data, grad = _g(B, x), _g(B, dy)
			y, dx = B.gpuarray.empty(x.shape, dtype), B.gpuarray.empty(x.shape, dtype)
			getattr(E, name + "Ker")(np.dtype(dtype))(y, data, *args)
			getattr(E, name + "DerKer")(np.dtype(dtype))(dx, grad, data if name == "gelu" else y, *args)
			out[name] = y.get()
			out[name + "Der"] = dx.get()
		return out

	return case


def blas1Case(dtype):
	def case(B, rng):
		x, y = rng.randn(777).astype(dtype), rng.randn(777).astype(dtype)
		gx, gy = _g(B, x), _g(B, y)
		out = {"in_x": x, "in_y": y}
		acc = _g(B, y)
		B.Blas.toVectorAddVector(acc, gx, alpha=-0.75)
		out["axpy"] = acc.get()
		out["axpby"] = B.Blas.addVectorToVector(gx, gy, alpha=0.3, beta=-1.7).get()
		lin = B.gpuarray.empty(x.shape, dtype)
		B.ElementWise.linearKer(np.dtype(dtype))(lin, gx, 1.5, -0.25)
		out["linear"] = lin.get()
		mul = B.gpuarray.empty(x.shape, dtype)
		B.ElementWise.mulKer(np.dtype(dtype))(mul, gx, gy)
		out["mul"] = mul.get()
		return out

	return case


def optimizerCase(dtype):
	def case(B, rng):
		E = B.ElementWise
		n = 513
		p0, g, m0 = rng.randn(n).astype(dtype), rng.randn(n).astype(dtype), (0.1 * rng.randn(n)).astype(dtype)
		out = {"in_p": p0, "in_g": g, "in_m": m0}

		p, m = _g(B, p0), _g(B, m0)
		E.classicMomSGDKer(np.dtype(dtype))(p, _g(B, g), m, 0.01, 0.9)
		out["momsgd_p"], out["momsgd_m"] = p.get(), m.get()

		p, m = _g(B, p0), _g(B, m0)
		E.nesterovMomSGDKer(np.dtype(dtype))(p, _g(B, g), m, 0.01, 0.9)
		out["nesterov_p"], out["nesterov_m"] = p.get(), m.get()

		mg0, ms0 = (0.1 * rng.randn(n)).astype(np.float32), (0.1 * rng.rand(n)).astype(np.float32)
		out["in_mg"], out["in_ms"] = mg0, ms0
		p, mg, ms = _g(B, p0), _g(B, mg0), _g(B, ms0)
		E.adamKer(np.dtype(dtype))(p, _g(B, g), mg, ms, 1e-3, 0.1, 0.001, 1e-8)
		out["adam_p"], out["adam_mg"], out["adam_ms"] = p.get(), mg.get(), ms.get()
		return out

	return case


def crossEntropyCase(shape):
	def case(B, rng):
		scores = (2.0 * rng.randn(*shape)).astype(np.float32)
		lshape = (shape[0], ) + tuple(shape[2:])
		labels = rng.randint(0, shape[1], size=lshape).astype(np.int32)
		error, grad = B.Costs.crossEntropyKernel(_g(B, scores), _g(B, labels), None, None)
		weights = (0.5 + rng.rand(shape[1])).astype(np.float32)
		werror, wgrad = B.Costs.crossEntropyKernel(_g(B, scores), _g(B, labels), _g(B, weights), None)
		return {"in_scores": scores, "in_labels": labels, "in_weights": weights, "error": error.get(), "grad": grad.get(),
				"werror": werror.get(), "wgrad": wgrad.get()}

	return case


def memoryCase(dtype):
	def case(B, rng):
		x = rng.randn(3, 4, 5, 6).astype(dtype)
		gx = _g(B, x)
		a, b = rng.randn(2, 3, 6, 6).astype(dtype), rng.randn(2, 2, 4, 4).astype(dtype)
		cat = B.Memory.depthConcat([_g(B, a), _g(B, b)])
		g = rng.randn(*cat.shape).astype(dtype)
		ga, gb = B.Memory.depthSplit(_g(B, g), [_g(B, a), _g(B, b)])
		return {
			"in_x": x, "in_a": a, "in_b": b, "in_g": g,
			"transpose": B.Memory.transpose(gx, (2, 0, 3, 1)).get(), "moveaxis": B.Memory.moveaxis(gx, 1, 3).get(),
			"swapaxes": B.Memory.swapaxes(gx, 0, 2).get(), "depthconcat": cat.get(), "splita": ga.get(), "splitb": gb.get(),
		}

	return case


# ---------------------------------------------------------------------------------------------------------- the table
f32, f16 = np.float32, np.float16

CASES = {}
for _name in CONV_GEOMETRIES:
	CASES["%s_f32" % _name] = convCase(_name, f32)
for _name in ("conv3x3", "conv3x3s2", "conv1x1s2", "convgroups"):
	CASES["%s_f16" % _name] = convCase(_name, f16)
for _name in DECONV_GEOMETRIES:
	CASES["%s_f32" % _name] = deconvCase(_name, f32)
CASES["deconv3x3s2_f16"] = deconvCase("deconv3x3s2", f16)

CASES["bn2d_f32"] = batchNormCase((4, 6, 5, 7), f32)
CASES["bn2d_big_f32"] = batchNormCase((8, 32, 14, 14), f32, factor=0.1)
CASES["bn2d_f16"] = batchNormCase((4, 6, 5, 7), f16)
CASES["bn3d_f32"] = batchNormCase((3, 4, 3, 4, 5), f32)
CASES["bn1d_f32"] = batchNormCase((16, 10, 1, 1), f32, factor=1.0)
CASES["bnperact_f32"] = batchNormCase((12, 30, 1, 1), f32, mode="perActivation")
CASES["instnorm_f32"] = instanceNormCase((3, 4, 6, 5), f32)

CASES["maxpool3s2_ties_f32"] = poolCase((2, 3, 11, 11), 3, 2, 0, "max", f32, True)
CASES["maxpool2s2_f32"] = poolCase((2, 4, 8, 10), 2, 2, 0, "max", f32, False)
CASES["maxpool3s2p1_ties_f32"] = poolCase((2, 3, 9, 9), 3, 2, 1, "max", f32, True)
CASES["maxpool3s1p1_ties_f32"] = poolCase((1, 2, 6, 7), 3, 1, 1, "max", f32, True)
CASES["avgpadpool_f32"] = poolCase((2, 3, 9, 9), 3, 2, 1, "avgWithPad", f32, False)
CASES["avgnopadpool_f32"] = poolCase((2, 3, 9, 9), 3, 2, 1, "avgNoPad", f32, False)
CASES["avgpool7_f32"] = poolCase((2, 5, 7, 7), 7, 1, 0, "avgWithPad", f32, False)
CASES["maxpool2s2_f16"] = poolCase((2, 4, 8, 10), 2, 2, 0, "max", f16, False)
CASES["maxpool3d_ties_f32"] = poolCase((1, 2, 6, 6, 6), 2, 2, 0, "max", f32, True)
CASES["maxpoolmask_f32"] = maxpoolMaskCase((2, 3, 9, 10), 3, 2, 1, False)
CASES["maxpoolmask_ties_f32"] = maxpoolMaskCase((2, 3, 11, 11), 3, 2, 0, True)

CASES["softmax_flat_f32"] = softmaxCase((6, 10, 1, 1), f32)
CASES["softmax_spatial_f32"] = softmaxCase((2, 5, 3, 4), f32)
CASES["softmax_flat_f16"] = softmaxCase((6, 10, 1, 1), f16)
CASES["lrn_cross_f32"] = lrnCase((2, 7, 5, 5), "cross", 5, f32)
CASES["lrn_map_f32"] = lrnCase((2, 3, 8, 8), "map", 3, f32)

CASES["gemm_f32"] = gemmCase(33, 50, 70, f32)
CASES["gemm_f16"] = gemmCase(32, 48, 64, f16)
CASES["matvec_f32"] = matvecCase(f32)
CASES["act_f32"] = activationCase(f32)
CASES["act_f16"] = activationCase(f16)
CASES["blas1_f32"] = blas1Case(f32)
CASES["blas1_f16"] = blas1Case(f16)
CASES["optim_f32"] = optimizerCase(f32)
CASES["optim_f16"] = optimizerCase(f16)
CASES["xent_flat"] = crossEntropyCase((16, 10))
CASES["xent_spatial"] = crossEntropyCase((4, 6, 3, 3))
CASES["memory_f32"] = memoryCase(f32)


def seedOf(name):
	return 1234 + sum(ord(c) * (i + 1) for i, c in enumerate(name)) % 100003


def run(B, names=None):
	"""-> {"case/key": array} for every case in `names` (default: all of CASES), each with its own fixed seed"""
	out = {}
	for name in (CASES if names is None else names):
		rng = np.random.RandomState(seedOf(name))
		fn = CASES[name] if name in CASES else SIDE_CASES[name]
		for key, val in fn(B, rng).items():
			out["%s/%s" % (name, key)] = np.asarray(val)
	return out


# ---------------------------------------------------------------------------------------------------------- side modules
# PReLU, reflection padding, embedding lookup, up-sampling, divisive normalisation: the kernel modules next to the hot path
# (Backend/Kernels/{PRelu,Pad,Embedder,Upsample}.py, Dnn.mapLRN with a means tensor).  Stored in tests/golden/ref_cuda_side.npz.
def preluCase(shared):
	def case(B, rng):
		x = rng.randn(3, 5, 4, 6).astype(np.float32)
		slopes = rng.randn(1 if shared else 5).astype(np.float32)
		dy = rng.randn(*x.shape).astype(np.float32)
		data, gs, grad = _g(B, x), _g(B, slopes), _g(B, dy)
		y = B.PRelu.prelu(data, gs, False, shared)
		dx = B.PRelu.preluBackwardData(grad, gs, data, shared)
		ds = B.PRelu.preluBackwardParams(data, grad, shared)
		return {"in_x": x, "in_slopes": slopes, "in_dy": dy, "y": y.get(), "dx": dx.get(), "dslopes": ds.get()}
	return case


def padCase(shape, pad, dtype):
	def case(B, rng):
		x = rng.randn(*shape).astype(dtype)
		fwd, bwd = (B.Pad.reflectpad1d, B.Pad.reflectpad1dBackward) if len(shape) == 3 else (B.Pad.reflectpad2d, B.Pad.reflectpad2dBackward)
		y = fwd(_g(B, x), pad)
		dy = rng.randn(*y.shape).astype(dtype)
		dx = bwd(_g(B, dy), pad)
		return {"in_x": x, "in_dy": dy, "y": y.get(), "dx": dx.get()}
	return case


def embedCase(dtype):
	def case(B, rng):
		vocab, emb = 50, 12
		idx = rng.randint(-1, vocab, size=(4, 9)).astype(np.int32)
		idx[0, :3] = 7                                                # repeated words: their gradient rows add up
		W = rng.randn(vocab, emb).astype(dtype)
		dy = rng.randn(4, 9, emb).astype(dtype)
		gW = _g(B, W)
		y = B.Embedder.embed(_g(B, idx), gW)
		B.Embedder.embedBackwardParams(_g(B, idx), _g(B, dy), gW, 0.25)
		return {"in_idx": idx, "in_W": W, "in_dy": dy, "y": y.get(), "W_after": gW.get()}
	return case


def upsampleCase(shape, scale, mode):
	def case(B, rng):
		x = rng.randn(*shape).astype(np.float32)
		fwd, bwd = (B.Upsample.upsample2d, B.Upsample.upsample2dBackward) if len(shape) == 4 else \
				   (B.Upsample.upsample3d, B.Upsample.upsample3dBackward)
		y = fwd(_g(B, x), scale, mode)
		dy = rng.randn(*y.shape).astype(np.float32)
		dx = bwd(_g(B, dy), scale, mode)
		return {"in_x": x, "in_dy": dy, "y": y.get(), "dx": dx.get()}
	return case


def lcnCase(shape, N):
	def case(B, rng):
		Dnn = B.Dnn
		x = rng.randn(*shape).astype(np.float32)
		dy = rng.randn(*shape).astype(np.float32)
		data, grad = _g(B, x), _g(B, dy)
		alpha, beta, K = 1e-2, 0.75, 2.0
		means, _ = Dnn.poolNd(data, (N, N), 1, (N // 2, N // 2), Dnn.PoolMode.avgWithPad, False)
		y, ws = Dnn.mapLRN(data, means, N, alpha, beta, K, False)
		dx, dmeans = Dnn.mapLRNBackward(data, y, grad, means, ws, N, alpha, beta, K)
		return {"in_x": x, "in_dy": dy, "means": means.get(), "y": y.get(), "dx": dx.get(), "dmeans": dmeans.get()}
	return case


SIDE_CASES = {
	"side_prelu": preluCase(False), "side_prelu_shared": preluCase(True),
	"side_pad1d_f32": padCase((2, 3, 11), (2, 3), np.float32), "side_pad2d_f32": padCase((2, 3, 7, 9), (2, 1, 3, 0), np.float32),
	"side_pad2d_f16": padCase((2, 3, 7, 9), (1, 2, 2, 2), np.float16),
	"side_embed_f32": embedCase(np.float32),
	"side_up2d_nearest": upsampleCase((2, 3, 5, 7), (2, 3), "nearest"), "side_up2d_linear": upsampleCase((2, 3, 5, 7), (2, 3), "linear"),
	"side_up3d_nearest": upsampleCase((1, 2, 3, 4, 5), (2, 1, 2), "nearest"),
	"side_up3d_linear": upsampleCase((1, 2, 3, 4, 4), (2, 2, 3), "linear"),
	"side_up3d_linear_hw": upsampleCase((1, 2, 3, 5, 4), (2, 2, 2), "linear"),     # inh != inw: the reference's addressing quirk shows
	"side_lcn": lcnCase((2, 3, 8, 9), 5),
}


# ---------------------------------------------------------------------------------------------------------- whole nets
def netCase(B, which, batch):
	"""Forward + backward of a whole reference model with seeded parameters; returns a SMALL summary (the output, a few
	gradient tensors and fp64 sums of the rest), enough to catch any operator going wrong at depth."""
	from PuzzleLib.Models.Nets.LeNet import loadLeNet
	from PuzzleLib.Models.Nets.ResNet import loadResNet

	np.random.seed(4321)
	rng = np.random.RandomState(99)
	if which == "lenet":
		net = loadLeNet(None, initscheme="he")
		x = rng.randn(batch, 1, 28, 28).astype(np.float32)
	else:
		net = loadResNet(None, "50", initscheme="he")
		x = rng.randn(batch, 3, 224, 224).astype(np.float32)

	net.zeroGradParams()
	y = net(_g(B, x))
	dy = (rng.randn(*y.shape) * 1e-2).astype(np.float32)
	net.backward(_g(B, dy))

	dx = net.grad.get().astype(np.float64)
	out = {"y": y.get(), "dxsum": np.array(dx.sum()), "dxabs": np.array(np.abs(dx).sum()), "dxhead": dx.ravel()[:4096].astype(np.float32)}

	table = sorted((names[0], var) for var, names in net.getVarTable().items())
	out["names"] = np.array([name for name, _ in table])
	grads = [var.grad.get().astype(np.float64) for _, var in table]
	out["gradsum"] = np.array([g.sum() for g in grads])
	out["gradabs"] = np.array([np.abs(g).sum() for g in grads])
	for (name, _), g in zip(table, grads):
		if g.size <= 10000 and ("conv1" in name or "fc" in name or name.count(".") == 0 or g.ndim == 4 and g.size <= 4096):
			out["grad:" + name] = g.astype(np.float32)
	return out


NETS = {"lenet_n4": ("lenet", 4), "resnet50_n2": ("resnet50", 2)}


def runNets(B, names=None):
	out = {}
	for name in (NETS if names is None else names):
		for key, val in netCase(B, *NETS[name]).items():
			out["%s/%s" % (name, key)] = np.asarray(val)
	return out
