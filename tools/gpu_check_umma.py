"""First-light check of the tcgen05 GEMM / conv engine against numpy fp64 (run on the GPU box)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from puzzlelib_b200 import driver
from puzzlelib_b200.driver import lib, check, Conv2dDesc
from puzzlelib_b200.gpuarray import GPUArray
from ctypes import byref

def relerr(a, b):
	return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))

def gemm(M, N, K, ta, tb, alpha=1.0, beta=0.0, bias=False, seed=0):
	rng = np.random.default_rng(seed)
	A = rng.standard_normal((K, M) if ta else (M, K)).astype(np.float32)
	B = rng.standard_normal((N, K) if tb else (K, N)).astype(np.float32)
	C0 = rng.standard_normal((M, N)).astype(np.float32)
	bv = rng.standard_normal(N).astype(np.float32)
	dA, dB, dC, dbias = GPUArray.toGpu(A), GPUArray.toGpu(B), GPUArray.toGpu(C0), GPUArray.toGpu(bv)
	check(lib.pz_gemm(0, dA.ptr, dB.ptr, dC.ptr, M, N, K, A.shape[1], B.shape[1], N, int(ta), int(tb), alpha, beta,
		dbias.ptr if bias else None, None))
	ref = alpha * ((A.T if ta else A).astype(np.float64) @ (B.T if tb else B).astype(np.float64)) + beta * C0
	if bias: ref = ref + bv
	out = dC.get()
	e = relerr(out, ref)
	print("gemm M=%d N=%d K=%d ta=%d tb=%d alpha=%g beta=%g bias=%d  relerr=%.3e %s" % (M, N, K, ta, tb, alpha, beta, bias, e, "OK" if e < 2e-3 else "FAIL"))
	return e < 2e-3

def conv_ref(x, w, stride, pad, dil, groups):
	N, C, H, W = x.shape; K, Cg, R, S = w.shape
	P = (H + 2*pad - dil*(R-1) - 1)//stride + 1; Q = (W + 2*pad - dil*(S-1) - 1)//stride + 1
	xp = np.zeros((N, C, H+2*pad, W+2*pad), np.float64); xp[:, :, pad:pad+H, pad:pad+W] = x
	y = np.zeros((N, K, P, Q), np.float64); Kg = K // groups
	for g in range(groups):
		for r in range(R):
			for s in range(S):
				patch = xp[:, g*Cg:(g+1)*Cg, r*dil:r*dil+(P-1)*stride+1:stride, s*dil:s*dil+(Q-1)*stride+1:stride]
				y[:, g*Kg:(g+1)*Kg] += np.einsum("ncpq,kc->nkpq", patch, w[g*Kg:(g+1)*Kg, :, r, s].astype(np.float64))
	return y

def conv(N, C, H, W, K, R, stride, pad, dil=1, groups=1, bias=False, seed=0):
	rng = np.random.default_rng(seed)
	x = rng.standard_normal((N, C, H, W)).astype(np.float32)
	w = rng.standard_normal((K, C//groups, R, R)).astype(np.float32)
	b = rng.standard_normal(K).astype(np.float32)
	y = conv_ref(x, w, stride, pad, dil, groups)
	if bias: y = y + b[None, :, None, None]
	P, Q = y.shape[2:]
	d = Conv2dDesc(N, C, H, W, K, R, R, P, Q, stride, stride, pad, pad, dil, dil, groups)
	dx_, dw_, db_ = GPUArray.toGpu(x), GPUArray.toGpu(w), GPUArray.toGpu(b)
	dy_ = GPUArray.empty(y.shape, np.float32)
	check(lib.pz_conv2d_fprop(0, byref(d), dx_.ptr, dw_.ptr, db_.ptr if bias else None, dy_.ptr, None))
	e1 = relerr(dy_.get(), y)
	# dgrad / wgrad via autograd identities in fp64
	g = rng.standard_normal(y.shape).astype(np.float32)
	gg = GPUArray.toGpu(g)
	# reference dgrad: loop
	xp = np.zeros((N, C, H+2*pad, W+2*pad), np.float64)
	wg = np.zeros(w.shape, np.float64)
	xpad = np.zeros_like(xp); xpad[:, :, pad:pad+H, pad:pad+W] = x
	Cg, Kg = C//groups, K//groups
	for gi in range(groups):
		for r in range(R):
			for s in range(R):
				sl = (slice(None), slice(gi*Cg, (gi+1)*Cg), slice(r*dil, r*dil+(P-1)*stride+1, stride), slice(s*dil, s*dil+(Q-1)*stride+1, stride))
				xp[sl] += np.einsum("nkpq,kc->ncpq", g[:, gi*Kg:(gi+1)*Kg].astype(np.float64), w[gi*Kg:(gi+1)*Kg, :, r, s].astype(np.float64))
				wg[gi*Kg:(gi+1)*Kg, :, r, s] = np.einsum("nkpq,ncpq->kc", g[:, gi*Kg:(gi+1)*Kg].astype(np.float64), xpad[sl])
	dxref = xp[:, :, pad:pad+H, pad:pad+W]
	wsz = int(lib.pz_conv2d_dgrad_workspace(0, byref(d)))
	ws = GPUArray.empty((max(wsz, 4)//4,), np.float32)
	dxo = GPUArray.empty(x.shape, np.float32)
	check(lib.pz_conv2d_dgrad(0, byref(d), gg.ptr, dw_.ptr, None, dxo.ptr, ws.ptr, wsz, None))
	e2 = relerr(dxo.get(), dxref)
	w0 = rng.standard_normal(w.shape).astype(np.float32)
	dwo = GPUArray.toGpu(w0)
	check(lib.pz_conv2d_wgrad(0, byref(d), dx_.ptr, gg.ptr, dwo.ptr, 0.5, 0.25, None))
	e3 = relerr(dwo.get(), 0.5*wg + 0.25*w0)
	ok = max(e1, e2, e3) < 2e-3
	print("conv N=%d C=%d HW=%d K=%d R=%d s=%d p=%d d=%d g=%d bias=%d  fprop=%.2e dgrad=%.2e wgrad=%.2e %s" % (N, C, H, K, R, stride, pad, dil, groups, bias, e1, e2, e3, "OK" if ok else "FAIL"))
	return ok

if __name__ == "__main__":
	print(driver.Device(0).name(), driver.Device(0).computeCapability())
	ok = True
	ok &= gemm(128, 128, 32, 0, 1)
	ok &= gemm(128, 128, 64, 0, 1)
	ok &= gemm(64, 128, 256, 0, 0)
	ok &= gemm(100, 200, 77, 0, 0, bias=True)
	ok &= gemm(33, 1000, 2048, 0, 0, alpha=0.5, beta=0.5)
	ok &= gemm(300, 70, 129, 1, 0)
	ok &= gemm(300, 70, 129, 1, 1, alpha=2.0, beta=1.0)
	ok &= gemm(64, 1000, 2048, 0, 0, bias=True)
	ok &= conv(2, 8, 12, 12, 16, 1, 1, 0)
	ok &= conv(2, 8, 12, 12, 16, 3, 1, 1, bias=True)
	ok &= conv(2, 3, 20, 20, 8, 7, 2, 3)
	ok &= conv(2, 16, 11, 11, 32, 1, 2, 0)
	ok &= conv(3, 6, 9, 10, 4, 2, 1, 0, groups=2)
	ok &= conv(2, 4, 13, 13, 6, 3, 2, 1, dil=2)
	ok &= conv(4, 64, 14, 14, 64, 3, 1, 1)
	ok &= conv(2, 256, 7, 7, 512, 1, 1, 0)
	print("ALL OK" if ok else "SOME FAILED")
	# quick timing of a ResNet-ish conv
	from puzzlelib_b200.driver import Event
	for (N, C, H, K, R, s, p) in [(64, 64, 55, 64, 3, 1, 1), (64, 256, 55, 64, 1, 1, 0), (64, 512, 7, 512, 3, 1, 1), (64, 1024, 14, 256, 1, 1, 0)]:
		P = (H + 2*p - R)//s + 1
		d = Conv2dDesc(N, C, H, H, K, R, R, P, P, s, s, p, p, 1, 1, 1)
		x = GPUArray.zeros((N, C, H, H), np.float32); w = GPUArray.zeros((K, C, R, R), np.float32); y = GPUArray.empty((N, K, P, P), np.float32)
		for _ in range(3): check(lib.pz_conv2d_fprop(0, byref(d), x.ptr, w.ptr, None, y.ptr, None))
		e0, e1 = Event(), Event(); e0.record()
		for _ in range(10): check(lib.pz_conv2d_fprop(0, byref(d), x.ptr, w.ptr, None, y.ptr, None))
		e1.record(); e1.synchronize()
		ms = e0.timeTill(e1) / 10
		fl = 2.0 * N * K * P * P * C * R * R
		print("fprop N=%d C=%d H=%d K=%d R=%d: %.3f ms  %.1f TFLOP/s" % (N, C, H, K, R, ms, fl / ms / 1e9))
	sys.exit(0 if ok else 1)
