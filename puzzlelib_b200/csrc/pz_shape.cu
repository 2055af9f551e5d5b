// pz_shape.cu -- the kernel modules of the reference backend object that sit next to the hot path: PReLU, reflection padding,
// embedding lookup, nearest / linear up-sampling (reference: Cuda/Kernels/PRelu.py, Pad.py, Embedder.py, Upsample.py -- NVRTC
// kernels with one thread per element there).  All bandwidth-bound; grid-stride kernels sized from the SM count, plane-major
// indexing so that the per-element work is a handful of integer instructions, gather-form (atomic-free, deterministic) backward
// passes wherever the inverse map is closed-form.
#include "pz_common.h"

namespace {

constexpr int kThreads = 256;

inline unsigned grid_for(int64_t work)
{
	const int64_t blocks = pz_cdiv(work, kThreads), cap = (int64_t)pz_num_sms() * 8;
	return (unsigned)(blocks < 1 ? 1 : (blocks < cap ? blocks : cap));
}

template <typename T> __device__ __forceinline__ float ldf(const T* p) { return (float)*p; }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v) { *p = (T)v; }
template <> __device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = __float2half_rn(v); }

// ------------------------------------------------------------------------------------------ PReLU (fp32, PRelu.py:12-57)
// slope index of element i of an [N][C][S] tensor: channel (i / S) % C, or 0 when one slope is shared by all maps
__global__ void __launch_bounds__(kThreads) prelu_fwd_kernel(const float* __restrict__ x, const float* __restrict__ slopes, float* __restrict__ y,
															  int64_t total, int64_t S, int C, int shared)
{
	for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
		const int c = shared ? 0 : (int)((i / S) % C);
		const float v = x[i];
		y[i] = v * (v > 0.0f ? 1.0f : __ldg(slopes + c));
	}
}

__global__ void __launch_bounds__(kThreads) prelu_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ slopes,
																   const float* __restrict__ x, float* __restrict__ dx, int64_t total, int64_t S,
																   int C, int shared)
{
	for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
		const int c = shared ? 0 : (int)((i / S) % C);
		const float v = x[i];
		dx[i] = dy[i] * ((v > 0.0f ? 1.0f : 0.0f) + (v <= 0.0f ? 1.0f : 0.0f) * __ldg(slopes + c));
	}
}

// dslope[c] = sum over images and positions of dy * x * (x <= 0): one block per channel (all channels for a shared slope), fixed
// summation tree -- deterministic, unlike an atomic reduction
__global__ void __launch_bounds__(kThreads) prelu_bwd_params_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ ds,
																	 int64_t N, int C, int64_t S, int shared)
{
	__shared__ float red[kThreads / 32];
	const int c = blockIdx.x;
	const int64_t per_image = shared ? (int64_t)C * S : S, count = N * per_image;
	float acc = 0.0f;
	for (int64_t j = threadIdx.x; j < count; j += kThreads) {
		const int64_t n = j / per_image, r = j - n * per_image;
		const int64_t idx = n * C * S + (shared ? r : (int64_t)c * S + r);
		const float v = x[idx];
		acc += dy[idx] * v * (v <= 0.0f ? 1.0f : 0.0f);
	}
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x == 0) {
		float total = 0.0f;
		for (int w = 0; w < kThreads / 32; w++) total += red[w];
		ds[c] = total;
	}
}

// ------------------------------------------------------------------------------------------ reflection padding (Pad.py:33-142)
// numpy "reflect": out[i] = in[mirror(i - pad)] with the edge sample not repeated.  1-d tensors are planes of height 1.
__device__ __forceinline__ int mirror(int i, int n)
{
	if (i < 0) i = -i;
	if (i >= n) i = 2 * (n - 1) - i;
	return i;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) reflectpad_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t planes, int H, int W, int up,
																   int lp, int oH, int oW)
{
	const int64_t osize = (int64_t)oH * oW, total = planes * osize;
	for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
		const int64_t p = i / osize;
		const int r = (int)(i - p * osize), oy = r / oW, ox = r - oy * oW;
		y[i] = x[p * H * W + (int64_t)mirror(oy - up, H) * W + mirror(ox - lp, W)];
	}
}

// gather form of the adjoint: input position j receives its own output sample plus the mirrored ones -- the left mirror when
// 1 <= j <= lpad, the right one when n-1-rpad <= j <= n-2 (the reference scatters with atomicAdd, Pad.py:77-139)
__device__ __forceinline__ int mirror_sources(int j, int n, int before, int after, int (&src)[3])
{
	int k = 0;
	src[k++] = j + before;
	if (j >= 1 && j <= before) src[k++] = before - j;
	if (j <= n - 2 && j >= n - 1 - after) src[k++] = before + 2 * (n - 1) - j;
	return k;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) reflectpad_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int64_t planes, int H, int W, int up,
																   int bp, int lp, int rp, int oH, int oW)
{
	const int64_t isize = (int64_t)H * W, total = planes * isize;
	for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
		const int64_t p = i / isize;
		const int r = (int)(i - p * isize), iy = r / W, ix = r - iy * W;
		int ys[3], xs[3];
		const int ny = mirror_sources(iy, H, up, bp, ys), nx = mirror_sources(ix, W, lp, rp, xs);
		const T* g = dy + p * (int64_t)oH * oW;
		float acc = 0.0f;
		for (int a = 0; a < ny; a++)
			for (int b = 0; b < nx; b++) acc += ldf(g + (int64_t)ys[a] * oW + xs[b]);
		stf(dx + i, acc);
	}
}

// ------------------------------------------------------------------------------------------ embedding lookup (Embedder.py:11-42)
// index -1 = padding: the output row is zero (the reference zero-fills first and skips the row), no gradient
template <typename T>
__global__ void __launch_bounds__(kThreads) embed_fwd_kernel(const int* __restrict__ idx, const T* __restrict__ W, T* __restrict__ out, int64_t size,
															  int64_t emb)
{
	const int64_t total = size * emb;
	for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
		const int64_t row = i / emb, col = i - row * emb;
		const int word = __ldg(idx + row);
		out[i] = word == -1 ? (T)0.0f : W[(int64_t)word * emb + col];
	}
}

__device__ __forceinline__ void atomic_add(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add(__half* p, float v) { atomicAdd(p, __float2half_rn(v)); }

template <typename T>
__global__ void __launch_bounds__(kThreads) embed_bwd_kernel(const int* __restrict__ idx, const T* __restrict__ grad, T* W, float scale, int64_t size,
															  int64_t emb)
{
	const int64_t total = size * emb;
	for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
		const int64_t row = i / emb, col = i - row * emb;
		const int word = __ldg(idx + row);
		if (word == -1) continue;
		// the vocabulary update W[word] += scale * grad: rows of repeated words collide, hence atomics (as the reference)
		atomic_add(W + (int64_t)word * emb + col, scale * ldf(grad + i));
	}
}

// ------------------------------------------------------------------------------------------ up-sampling (Upsample.py:9-297), fp32
// nearest: out[z][d*ds+i][y*hs+j][x*ws+k] = in[z][d][y][x]; one thread per OUTPUT element (coalesced stores, the input read hits L1)
__global__ void __launch_bounds__(kThreads) upsample_nearest_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t planes, int D, int H,
																		 int W, int ds, int hs, int ws)
{
	const int oD = D * ds, oH = H * hs, oW = W * ws;
	const int64_t osize = (int64_t)oD * oH * oW, total = planes * osize;
	for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
		const int64_t p = i / osize;
		int64_t r = i - p * osize;
		const int od = (int)(r / ((int64_t)oH * oW));
		r -= (int64_t)od * oH * oW;
		const int oy = (int)(r / oW), ox = (int)(r - (int64_t)oy * oW);
		y[i] = x[((p * D + od / ds) * H + oy / hs) * W + ox / ws];
	}
}

// the adjoint: every input element sums its ds x hs x ws block of output gradients
__global__ void __launch_bounds__(kThreads) upsample_nearest_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int64_t planes, int D, int H,
																		 int W, int ds, int hs, int ws)
{
	const int oH = H * hs, oW = W * ws;
	const int64_t isize = (int64_t)D * H * W, total = planes * isize;
	for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
		const int64_t p = i / isize;
		int64_t r = i - p * isize;
		const int d = (int)(r / ((int64_t)H * W));
		r -= (int64_t)d * H * W;
		const int yy = (int)(r / W), xx = (int)(r - (int64_t)yy * W);
		const float* g = dy + ((p * D * ds + (int64_t)d * ds) * oH + (int64_t)yy * hs) * oW + (int64_t)xx * ws;
		float acc = 0.0f;
		for (int a = 0; a < ds; a++)
			for (int b = 0; b < hs; b++)
				for (int c = 0; c < ws; c++) acc += g[((int64_t)a * oH + b) * oW + c];
		dx[i] = acc;
	}
}

// linear ("align corners": source coordinate = r * output coordinate, r = (in - 1) / (out - 1) as float32).  The interpolation
// expressions keep the reference's evaluation order.  Quirk kept on purpose: one of the eight taps of the reference's 3-d FORWARD
// kernel is addressed with d1 * inw * inw instead of d1 * inh * inw (Upsample.py:241) -- identical whenever inh == inw.
struct Lerp {
	int i0, step;
	float w0, w1;
};
__device__ __forceinline__ Lerp lerp_of(float ratio, int o, int n)
{
	Lerp l;
	const float src = ratio * (float)o;
	l.i0 = (int)src;
	l.step = l.i0 < n - 1 ? 1 : 0;
	l.w1 = src - (float)l.i0;
	l.w0 = 1.0f - l.w1;
	return l;
}

__global__ void __launch_bounds__(kThreads) upsample_linear_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t planes, int D, int H,
																		int W, int oD, int oH, int oW, float rd, float rh, float rw, int three_d)
{
	const int64_t osize = (int64_t)oD * oH * oW, total = planes * osize;
	for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
		const int64_t p = i / osize;
		int64_t r = i - p * osize;
		const int od = (int)(r / ((int64_t)oH * oW));
		r -= (int64_t)od * oH * oW;
		const int oy = (int)(r / oW), ox = (int)(r - (int64_t)oy * oW);
		const Lerp lh = lerp_of(rh, oy, H), lw = lerp_of(rw, ox, W);
		const float* src = x + p * (int64_t)D * H * W;
		if (!three_d) {
			const float* a = src + (int64_t)lh.i0 * W + lw.i0;
			const float* b = src + (int64_t)(lh.i0 + lh.step) * W + lw.i0;
			y[i] = lh.w0 * (lw.w0 * a[0] + lw.w1 * a[lw.step]) + lh.w1 * (lw.w0 * b[0] + lw.w1 * b[lw.step]);
		} else {
			const Lerp ld = lerp_of(rd, od, D);
			const int64_t hw = (int64_t)H * W, d0 = (int64_t)ld.i0 * hw, d1 = (int64_t)(ld.i0 + ld.step) * hw;
			const int64_t h0 = (int64_t)lh.i0 * W, h1 = (int64_t)(lh.i0 + lh.step) * W;
			int64_t quirk = (int64_t)ld.i0 * W * W + h0 + lw.i0 + lw.step;      // Upsample.py:241
			if (quirk >= (int64_t)D * hw) quirk = d0 + h0 + lw.i0 + lw.step;       // (the reference would read past the volume there)
			const float near =
				lh.w0 * (lw.w0 * src[d0 + h0 + lw.i0] + lw.w1 * src[quirk]) +
				lh.w1 * (lw.w0 * src[d0 + h1 + lw.i0] + lw.w1 * src[d0 + h1 + lw.i0 + lw.step]);
			const float far =
				lh.w0 * (lw.w0 * src[d1 + h0 + lw.i0] + lw.w1 * src[d1 + h0 + lw.i0 + lw.step]) +
				lh.w1 * (lw.w0 * src[d1 + h1 + lw.i0] + lw.w1 * src[d1 + h1 + lw.i0 + lw.step]);
			y[i] = ld.w0 * near + ld.w1 * far;
		}
	}
}

// the adjoint scatters each output gradient to its 4 / 8 taps (red.add: the taps of neighbouring outputs collide, as in the
// reference, Upsample.py:141-184, 254-296); dx is zeroed by the caller
__global__ void __launch_bounds__(kThreads) upsample_linear_bwd_kernel(const float* __restrict__ dy, float* dx, int64_t planes, int D, int H, int W,
																		int oD, int oH, int oW, float rd, float rh, float rw, int three_d)
{
	const int64_t osize = (int64_t)oD * oH * oW, total = planes * osize;
	for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
		const int64_t p = i / osize;
		int64_t r = i - p * osize;
		const int od = (int)(r / ((int64_t)oH * oW));
		r -= (int64_t)od * oH * oW;
		const int oy = (int)(r / oW), ox = (int)(r - (int64_t)oy * oW);
		const Lerp lh = lerp_of(rh, oy, H), lw = lerp_of(rw, ox, W);
		float* dst = dx + p * (int64_t)D * H * W;
		const float v = dy[i];
		const int64_t h0 = (int64_t)lh.i0 * W, h1 = (int64_t)(lh.i0 + lh.step) * W;
		if (!three_d) {
			atomicAdd(dst + h0 + lw.i0, lh.w0 * lw.w0 * v);
			atomicAdd(dst + h0 + lw.i0 + lw.step, lh.w0 * lw.w1 * v);
			atomicAdd(dst + h1 + lw.i0, lh.w1 * lw.w0 * v);
			atomicAdd(dst + h1 + lw.i0 + lw.step, lh.w1 * lw.w1 * v);
		} else {
			const Lerp ld = lerp_of(rd, od, D);
			const int64_t hw = (int64_t)H * W, d0 = (int64_t)ld.i0 * hw, d1 = (int64_t)(ld.i0 + ld.step) * hw;
			atomicAdd(dst + d0 + h0 + lw.i0, ld.w0 * lh.w0 * lw.w0 * v);
			atomicAdd(dst + d0 + h0 + lw.i0 + lw.step, ld.w0 * lh.w0 * lw.w1 * v);
			atomicAdd(dst + d0 + h1 + lw.i0, ld.w0 * lh.w1 * lw.w0 * v);
			atomicAdd(dst + d0 + h1 + lw.i0 + lw.step, ld.w0 * lh.w1 * lw.w1 * v);
			atomicAdd(dst + d1 + h0 + lw.i0, ld.w1 * lh.w0 * lw.w0 * v);
			atomicAdd(dst + d1 + h0 + lw.i0 + lw.step, ld.w1 * lh.w0 * lw.w1 * v);
			atomicAdd(dst + d1 + h1 + lw.i0, ld.w1 * lh.w1 * lw.w0 * v);
			atomicAdd(dst + d1 + h1 + lw.i0 + lw.step, ld.w1 * lh.w1 * lw.w1 * v);
		}
	}
}

// ------------------------------------------------------------------------------------------ spatial transformer
// cudnnSpatialTfGridGenerator* + cudnnSpatialTfSampler* (CuDnnSpatialTf.c:20-222; host formulas of Cuda/Wrappers/CuDnnSpatialTf.py):
//   grid[b][y][x] = theta[b] (2 x 3) . (xn, yn, 1),  xn = -1 + 2 x / (oW - 1),  yn = -1 + 2 y / (oH - 1)
//   out[b][c][y][x] = bilinear sample of data[b][c] at ((gx + 1) (W - 1) / 2, (gy + 1) (H - 1) / 2), zero outside the map
__device__ __forceinline__ float norm_coord(int i, int n) { return -1.0f + (float)i * (2.0f / (float)(n - 1)); }

template <typename T>
__global__ void __launch_bounds__(kThreads) stf_grid_kernel(const T* __restrict__ theta, T* __restrict__ grid, int64_t B, int oH, int oW)
{
	const int64_t osize = (int64_t)oH * oW, total = B * osize;
	for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
		const int64_t b = i / osize;
		const int r = (int)(i - b * osize), y = r / oW, x = r - y * oW;
		const T* t = theta + b * 6;
		const float xn = norm_coord(x, oW), yn = norm_coord(y, oH);
		stf(grid + 2 * i, ldf(t) * xn + ldf(t + 1) * yn + ldf(t + 2));
		stf(grid + 2 * i + 1, ldf(t + 3) * xn + ldf(t + 4) * yn + ldf(t + 5));
	}
}

struct Taps {
	int x0, y0;
	float fx, fy;
	bool okx0, okx1, oky0, oky1;
};
__device__ __forceinline__ Taps taps_of(float gx, float gy, int H, int W)
{
	Taps t;
	const float nx = (gx + 1.0f) * (0.5f * (float)(W - 1)), ny = (gy + 1.0f) * (0.5f * (float)(H - 1));
	const float flx = floorf(nx), fly = floorf(ny);
	t.x0 = (int)flx; t.y0 = (int)fly;
	t.fx = nx - flx; t.fy = ny - fly;
	t.okx0 = t.x0 >= 0 && t.x0 < W; t.okx1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
	t.oky0 = t.y0 >= 0 && t.y0 < H; t.oky1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
	return t;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) stf_sample_fwd_kernel(const T* __restrict__ data, const T* __restrict__ grid, T* __restrict__ out, int64_t B,
																   int C, int H, int W, int oH, int oW)
{
	const int64_t osize = (int64_t)oH * oW, total = B * C * osize;
	for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
		const int64_t bc = i / osize, b = bc / C, pos = i - bc * osize;
		const T* g = grid + 2 * (b * osize + pos);
		const Taps t = taps_of(ldf(g), ldf(g + 1), H, W);
		const T* src = data + bc * (int64_t)H * W + (int64_t)t.y0 * W + t.x0;
		float v = 0.0f;
		if (t.oky0 && t.okx0) v += ldf(src) * (1.0f - t.fy) * (1.0f - t.fx);
		if (t.oky0 && t.okx1) v += ldf(src + 1) * (1.0f - t.fy) * t.fx;
		if (t.oky1 && t.okx0) v += ldf(src + W) * t.fy * (1.0f - t.fx);
		if (t.oky1 && t.okx1) v += ldf(src + W + 1) * t.fy * t.fx;
		stf(out + i, v);
	}
}

// one thread per (image, output position): the taps are shared by the channels; the data gradient is scattered (red.add into a
// zeroed tensor, as taps of neighbouring outputs collide), the grid gradient is summed over the channels in registers
template <typename T>
__global__ void __launch_bounds__(kThreads) stf_sample_bwd_kernel(const T* __restrict__ grad, const T* __restrict__ data, const T* __restrict__ grid,
																   T* dx, T* __restrict__ dgrid, int64_t B, int C, int H, int W, int oH, int oW)
{
	const int64_t osize = (int64_t)oH * oW, total = B * osize;
	const float sx = 0.5f * (float)(W - 1), sy = 0.5f * (float)(H - 1);          // d(nx) / d(gx), d(ny) / d(gy)
	for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
		const int64_t b = i / osize, pos = i - b * osize;
		const Taps t = taps_of(ldf(grid + 2 * i), ldf(grid + 2 * i + 1), H, W);
		float gx = 0.0f, gy = 0.0f;
		for (int c = 0; c < C; c++) {
			const int64_t plane = (b * C + c) * (int64_t)H * W + (int64_t)t.y0 * W + t.x0;
			const float g = ldf(grad + (b * C + c) * osize + pos);
			float vx = 0.0f, vy = 0.0f;
			if (t.oky0 && t.okx0) { const float d = ldf(data + plane); atomic_add(dx + plane, g * (1.0f - t.fy) * (1.0f - t.fx)); vx -= d * (1.0f - t.fy); vy -= d * (1.0f - t.fx); }
			if (t.oky0 && t.okx1) { const float d = ldf(data + plane + 1); atomic_add(dx + plane + 1, g * (1.0f - t.fy) * t.fx); vx += d * (1.0f - t.fy); vy -= d * t.fx; }
			if (t.oky1 && t.okx0) { const float d = ldf(data + plane + W); atomic_add(dx + plane + W, g * t.fy * (1.0f - t.fx)); vx -= d * t.fy; vy += d * (1.0f - t.fx); }
			if (t.oky1 && t.okx1) { const float d = ldf(data + plane + W + 1); atomic_add(dx + plane + W + 1, g * t.fy * t.fx); vx += d * t.fy; vy += d * t.fx; }
			gx = fmaf(g, vx * sx, gx);
			gy = fmaf(g, vy * sy, gy);
		}
		stf(dgrid + 2 * i, gx);
		stf(dgrid + 2 * i + 1, gy);
	}
}

// dtheta[b] = sum over positions of outer(dgrid[b][y][x], (xn, yn, 1)): one block per image, fixed summation tree
template <typename T>
__global__ void __launch_bounds__(kThreads) stf_dtheta_kernel(const T* __restrict__ dgrid, T* __restrict__ dtheta, int oH, int oW)
{
	__shared__ float red[6][kThreads / 32];
	const int64_t b = blockIdx.x;
	const int osize = oH * oW;
	float acc[6] = {};
	for (int r = threadIdx.x; r < osize; r += kThreads) {
		const int y = r / oW, x = r - y * oW;
		const float xn = norm_coord(x, oW), yn = norm_coord(y, oH);
		const float gx = ldf(dgrid + 2 * (b * osize + r)), gy = ldf(dgrid + 2 * (b * osize + r) + 1);
		acc[0] += gx * xn; acc[1] += gx * yn; acc[2] += gx;
		acc[3] += gy * xn; acc[4] += gy * yn; acc[5] += gy;
	}
	for (int k = 0; k < 6; k++) {
		float v = acc[k];
		for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
		if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = v;
	}
	__syncthreads();
	if (threadIdx.x < 6) {
		float v = 0.0f;
		for (int w = 0; w < kThreads / 32; w++) v += red[threadIdx.x][w];
		stf(dtheta + b * 6 + threadIdx.x, v);
	}
}

// ------------------------------------------------------------------------------------------ CTC loss (Cuda/Kernels/CTC.py:9-192)
// One block per sequence; extended label string l' (blanks interleaved, S = 2 L + 1 states) in shared memory.  alphas [T][S] go to
// global memory in the reference's layout (sequence b at T * (2 * offset_b + b)), betas are double-buffered in shared memory.
// The gradient with respect to the SOFTMAX OUTPUT y:  grad[t][v] = -y[t][v] + sum_{s: l'_s = v} exp(alpha + beta - log y + nll).
// States that share a label are chained in ascending order (`nxt`) instead of the reference's block radix sort: the log-sum over a
// label visits them in the same order.
constexpr int kCtcThreads = 128;
#define PZ_NEG_INF (-__int_as_float(0x7f800000))

__device__ __forceinline__ float log_plus(float a, float b)
{
	if (a <= PZ_NEG_INF) return b;
	if (b <= PZ_NEG_INF) return a;
	return log1pf(expf(-fabsf(a - b))) + fmaxf(a, b);
}

__global__ void __launch_bounds__(kCtcThreads) ctc_alpha_kernel(const float* __restrict__ y, const int* __restrict__ datalen, int T, int B, int V,
																 const int* __restrict__ labels, const int* __restrict__ offsets, float* __restrict__ alphas,
																 int blank, float* __restrict__ nll, float* error)
{
	extern __shared__ int ctc_smem[];
	int* shl = ctc_smem;
	const int b = blockIdx.x, offset = offsets[b], S = 2 * (offsets[b + 1] - offset) + 1;
	y += (int64_t)b * V;
	labels += offset;
	float* a = alphas + (int64_t)T * (2 * offset + b);

	for (int i = threadIdx.x; i < S; i += kCtcThreads) {
		const int label = (i % 2 == 0) ? blank : labels[i / 2];
		shl[i] = label;
		a[i] = i < 2 ? logf(y[label]) : PZ_NEG_INF;
	}
	__syncthreads();
	const int Tb = datalen[b];
	for (int t = 1; t < Tb; t++) {
		for (int i = threadIdx.x; i < S; i += kCtcThreads) {
			float prev = a[(int64_t)(t - 1) * S + i];
			if (i > 0) {
				prev = log_plus(prev, a[(int64_t)(t - 1) * S + i - 1]);
				if (i > 1 && shl[i] != blank && shl[i] != shl[i - 2]) prev = log_plus(prev, a[(int64_t)(t - 1) * S + i - 2]);
			}
			a[(int64_t)t * S + i] = prev + logf(y[(int64_t)t * B * V + shl[i]]);
		}
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		const float tail = S >= 2 ? a[(int64_t)(Tb - 1) * S + S - 2] : PZ_NEG_INF;
		const float loglike = log_plus(tail, a[(int64_t)(Tb - 1) * S + S - 1]);
		nll[b] = -loglike;
		atomicAdd(error, -loglike);
	}
}

__global__ void __launch_bounds__(kCtcThreads) ctc_beta_kernel(const float* __restrict__ y, const int* __restrict__ datalen, int T, int B, int V,
																const int* __restrict__ labels, const int* __restrict__ offsets, const float* __restrict__ alphas,
																int blank, const float* __restrict__ nll, float* __restrict__ grad, int Smax)
{
	extern __shared__ int ctc_smem[];
	int* shl = ctc_smem;                       // [Smax] extended labels
	int* nxt = ctc_smem + Smax;                // [Smax] 1 + next state with the same label (0: none) | first occurrence << 30
	float* betas = reinterpret_cast<float*>(ctc_smem + 2 * Smax);      // [2][Smax]
	const int b = blockIdx.x, offset = offsets[b], S = 2 * (offsets[b + 1] - offset) + 1;
	y += (int64_t)b * V;
	grad += (int64_t)b * V;
	labels += offset;
	const float* a = alphas + (int64_t)T * (2 * offset + b);
	const float loglike = nll[b];
	if (loglike >= -PZ_NEG_INF) return;       // no valid alignment: the gradient stays zero

	for (int i = threadIdx.x; i < S; i += kCtcThreads) shl[i] = (i % 2 == 0) ? blank : labels[i / 2];
	__syncthreads();
	for (int i = threadIdx.x; i < S; i += kCtcThreads) {
		const int label = shl[i];
		int next = -1, first = 1;
		for (int j = i + 1; j < S; j++)
			if (shl[j] == label) { next = j; break; }
		for (int j = i - 1; j >= 0; j--)
			if (shl[j] == label) { first = 0; break; }
		nxt[i] = (next + 1) | (first << 30);                       // 0 in the low bits: no further state with this label
	}
	__syncthreads();

	const int Tb = datalen[b];
	int src = 0, dst = 1;
	for (int t = Tb - 1; t >= 0; t--) {
		if (t < Tb - 1) {
			for (int i = threadIdx.x; i < S; i += kCtcThreads) {
				float next = betas[src * Smax + i];
				if (i < S - 1) {
					next = log_plus(next, betas[src * Smax + i + 1]);
					if (i < S - 2 && shl[i] != blank && shl[i] != shl[i + 2]) next = log_plus(next, betas[src * Smax + i + 2]);
				}
				betas[dst * Smax + i] = next + logf(y[(int64_t)t * B * V + shl[i]]);
			}
			src ^= 1;
			dst ^= 1;
		} else {
			for (int i = threadIdx.x; i < S; i += kCtcThreads)
				betas[i] = i >= S - 2 ? logf(y[(int64_t)(Tb - 1) * B * V + shl[i]]) : PZ_NEG_INF;
		}
		__syncthreads();
		for (int v = threadIdx.x; v < V; v += kCtcThreads) grad[(int64_t)t * B * V + v] = -y[(int64_t)t * B * V + v];
		__syncthreads();
		for (int i = threadIdx.x; i < S; i += kCtcThreads) {
			if (!(nxt[i] >> 30 & 1)) continue;                          // the first state of a label sums the whole chain
			float gr = PZ_NEG_INF;
			for (int j = i; j >= 0; j = (nxt[j] & ~(1 << 30)) - 1) gr = log_plus(gr, a[(int64_t)t * S + j] + betas[src * Smax + j]);
			const int64_t off = (int64_t)t * B * V + shl[i];
			const float data = y[off];
			if (data > 0.0f) grad[off] += expf(gr - logf(data) + loglike);
		}
		__syncthreads();
	}
}

// ------------------------------------------------------------------------------------------ divisive normalisation (LCN)
// mapLRN with a means tensor = cudnnDivisiveNormalization (CuDnnNorm.c:329-510; host formulas of Modules/LCN.py:62-143):
//   norm_i = K + alpha / N^2 * sum_{j in win(i)} (x_j - m_i)^2,   y_i = x_i * norm_i^-beta,   win(i) = [i - lb, i + la) clipped
//   t_j    = g_j * x_j * norm_j^-(beta + 1)
//   dx_i   = g_i * norm_i^-beta - (2 alpha beta / N^2) * (x_i * sum_{j in win(i)} t_j - sum_{j in win(i)} t_j * m_j)
//   dm_i   = (2 alpha beta / N^2) * t_i * sum_{j in win(i)} (x_j - m_i)
struct DivGeo {
	int H, W, lb, la;
	float scale, beta, K;          // scale = alpha / N^2
};

template <typename T>
__device__ __forceinline__ float div_norm(const T* __restrict__ xp, float m, const DivGeo& g, int h, int w, float& sumdiff)
{
	float s = 0.0f, d = 0.0f;
	for (int yy = max(0, h - g.lb); yy < min(g.H, h + g.la); yy++)
		for (int xx = max(0, w - g.lb); xx < min(g.W, w + g.la); xx++) {
			const float v = ldf(xp + (int64_t)yy * g.W + xx) - m;
			s = fmaf(v, v, s);
			d += v;
		}
	sumdiff = d;
	return g.K + g.scale * s;
}

// PASS 0: y.  PASS 1: t (fp32 scratch) and dm.  PASS 2: dx from g, x, m, t.
template <typename T, int PASS>
__global__ void __launch_bounds__(kThreads) divnorm_kernel(const T* __restrict__ x, const T* __restrict__ means, const T* __restrict__ grad,
															T* __restrict__ out, T* __restrict__ dmeans, float* __restrict__ tmp, DivGeo g, int64_t total)
{
	const int64_t plane = (int64_t)g.H * g.W;
	for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
		const int64_t p = i / plane;
		const int r = (int)(i - p * plane), h = r / g.W, w = r - h * g.W;
		const T* xp = x + p * plane;
		float sumdiff;
		const float xv = ldf(x + i), norm = div_norm(xp, ldf(means + i), g, h, w, sumdiff);
		if (PASS == 0) {
			stf(out + i, xv * powf(norm, -g.beta));
		} else if (PASS == 1) {
			const float t = ldf(grad + i) * xv * powf(norm, -(g.beta + 1.0f));
			tmp[i] = t;
			stf(dmeans + i, 2.0f * g.beta * g.scale * t * sumdiff);
		} else {
			const float* tp = tmp + p * plane;
			const T* mp = means + p * plane;
			float s1 = 0.0f, s2 = 0.0f;
			for (int yy = max(0, h - g.lb); yy < min(g.H, h + g.la); yy++)
				for (int xx = max(0, w - g.lb); xx < min(g.W, w + g.la); xx++) {
					const float t = tp[(int64_t)yy * g.W + xx];
					s1 += t;
					s2 = fmaf(t, ldf(mp + (int64_t)yy * g.W + xx), s2);
				}
			stf(out + i, ldf(grad + i) * powf(norm, -g.beta) - 2.0f * g.beta * g.scale * (xv * s1 - s2));
		}
	}
}

template <typename T>
int divnorm_launch(int pass, const void* x, const void* means, const void* grad, void* out, void* dmeans, float* tmp, const DivGeo& g, int64_t total,
				   cudaStream_t s)
{
	const unsigned grid = grid_for(total);
	if (pass == 0) divnorm_kernel<T, 0><<<grid, kThreads, 0, s>>>((const T*)x, (const T*)means, nullptr, (T*)out, nullptr, nullptr, g, total);
	else {
		divnorm_kernel<T, 1><<<grid, kThreads, 0, s>>>((const T*)x, (const T*)means, (const T*)grad, nullptr, (T*)dmeans, tmp, g, total);
		divnorm_kernel<T, 2><<<grid, kThreads, 0, s>>>((const T*)x, (const T*)means, (const T*)grad, (T*)out, nullptr, tmp, g, total);
		pz_count_launch(1);
	}
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

int div_geo(DivGeo& g, int dtype, int64_t planes, int64_t H, int64_t W, int n, float alpha, float beta, float K)
{
	PZ_REQUIRE(dtype == PZ_F32 || dtype == PZ_F16, "divisive normalisation: unsupported dtype %d", dtype);
	PZ_REQUIRE(planes >= 0 && n >= 1 && H > 0 && W > 0 && H < (1ll << 31) && W < (1ll << 31), "divisive normalisation: bad geometry");
	g.H = (int)H; g.W = (int)W;
	g.lb = (n - 1) / 2; g.la = n - g.lb;
	g.scale = alpha / ((float)n * (float)n);
	g.beta = beta; g.K = K;
	return PZ_OK;
}

}  // namespace

#define PZ_SHAPE_LAUNCH(KERNEL, WORK, ...)                                                   \
	do {                                                                                     \
		KERNEL<<<grid_for(WORK), kThreads, 0, pz_stream(stream)>>>(__VA_ARGS__);             \
		pz_count_launch(1);                                                                  \
		PZ_LAUNCH_CHECK();                                                                   \
	} while (0)

extern "C" {

int pz_prelu_fwd(const void* x, const void* slopes, void* y, int64_t N, int64_t C, int64_t S, int shared, void* stream)
{
	PZ_REQUIRE(N >= 0 && C > 0 && S > 0 && C < (1ll << 31), "prelu: invalid shape");
	const int64_t total = N * C * S;
	if (total == 0) return PZ_OK;
	PzProfScope prof(PZ_PROF_ELTWISE, pz_stream(stream), 0.0, 8.0 * (double)total);
	PZ_SHAPE_LAUNCH(prelu_fwd_kernel, total, (const float*)x, (const float*)slopes, (float*)y, total, S, (int)C, shared);
	return PZ_OK;
}

int pz_prelu_bwd_data(const void* dy, const void* slopes, const void* x, void* dx, int64_t N, int64_t C, int64_t S, int shared, void* stream)
{
	PZ_REQUIRE(N >= 0 && C > 0 && S > 0 && C < (1ll << 31), "prelu: invalid shape");
	const int64_t total = N * C * S;
	if (total == 0) return PZ_OK;
	PzProfScope prof(PZ_PROF_ELTWISE, pz_stream(stream), 0.0, 12.0 * (double)total);
	PZ_SHAPE_LAUNCH(prelu_bwd_data_kernel, total, (const float*)dy, (const float*)slopes, (const float*)x, (float*)dx, total, S, (int)C, shared);
	return PZ_OK;
}

int pz_prelu_bwd_params(const void* x, const void* dy, void* dslopes, int64_t N, int64_t C, int64_t S, int shared, void* stream)
{
	PZ_REQUIRE(N >= 0 && C > 0 && S > 0 && C < (1ll << 31), "prelu: invalid shape");
	PzProfScope prof(PZ_PROF_ELTWISE, pz_stream(stream), 0.0, 8.0 * (double)(N * C * S));
	prelu_bwd_params_kernel<<<(unsigned)(shared ? 1 : C), kThreads, 0, pz_stream(stream)>>>((const float*)x, (const float*)dy, (float*)dslopes, N, (int)C,
																						  S, shared);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

static int check_pad(int dtype, int64_t planes, int H, int W, int up, int bp, int lp, int rp)
{
	PZ_REQUIRE(dtype == PZ_F32 || dtype == PZ_F16, "reflectpad: unsupported dtype %d", dtype);
	PZ_REQUIRE(planes >= 0 && H > 0 && W > 0, "reflectpad: invalid shape");
	PZ_REQUIRE(up >= 0 && bp >= 0 && lp >= 0 && rp >= 0, "reflectpad: negative padding is not supported");
	PZ_REQUIRE(H >= (up > bp ? up : bp) + 1 || (up == 0 && bp == 0), "reflectpad: padding exceeds the map height");
	PZ_REQUIRE(W >= (lp > rp ? lp : rp) + 1, "reflectpad: padding exceeds the map width");
	return PZ_OK;
}

int pz_reflectpad_fwd(int dtype, const void* x, void* y, int64_t planes, int H, int W, int up, int bp, int lp, int rp, void* stream)
{
	int st = check_pad(dtype, planes, H, W, up, bp, lp, rp);
	if (st != PZ_OK) return st;
	const int oH = H + up + bp, oW = W + lp + rp;
	const int64_t total = planes * oH * oW;
	if (total == 0) return PZ_OK;
	if (dtype == PZ_F32) PZ_SHAPE_LAUNCH(reflectpad_fwd_kernel<float>, total, (const float*)x, (float*)y, planes, H, W, up, lp, oH, oW);
	else PZ_SHAPE_LAUNCH(reflectpad_fwd_kernel<__half>, total, (const __half*)x, (__half*)y, planes, H, W, up, lp, oH, oW);
	return PZ_OK;
}

int pz_reflectpad_bwd(int dtype, const void* dy, void* dx, int64_t planes, int H, int W, int up, int bp, int lp, int rp, void* stream)
{
	int st = check_pad(dtype, planes, H, W, up, bp, lp, rp);
	if (st != PZ_OK) return st;
	const int oH = H + up + bp, oW = W + lp + rp;
	const int64_t total = planes * H * W;
	if (total == 0) return PZ_OK;
	if (dtype == PZ_F32) PZ_SHAPE_LAUNCH(reflectpad_bwd_kernel<float>, total, (const float*)dy, (float*)dx, planes, H, W, up, bp, lp, rp, oH, oW);
	else PZ_SHAPE_LAUNCH(reflectpad_bwd_kernel<__half>, total, (const __half*)dy, (__half*)dx, planes, H, W, up, bp, lp, rp, oH, oW);
	return PZ_OK;
}

int pz_embed_fwd(int dtype, const void* idx, const void* W, void* out, int64_t size, int64_t emb, void* stream)
{
	PZ_REQUIRE(dtype == PZ_F32 || dtype == PZ_F16, "embed: unsupported dtype %d", dtype);
	PZ_REQUIRE(size >= 0 && emb > 0, "embed: invalid shape");
	if (size == 0) return PZ_OK;
	if (dtype == PZ_F32) PZ_SHAPE_LAUNCH(embed_fwd_kernel<float>, size * emb, (const int*)idx, (const float*)W, (float*)out, size, emb);
	else PZ_SHAPE_LAUNCH(embed_fwd_kernel<__half>, size * emb, (const int*)idx, (const __half*)W, (__half*)out, size, emb);
	return PZ_OK;
}

int pz_embed_bwd(int dtype, const void* idx, const void* grad, void* W, float scale, int64_t size, int64_t emb, void* stream)
{
	PZ_REQUIRE(dtype == PZ_F32 || dtype == PZ_F16, "embed: unsupported dtype %d", dtype);
	PZ_REQUIRE(size >= 0 && emb > 0, "embed: invalid shape");
	if (size == 0) return PZ_OK;
	if (dtype == PZ_F32) PZ_SHAPE_LAUNCH(embed_bwd_kernel<float>, size * emb, (const int*)idx, (const float*)grad, (float*)W, scale, size, emb);
	else PZ_SHAPE_LAUNCH(embed_bwd_kernel<__half>, size * emb, (const int*)idx, (const __half*)grad, (__half*)W, scale, size, emb);
	return PZ_OK;
}

int pz_upsample_nearest_fwd(const void* x, void* y, int64_t planes, int D, int H, int W, int ds, int hs, int ws, void* stream)
{
	PZ_REQUIRE(planes >= 0 && D > 0 && H > 0 && W > 0 && ds > 0 && hs > 0 && ws > 0, "upsample: invalid shape");
	const int64_t total = planes * D * ds * H * hs * W * ws;
	if (total == 0) return PZ_OK;
	PZ_SHAPE_LAUNCH(upsample_nearest_fwd_kernel, total, (const float*)x, (float*)y, planes, D, H, W, ds, hs, ws);
	return PZ_OK;
}

int pz_upsample_nearest_bwd(const void* dy, void* dx, int64_t planes, int D, int H, int W, int ds, int hs, int ws, void* stream)
{
	PZ_REQUIRE(planes >= 0 && D > 0 && H > 0 && W > 0 && ds > 0 && hs > 0 && ws > 0, "upsample: invalid shape");
	const int64_t total = planes * D * H * W;
	if (total == 0) return PZ_OK;
	PZ_SHAPE_LAUNCH(upsample_nearest_bwd_kernel, total, (const float*)dy, (float*)dx, planes, D, H, W, ds, hs, ws);
	return PZ_OK;
}

int pz_upsample_linear_fwd(const void* x, void* y, int64_t planes, int D, int H, int W, int oD, int oH, int oW, float rd, float rh, float rw,
						   int three_d, void* stream)
{
	PZ_REQUIRE(planes >= 0 && D > 0 && H > 0 && W > 0 && oD > 0 && oH > 0 && oW > 0, "upsample: invalid shape");
	const int64_t total = planes * oD * oH * oW;
	if (total == 0) return PZ_OK;
	PZ_SHAPE_LAUNCH(upsample_linear_fwd_kernel, total, (const float*)x, (float*)y, planes, D, H, W, oD, oH, oW, rd, rh, rw, three_d);
	return PZ_OK;
}

int pz_upsample_linear_bwd(const void* dy, void* dx, int64_t planes, int D, int H, int W, int oD, int oH, int oW, float rd, float rh, float rw,
						   int three_d, void* stream)
{
	PZ_REQUIRE(planes >= 0 && D > 0 && H > 0 && W > 0 && oD > 0 && oH > 0 && oW > 0, "upsample: invalid shape");
	const int64_t total = planes * oD * oH * oW;
	if (total == 0) return PZ_OK;
	PZ_SHAPE_LAUNCH(upsample_linear_bwd_kernel, total, (const float*)dy, (float*)dx, planes, D, H, W, oD, oH, oW, rd, rh, rw, three_d);
	return PZ_OK;
}

int pz_divnorm_fwd(int dtype, const void* x, const void* means, void* y, int64_t planes, int64_t H, int64_t W, int n, float alpha, float beta, float K,
				   void* stream)
{
	DivGeo g{};
	int st = div_geo(g, dtype, planes, H, W, n, alpha, beta, K);
	if (st != PZ_OK) return st;
	const int64_t total = planes * H * W;
	if (total == 0) return PZ_OK;
	return dtype == PZ_F32 ? divnorm_launch<float>(0, x, means, nullptr, y, nullptr, nullptr, g, total, pz_stream(stream))
						   : divnorm_launch<__half>(0, x, means, nullptr, y, nullptr, nullptr, g, total, pz_stream(stream));
}

int pz_divnorm_bwd(int dtype, const void* x, const void* means, const void* grad, void* dx, void* dmeans, void* tmp, int64_t planes, int64_t H,
				   int64_t W, int n, float alpha, float beta, float K, void* stream)
{
	DivGeo g{};
	int st = div_geo(g, dtype, planes, H, W, n, alpha, beta, K);
	if (st != PZ_OK) return st;
	const int64_t total = planes * H * W;
	if (total == 0) return PZ_OK;
	return dtype == PZ_F32 ? divnorm_launch<float>(1, x, means, grad, dx, dmeans, (float*)tmp, g, total, pz_stream(stream))
						   : divnorm_launch<__half>(1, x, means, grad, dx, dmeans, (float*)tmp, g, total, pz_stream(stream));
}

int pz_spatialtf_fwd(int dtype, const void* data, const void* theta, void* grid, void* out, int64_t B, int64_t C, int H, int W, int oH, int oW,
					 void* stream)
{
	PZ_REQUIRE(dtype == PZ_F32 || dtype == PZ_F16, "spatialTf: unsupported dtype %d", dtype);
	PZ_REQUIRE(B >= 0 && C > 0 && C < (1ll << 31) && H > 0 && W > 0 && oH > 0 && oW > 0, "spatialTf: invalid shape");
	if (B == 0) return PZ_OK;
	const int64_t gtotal = B * oH * oW;
	if (dtype == PZ_F32) {
		PZ_SHAPE_LAUNCH(stf_grid_kernel<float>, gtotal, (const float*)theta, (float*)grid, B, oH, oW);
		PZ_SHAPE_LAUNCH(stf_sample_fwd_kernel<float>, gtotal * C, (const float*)data, (const float*)grid, (float*)out, B, (int)C, H, W, oH, oW);
	} else {
		PZ_SHAPE_LAUNCH(stf_grid_kernel<__half>, gtotal, (const __half*)theta, (__half*)grid, B, oH, oW);
		PZ_SHAPE_LAUNCH(stf_sample_fwd_kernel<__half>, gtotal * C, (const __half*)data, (const __half*)grid, (__half*)out, B, (int)C, H, W, oH, oW);
	}
	return PZ_OK;
}

int pz_spatialtf_bwd(int dtype, const void* grad, const void* data, const void* grid, void* dx, void* dtheta, void* dgrid, int64_t B, int64_t C,
					 int H, int W, int oH, int oW, void* stream)
{
	PZ_REQUIRE(dtype == PZ_F32 || dtype == PZ_F16, "spatialTf: unsupported dtype %d", dtype);
	PZ_REQUIRE(B >= 0 && C > 0 && C < (1ll << 31) && H > 0 && W > 0 && oH > 0 && oW > 0, "spatialTf: invalid shape");
	if (B == 0) return PZ_OK;
	const size_t es = dtype == PZ_F32 ? 4 : 2;
	PZ_CHECK_CUDA(cudaMemsetAsync(dx, 0, (size_t)(B * C) * H * W * es, pz_stream(stream)));
	const int64_t gtotal = B * oH * oW;
	if (dtype == PZ_F32) {
		PZ_SHAPE_LAUNCH(stf_sample_bwd_kernel<float>, gtotal, (const float*)grad, (const float*)data, (const float*)grid, (float*)dx, (float*)dgrid, B,
						(int)C, H, W, oH, oW);
		stf_dtheta_kernel<float><<<(unsigned)B, kThreads, 0, pz_stream(stream)>>>((const float*)dgrid, (float*)dtheta, oH, oW);
	} else {
		PZ_SHAPE_LAUNCH(stf_sample_bwd_kernel<__half>, gtotal, (const __half*)grad, (const __half*)data, (const __half*)grid, (__half*)dx,
						(__half*)dgrid, B, (int)C, H, W, oH, oW);
		stf_dtheta_kernel<__half><<<(unsigned)B, kThreads, 0, pz_stream(stream)>>>((const __half*)dgrid, (__half*)dtheta, oH, oW);
	}
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

int pz_ctc_loss(const void* y, const void* datalen, const void* labels, const void* offsets, void* alphas, void* nll, void* error, void* grad, int T,
				int B, int V, int max_label_len, int blank, void* stream)
{
	PZ_REQUIRE(T > 0 && B > 0 && V > 0 && max_label_len >= 0 && blank >= 0 && blank < V, "ctcLoss: invalid arguments");
	const int Smax = 2 * max_label_len + 1;
	const size_t smem_a = (size_t)Smax * sizeof(int), smem_b = (size_t)Smax * (2 * sizeof(int) + 2 * sizeof(float));
	PZ_REQUIRE(smem_b <= 200 * 1024, "ctcLoss: label sequences of %d symbols do not fit shared memory", max_label_len);
	cudaStream_t s = pz_stream(stream);
	if (smem_b > 48 * 1024) {
		PZ_CHECK_CUDA(cudaFuncSetAttribute(ctc_alpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_a));
		PZ_CHECK_CUDA(cudaFuncSetAttribute(ctc_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
	}
	ctc_alpha_kernel<<<(unsigned)B, kCtcThreads, smem_a, s>>>((const float*)y, (const int*)datalen, T, B, V, (const int*)labels, (const int*)offsets,
															  (float*)alphas, blank, (float*)nll, (float*)error);
	ctc_beta_kernel<<<(unsigned)B, kCtcThreads, smem_b, s>>>((const float*)y, (const int*)datalen, T, B, V, (const int*)labels, (const int*)offsets,
															 (const float*)alphas, blank, (const float*)nll, (float*)grad, Smax);
	pz_count_launch(2);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

}  // extern "C"
