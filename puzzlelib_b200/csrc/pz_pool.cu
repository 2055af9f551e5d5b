// pz_pool.cu -- 2-D pooling forward / backward (HBM-bandwidth bound).
//
// Two families, both working on [planes = N*C][H][W] row-major tensors:
//   * pz_pool2d_{fwd,bwd}: replaces cudnnPoolingForward / cudnnPoolingBackward as configured by the reference
//     (Cuda/Source/Libs/CuDnnPool.c:44-96,155-190: max / avg-with-pad / avg-no-pad / max-deterministic,
//     CUDNN_NOT_PROPAGATE_NAN).  Out size = (in + 2*pad - size)/stride + 1 (CuDnnPool.c:24-41).
//   * pz_maxpool2d_mask_{fwd,bwd}, pz_maxunpool2d_{fwd,bwd}: bit-exact restatement of the reference's own
//     JIT kernels (Cuda/Kernels/Pool.py:10-112): int32 in-plane argmax = first strict '>' winner of a
//     row-major window scan starting from -FLT_MAX, -1 for an empty window; backward adds the candidate
//     windows in (ph, pw) order.
//
// B200 design: one thread per in-plane output position with threadIdx.x running along the contiguous W axis, so each
// warp writes full 128-byte lines and its overlapping window reads are served from L1 (every input line is
// fetched from HBM once).  A thread decodes its (h, w) and window bounds ONCE and then walks over planes
// (blockIdx.y-strided), so the per-element work is loads + compares only -- no integer division in the loop.
// Tiny planes (global average pooling) use flat-indexed variants instead.  The backward kernels are GATHERS (one
// thread per dx element) -- no atomics, so results are deterministic run to run, which the reference's
// scatter-free mask kernel also guarantees.
#include "pz_common.h"

#include <cfloat>

namespace {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

struct PoolGeo {
	int H, W, OH, OW, fh, fw, sh, sw, ph, pw;
};

constexpr int kThreads = 256;

// ------------------------------------------------------------------------------------------ cuDNN-style forward
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads) pool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t total, PoolGeo g)
{
	for (int64_t index = (int64_t)blockIdx.x * kThreads + threadIdx.x; index < total; index += (int64_t)gridDim.x * kThreads) {
		const int ow = (int)(index % g.OW);
		const int64_t t = index / g.OW;
		const int oh = (int)(t % g.OH);
		const int64_t plane = t / g.OH;

		int h0 = oh * g.sh - g.ph, w0 = ow * g.sw - g.pw;
		const int h1 = min(h0 + g.fh, g.H), w1 = min(w0 + g.fw, g.W);
		h0 = max(h0, 0);
		w0 = max(w0, 0);
		const T* slice = x + plane * (int64_t)g.H * g.W;

		float r;
		if (MODE == PZ_POOL_MAX || MODE == PZ_POOL_MAX_DETERMINISM) {
			r = -FLT_MAX;
			for (int h = h0; h < h1; h++)
				for (int w = w0; w < w1; w++) r = fmaxf(r, to_f<T>(slice[h * g.W + w]));   // fmaxf drops NaNs (NOT_PROPAGATE_NAN)
		} else {
			float acc = 0.0f;
			for (int h = h0; h < h1; h++)
				for (int w = w0; w < w1; w++) acc += to_f<T>(slice[h * g.W + w]);
			const int cnt = MODE == PZ_POOL_AVG_WITH_PAD ? g.fh * g.fw : (h1 - h0) * (w1 - w0);
			r = acc / (float)cnt;
		}
		y[index] = from_f<T>(r);
	}
}

// ------------------------------------------------------------------------------------------ cuDNN-style backward
// max: dy of a window goes to the FIRST element (row-major scan) equal to the window maximum.  cuDNN's tie
// routing is not asserted by any reference test (SURVEY 8c "not pinned"); first-maximum is what both the
// reference's own mask kernel and the deterministic cuDNN mode do.
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads) pool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ y,
															const T* __restrict__ dy, T* __restrict__ dx, int64_t total, PoolGeo g)
{
	for (int64_t index = (int64_t)blockIdx.x * kThreads + threadIdx.x; index < total; index += (int64_t)gridDim.x * kThreads) {
		const int w = (int)(index % g.W);
		const int64_t t = index / g.W;
		const int h = (int)(t % g.H);
		const int64_t plane = t / g.H;

		const int oh0 = (h + g.ph < g.fh) ? 0 : (h + g.ph - g.fh) / g.sh + 1;
		const int oh1 = min((h + g.ph) / g.sh + 1, g.OH);
		const int ow0 = (w + g.pw < g.fw) ? 0 : (w + g.pw - g.fw) / g.sw + 1;
		const int ow1 = min((w + g.pw) / g.sw + 1, g.OW);

		const T* xs = x + plane * (int64_t)g.H * g.W;
		const int64_t ooff = plane * (int64_t)g.OH * g.OW;
		float grad = 0.0f;

		if (MODE == PZ_POOL_MAX || MODE == PZ_POOL_MAX_DETERMINISM) {
			const float xv = to_f<T>(xs[h * g.W + w]);
			for (int oh = oh0; oh < oh1; oh++)
				for (int ow = ow0; ow < ow1; ow++) {
					const float yv = to_f<T>(y[ooff + oh * g.OW + ow]);
					if (xv != yv) continue;
					// is (h, w) the first element of this window that equals the maximum?
					const int hs = max(oh * g.sh - g.ph, 0), ws = max(ow * g.sw - g.pw, 0);
					const int we = min(ow * g.sw - g.pw + g.fw, g.W);
					bool first = true;
					for (int hh = hs; hh <= h && first; hh++) {
						const int wend = hh == h ? w : we;
						for (int ww = ws; ww < wend; ww++)
							if (to_f<T>(xs[hh * g.W + ww]) == yv) { first = false; break; }
					}
					if (first) grad += to_f<T>(dy[ooff + oh * g.OW + ow]);
				}
		} else {
			for (int oh = oh0; oh < oh1; oh++)
				for (int ow = ow0; ow < ow1; ow++) {
					int cnt = g.fh * g.fw;
					if (MODE == PZ_POOL_AVG_NO_PAD) {
						const int hs = max(oh * g.sh - g.ph, 0), he = min(oh * g.sh - g.ph + g.fh, g.H);
						const int ws = max(ow * g.sw - g.pw, 0), we = min(ow * g.sw - g.pw + g.fw, g.W);
						cnt = (he - hs) * (we - ws);
					}
					grad += to_f<T>(dy[ooff + oh * g.OW + ow]) / (float)cnt;
				}
		}
		dx[index] = from_f<T>(grad);
	}
}

// ------------------------------------------------------------------------------------------ mask kernels (fp32, bit-exact)
__global__ void __launch_bounds__(kThreads) maxpool_mask_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
																	 int32_t* __restrict__ mask, int64_t total, PoolGeo g)
{
	for (int64_t index = (int64_t)blockIdx.x * kThreads + threadIdx.x; index < total; index += (int64_t)gridDim.x * kThreads) {
		const int ow = (int)(index % g.OW);
		const int64_t t = index / g.OW;
		const int oh = (int)(t % g.OH);
		const int64_t plane = t / g.OH;

		int h0 = oh * g.sh - g.ph, w0 = ow * g.sw - g.pw;
		const int h1 = min(h0 + g.fh, g.H), w1 = min(w0 + g.fw, g.W);
		h0 = max(h0, 0);
		w0 = max(w0, 0);
		const float* slice = x + plane * (int64_t)g.H * g.W;

		float maxval = -FLT_MAX;
		int maxidx = -1;
		for (int h = h0; h < h1; h++)
			for (int w = w0; w < w1; w++) {
				const float v = slice[h * g.W + w];
				if (v > maxval) { maxidx = h * g.W + w; maxval = v; }
			}
		y[index] = maxval;
		mask[index] = maxidx;
	}
}

__global__ void __launch_bounds__(kThreads) maxpool_mask_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ mask,
																	 float* __restrict__ dx, int64_t total, PoolGeo g)
{
	for (int64_t index = (int64_t)blockIdx.x * kThreads + threadIdx.x; index < total; index += (int64_t)gridDim.x * kThreads) {
		const int w = (int)(index % g.W);
		const int64_t t = index / g.W;
		const int h = (int)(t % g.H);
		const int64_t plane = t / g.H;

		const int oh0 = (h + g.ph < g.fh) ? 0 : (h + g.ph - g.fh) / g.sh + 1;
		const int oh1 = min((h + g.ph) / g.sh + 1, g.OH);
		const int ow0 = (w + g.pw < g.fw) ? 0 : (w + g.pw - g.fw) / g.sw + 1;
		const int ow1 = min((w + g.pw) / g.sw + 1, g.OW);

		const int64_t ooff = plane * (int64_t)g.OH * g.OW;
		const int me = h * g.W + w;
		float grad = 0.0f;
		for (int oh = oh0; oh < oh1; oh++)
			for (int ow = ow0; ow < ow1; ow++)
				if (mask[ooff + oh * g.OW + ow] == me) grad += dy[ooff + oh * g.OW + ow];
		dx[index] = grad;
	}
}

__global__ void __launch_bounds__(kThreads) maxunpool_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ mask,
																  float* __restrict__ y, int64_t total, int inHW, int outHW)
{
	for (int64_t index = (int64_t)blockIdx.x * kThreads + threadIdx.x; index < total; index += (int64_t)gridDim.x * kThreads) {
		const int64_t plane = index / inHW;
		const int m = mask[index];
		if (m >= 0 && m < outHW) y[plane * outHW + m] = x[index];
	}
}

__global__ void __launch_bounds__(kThreads) maxunpool_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ mask,
																  float* __restrict__ dx, int64_t total, int inHW, int outHW)
{
	for (int64_t index = (int64_t)blockIdx.x * kThreads + threadIdx.x; index < total; index += (int64_t)gridDim.x * kThreads) {
		const int64_t plane = index / inHW;
		const int m = mask[index];
		dx[index] = (m >= 0 && m < outHW) ? dy[plane * outHW + m] : 0.0f;
	}
}

// ------------------------------------------------------------------------------------------ plane-walking variants
// grid.x * blockDim.x covers the positions of one plane, grid.y strides over planes.
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads) pool_fwd_plane_kernel(const T* __restrict__ x, T* __restrict__ y, int planes, PoolGeo g)
{
	const int o = blockIdx.x * blockDim.x + threadIdx.x;
	const int OHW = g.OH * g.OW, HW = g.H * g.W;
	if (o >= OHW) return;
	const int oh = o / g.OW, ow = o - oh * g.OW;
	int h0 = oh * g.sh - g.ph, w0 = ow * g.sw - g.pw;
	const int h1 = min(h0 + g.fh, g.H), w1 = min(w0 + g.fw, g.W);
	h0 = max(h0, 0);
	w0 = max(w0, 0);
	for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {
		const T* slice = x + (int64_t)pl * HW;
		float r;
		if (MODE == PZ_POOL_MAX || MODE == PZ_POOL_MAX_DETERMINISM) {
			r = -FLT_MAX;
			for (int h = h0; h < h1; h++)
				for (int w = w0; w < w1; w++) r = fmaxf(r, to_f<T>(slice[h * g.W + w]));
		} else {
			float acc = 0.0f;
			for (int h = h0; h < h1; h++)
				for (int w = w0; w < w1; w++) acc += to_f<T>(slice[h * g.W + w]);
			r = MODE == PZ_POOL_AVG_WITH_PAD ? acc / (float)(g.fh * g.fw) : acc / (float)((h1 - h0) * (w1 - w0));
		}
		y[(int64_t)pl * OHW + o] = from_f<T>(r);
	}
}

template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads) pool_bwd_plane_kernel(const T* __restrict__ x, const T* __restrict__ y,
																  const T* __restrict__ dy, T* __restrict__ dx, int planes, PoolGeo g)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int OHW = g.OH * g.OW, HW = g.H * g.W;
	if (i >= HW) return;
	const int h = i / g.W, w = i - h * g.W;
	const int oh0 = (h + g.ph < g.fh) ? 0 : (h + g.ph - g.fh) / g.sh + 1;
	const int oh1 = min((h + g.ph) / g.sh + 1, g.OH);
	const int ow0 = (w + g.pw < g.fw) ? 0 : (w + g.pw - g.fw) / g.sw + 1;
	const int ow1 = min((w + g.pw) / g.sw + 1, g.OW);

	for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {
		const T* xs = x + (int64_t)pl * HW;
		const T* ys = y + (int64_t)pl * OHW;
		const T* gs = dy + (int64_t)pl * OHW;
		float grad = 0.0f;
		if (MODE == PZ_POOL_MAX || MODE == PZ_POOL_MAX_DETERMINISM) {
			const float xv = to_f<T>(xs[i]);
			for (int oh = oh0; oh < oh1; oh++)
				for (int ow = ow0; ow < ow1; ow++) {
					const float yv = to_f<T>(ys[oh * g.OW + ow]);
					if (xv != yv) continue;
					// is (h, w) the first element of this window that equals the maximum?
					const int hs = max(oh * g.sh - g.ph, 0), ws = max(ow * g.sw - g.pw, 0);
					const int we = min(ow * g.sw - g.pw + g.fw, g.W);
					bool first = true;
					for (int hh = hs; hh <= h && first; hh++) {
						const int wend = hh == h ? w : we;
						for (int ww = ws; ww < wend; ww++)
							if (to_f<T>(xs[hh * g.W + ww]) == yv) { first = false; break; }
					}
					if (first) grad += to_f<T>(gs[oh * g.OW + ow]);
				}
		} else {
			for (int oh = oh0; oh < oh1; oh++)
				for (int ow = ow0; ow < ow1; ow++) {
					int cnt = g.fh * g.fw;
					if (MODE == PZ_POOL_AVG_NO_PAD) {
						const int hs = max(oh * g.sh - g.ph, 0), he = min(oh * g.sh - g.ph + g.fh, g.H);
						const int ws = max(ow * g.sw - g.pw, 0), we = min(ow * g.sw - g.pw + g.fw, g.W);
						cnt = (he - hs) * (we - ws);
					}
					grad += to_f<T>(gs[oh * g.OW + ow]) / (float)cnt;
				}
		}
		dx[(int64_t)pl * HW + i] = from_f<T>(grad);
	}
}

__global__ void __launch_bounds__(kThreads) maxpool_mask_fwd_plane_kernel(const float* __restrict__ x, float* __restrict__ y,
																			   int32_t* __restrict__ mask, int planes, PoolGeo g)
{
	const int o = blockIdx.x * blockDim.x + threadIdx.x;
	const int OHW = g.OH * g.OW, HW = g.H * g.W;
	if (o >= OHW) return;
	const int oh = o / g.OW, ow = o - oh * g.OW;
	int h0 = oh * g.sh - g.ph, w0 = ow * g.sw - g.pw;
	const int h1 = min(h0 + g.fh, g.H), w1 = min(w0 + g.fw, g.W);
	h0 = max(h0, 0);
	w0 = max(w0, 0);
	for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {
		const float* slice = x + (int64_t)pl * HW;
		float maxval = -FLT_MAX;
		int maxidx = -1;
		for (int h = h0; h < h1; h++)
			for (int w = w0; w < w1; w++) {
				const float v = slice[h * g.W + w];
				if (v > maxval) { maxidx = h * g.W + w; maxval = v; }
			}
		y[(int64_t)pl * OHW + o] = maxval;
		mask[(int64_t)pl * OHW + o] = maxidx;
	}
}

__global__ void __launch_bounds__(kThreads) maxpool_mask_bwd_plane_kernel(const float* __restrict__ dy, const int32_t* __restrict__ mask,
																			   float* __restrict__ dx, int planes, PoolGeo g)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int OHW = g.OH * g.OW, HW = g.H * g.W;
	if (i >= HW) return;
	const int h = i / g.W, w = i - h * g.W;
	const int oh0 = (h + g.ph < g.fh) ? 0 : (h + g.ph - g.fh) / g.sh + 1;
	const int oh1 = min((h + g.ph) / g.sh + 1, g.OH);
	const int ow0 = (w + g.pw < g.fw) ? 0 : (w + g.pw - g.fw) / g.sw + 1;
	const int ow1 = min((w + g.pw) / g.sw + 1, g.OW);
	for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {
		const int32_t* ms = mask + (int64_t)pl * OHW;
		const float* gs = dy + (int64_t)pl * OHW;
		float grad = 0.0f;
		for (int oh = oh0; oh < oh1; oh++)
			for (int ow = ow0; ow < ow1; ow++)
				if (ms[oh * g.OW + ow] == i) grad += gs[oh * g.OW + ow];
		dx[(int64_t)pl * HW + i] = grad;
	}
}

// cuDNN-style max backward (x, dy -> dx) in two passes through a library-owned int32 winner map: pass 1 finds the first
// maximum of every window (exactly the forward scan), pass 2 gathers dy from the windows whose winner is this element.
// Uniform work per thread -- the direct "am I the first maximum of this window?" test re-scans the window divergently.
template <typename T>
__global__ void __launch_bounds__(kThreads) pool_argmax_plane_kernel(const T* __restrict__ x, int32_t* __restrict__ winner, int planes,
																		  PoolGeo g)
{
	const int o = blockIdx.x * blockDim.x + threadIdx.x;
	const int OHW = g.OH * g.OW, HW = g.H * g.W;
	if (o >= OHW) return;
	const int oh = o / g.OW, ow = o - oh * g.OW;
	int h0 = oh * g.sh - g.ph, w0 = ow * g.sw - g.pw;
	const int h1 = min(h0 + g.fh, g.H), w1 = min(w0 + g.fw, g.W);
	h0 = max(h0, 0);
	w0 = max(w0, 0);
	for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {
		const T* slice = x + (int64_t)pl * HW;
		float maxval = -FLT_MAX;
		int maxidx = h0 * g.W + w0;          // a window of -FLT_MAX / NaN still routes its gradient somewhere (cuDNN does)
		for (int h = h0; h < h1; h++)
			for (int w = w0; w < w1; w++) {
				const float v = to_f<T>(slice[h * g.W + w]);
				if (v > maxval) { maxidx = h * g.W + w; maxval = v; }
			}
		winner[(int64_t)pl * OHW + o] = maxidx;
	}
}

template <typename T>
__global__ void __launch_bounds__(kThreads) pool_gather_plane_kernel(const T* __restrict__ dy, const int32_t* __restrict__ winner,
																		  T* __restrict__ dx, int planes, PoolGeo g)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int OHW = g.OH * g.OW, HW = g.H * g.W;
	if (i >= HW) return;
	const int h = i / g.W, w = i - h * g.W;
	const int oh0 = (h + g.ph < g.fh) ? 0 : (h + g.ph - g.fh) / g.sh + 1;
	const int oh1 = min((h + g.ph) / g.sh + 1, g.OH);
	const int ow0 = (w + g.pw < g.fw) ? 0 : (w + g.pw - g.fw) / g.sw + 1;
	const int ow1 = min((w + g.pw) / g.sw + 1, g.OW);
	// U planes per round: the winner lookups of the U planes are independent loads (one plane at a time is a chain of two
	// dependent loads per window and runs at a sixth of the HBM rate)
	constexpr int U = 8;
	for (int pl = blockIdx.y; pl < planes; pl += gridDim.y * U) {
		float grad[U];
		#pragma unroll
		for (int u = 0; u < U; u++) grad[u] = 0.0f;
		for (int oh = oh0; oh < oh1; oh++)
			for (int ow = ow0; ow < ow1; ow++) {
				const int o = oh * g.OW + ow;
				int32_t m[U];
				#pragma unroll
				for (int u = 0; u < U; u++) {
					const int q = pl + u * (int)gridDim.y;
					m[u] = q < planes ? winner[(int64_t)q * OHW + o] : -1;
				}
				#pragma unroll
				for (int u = 0; u < U; u++)
					if (m[u] == i) grad[u] += to_f<T>(dy[(int64_t)(pl + u * (int)gridDim.y) * OHW + o]);
			}
		#pragma unroll
		for (int u = 0; u < U; u++) {
			const int q = pl + u * (int)gridDim.y;
			if (q < planes) dx[(int64_t)q * HW + i] = from_f<T>(grad[u]);
		}
	}
}

// launch geometry of the plane-walking kernels: `elems` positions per plane
struct PlaneGrid { dim3 grid; unsigned threads; };
PlaneGrid plane_grid(int elems, int64_t planes)
{
	PlaneGrid pg;
	pg.threads = elems >= kThreads ? kThreads : (unsigned)((elems + 31) & ~31);
	const unsigned gx = (unsigned)pz_cdiv(elems, pg.threads);
	int64_t gy = pz_cdiv((int64_t)pz_num_sms() * 16, gx);
	if (gy > planes) gy = planes;
	if (gy > 65535) gy = 65535;
	if (gy < 1) gy = 1;
	pg.grid = dim3(gx, (unsigned)gy);
	return pg;
}
constexpr int kMinPlaneElems = 64;     // smaller planes: flat-indexed kernels (e.g. global average pooling)

unsigned grid_for(int64_t total)
{
	int64_t blocks = pz_cdiv(total, kThreads);
	const int64_t cap = (int64_t)pz_num_sms() * 32;
	if (blocks > cap) blocks = cap;
	return (unsigned)(blocks < 1 ? 1 : blocks);
}

int check_geo(int64_t planes, const PoolGeo& g)
{
	PZ_REQUIRE(planes > 0 && g.H > 0 && g.W > 0, "pool2d: empty input");
	PZ_REQUIRE(g.fh > 0 && g.fw > 0 && g.sh > 0 && g.sw > 0 && g.ph >= 0 && g.pw >= 0, "pool2d: invalid window parameters");
	PZ_REQUIRE(g.H + 2 * g.ph >= g.fh && g.W + 2 * g.pw >= g.fw, "pool2d: invalid input map size");
	PZ_REQUIRE(g.OH == (g.H + 2 * g.ph - g.fh) / g.sh + 1 && g.OW == (g.W + 2 * g.pw - g.fw) / g.sw + 1,
			   "pool2d: output size %dx%d inconsistent with input %dx%d", g.OH, g.OW, g.H, g.W);
	PZ_REQUIRE((int64_t)g.H * g.W < (1ll << 31), "pool2d: plane too large");
	return PZ_OK;
}

template <typename T>
int fwd_dispatch(int mode, const void* x, void* y, int64_t planes, const PoolGeo& g, void* stream)
{
	const int64_t total = planes * g.OH * g.OW;
	const unsigned grid = grid_for(total);
	cudaStream_t s = pz_stream(stream);
	PzProfScope prof(PZ_PROF_POOL, s, 0.0, (double)sizeof(T) * planes * ((double)g.H * g.W + (double)g.OH * g.OW));
	if (g.OH * g.OW >= kMinPlaneElems && planes < (1ll << 31)) {
		const PlaneGrid pg = plane_grid(g.OH * g.OW, planes);
		switch (mode) {
			case PZ_POOL_MAX:
			case PZ_POOL_MAX_DETERMINISM:
				pool_fwd_plane_kernel<T, PZ_POOL_MAX><<<pg.grid, pg.threads, 0, s>>>((const T*)x, (T*)y, (int)planes, g); break;
			case PZ_POOL_AVG_WITH_PAD:
				pool_fwd_plane_kernel<T, PZ_POOL_AVG_WITH_PAD><<<pg.grid, pg.threads, 0, s>>>((const T*)x, (T*)y, (int)planes, g); break;
			case PZ_POOL_AVG_NO_PAD:
				pool_fwd_plane_kernel<T, PZ_POOL_AVG_NO_PAD><<<pg.grid, pg.threads, 0, s>>>((const T*)x, (T*)y, (int)planes, g); break;
			default:
				pz_set_error(PZ_ERR_VALUE, "pool2d: unknown mode %d", mode);
				return PZ_ERR_VALUE;
		}
		pz_count_launch(1);
		PZ_LAUNCH_CHECK();
		return PZ_OK;
	}
	switch (mode) {
		case PZ_POOL_MAX:
		case PZ_POOL_MAX_DETERMINISM:
			pool_fwd_kernel<T, PZ_POOL_MAX><<<grid, kThreads, 0, s>>>((const T*)x, (T*)y, total, g); break;
		case PZ_POOL_AVG_WITH_PAD:
			pool_fwd_kernel<T, PZ_POOL_AVG_WITH_PAD><<<grid, kThreads, 0, s>>>((const T*)x, (T*)y, total, g); break;
		case PZ_POOL_AVG_NO_PAD:
			pool_fwd_kernel<T, PZ_POOL_AVG_NO_PAD><<<grid, kThreads, 0, s>>>((const T*)x, (T*)y, total, g); break;
		default:
			pz_set_error(PZ_ERR_VALUE, "pool2d: unknown mode %d", mode);
			return PZ_ERR_VALUE;
	}
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

template <typename T>
int bwd_dispatch(int mode, const void* x, const void* y, const void* dy, void* dx, int64_t planes, const PoolGeo& g, void* stream)
{
	const int64_t total = planes * g.H * g.W;
	const unsigned grid = grid_for(total);
	cudaStream_t s = pz_stream(stream);
	PzProfScope prof(PZ_PROF_POOL, s, 0.0, (double)sizeof(T) * planes * (2.0 * g.H * g.W + 2.0 * g.OH * g.OW));
	if (g.H * g.W >= kMinPlaneElems && planes < (1ll << 31)) {
		const PlaneGrid pg = plane_grid(g.H * g.W, planes);
		switch (mode) {
			case PZ_POOL_MAX:
			case PZ_POOL_MAX_DETERMINISM: {
				int32_t* winner = g.OH * g.OW >= 32 ? (int32_t*)pz_scratch((size_t)planes * g.OH * g.OW * sizeof(int32_t)) : nullptr;
				if (winner) {
					const PlaneGrid po = plane_grid(g.OH * g.OW, planes);
					pool_argmax_plane_kernel<T><<<po.grid, po.threads, 0, s>>>((const T*)x, winner, (int)planes, g);
					pool_gather_plane_kernel<T><<<pg.grid, pg.threads, 0, s>>>((const T*)dy, winner, (T*)dx, (int)planes, g);
					pz_count_launch(1);
				} else
					pool_bwd_plane_kernel<T, PZ_POOL_MAX><<<pg.grid, pg.threads, 0, s>>>((const T*)x, (const T*)y, (const T*)dy, (T*)dx, (int)planes, g);
				break;
			}
			case PZ_POOL_AVG_WITH_PAD:
				pool_bwd_plane_kernel<T, PZ_POOL_AVG_WITH_PAD><<<pg.grid, pg.threads, 0, s>>>((const T*)x, (const T*)y, (const T*)dy, (T*)dx, (int)planes, g); break;
			case PZ_POOL_AVG_NO_PAD:
				pool_bwd_plane_kernel<T, PZ_POOL_AVG_NO_PAD><<<pg.grid, pg.threads, 0, s>>>((const T*)x, (const T*)y, (const T*)dy, (T*)dx, (int)planes, g); break;
			default:
				pz_set_error(PZ_ERR_VALUE, "pool2d: unknown mode %d", mode);
				return PZ_ERR_VALUE;
		}
		pz_count_launch(1);
		PZ_LAUNCH_CHECK();
		return PZ_OK;
	}
	switch (mode) {
		case PZ_POOL_MAX:
		case PZ_POOL_MAX_DETERMINISM:
			pool_bwd_kernel<T, PZ_POOL_MAX><<<grid, kThreads, 0, s>>>((const T*)x, (const T*)y, (const T*)dy, (T*)dx, total, g); break;
		case PZ_POOL_AVG_WITH_PAD:
			pool_bwd_kernel<T, PZ_POOL_AVG_WITH_PAD><<<grid, kThreads, 0, s>>>((const T*)x, (const T*)y, (const T*)dy, (T*)dx, total, g); break;
		case PZ_POOL_AVG_NO_PAD:
			pool_bwd_kernel<T, PZ_POOL_AVG_NO_PAD><<<grid, kThreads, 0, s>>>((const T*)x, (const T*)y, (const T*)dy, (T*)dx, total, g); break;
		default:
			pz_set_error(PZ_ERR_VALUE, "pool2d: unknown mode %d", mode);
			return PZ_ERR_VALUE;
	}
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

#define PZ_DISPATCH_FLOAT(dtype, ...)                                                    \
	switch (dtype) {                                                                     \
		case PZ_F32: { using T = float; return __VA_ARGS__; }                            \
		case PZ_F16: { using T = __half; return __VA_ARGS__; }                           \
		case PZ_BF16: { using T = __nv_bfloat16; return __VA_ARGS__; }                   \
		default: pz_set_error(PZ_ERR_UNSUPPORTED, "unsupported dtype %d", (int)(dtype)); \
				 return PZ_ERR_UNSUPPORTED;                                              \
	}

}  // namespace

extern "C" {

int pz_pool2d_fwd(int dtype, int mode, const void* x, void* y, int64_t planes, int H, int W, int OH, int OW, int fh, int fw,
				  int sh, int sw, int ph, int pw, void* stream)
{
	PoolGeo g{H, W, OH, OW, fh, fw, sh, sw, ph, pw};
	int st = check_geo(planes, g);
	if (st != PZ_OK) return st;
	PZ_DISPATCH_FLOAT(dtype, fwd_dispatch<T>(mode, x, y, planes, g, stream));
}

int pz_pool2d_bwd(int dtype, int mode, const void* x, const void* y, const void* dy, void* dx, int64_t planes, int H, int W,
				  int OH, int OW, int fh, int fw, int sh, int sw, int ph, int pw, void* stream)
{
	PoolGeo g{H, W, OH, OW, fh, fw, sh, sw, ph, pw};
	int st = check_geo(planes, g);
	if (st != PZ_OK) return st;
	PZ_DISPATCH_FLOAT(dtype, bwd_dispatch<T>(mode, x, y, dy, dx, planes, g, stream));
}

int pz_maxpool2d_mask_fwd(const float* x, float* y, int32_t* mask, int64_t planes, int H, int W, int OH, int OW, int fh, int fw,
						  int sh, int sw, int ph, int pw, void* stream)
{
	PoolGeo g{H, W, OH, OW, fh, fw, sh, sw, ph, pw};
	int st = check_geo(planes, g);
	if (st != PZ_OK) return st;
	const int64_t total = planes * OH * OW;
	PzProfScope prof(PZ_PROF_POOL, pz_stream(stream), 0.0, 4.0 * planes * ((double)H * W + 2.0 * OH * OW));
	if (OH * OW >= kMinPlaneElems && planes < (1ll << 31)) {
		const PlaneGrid pg = plane_grid(OH * OW, planes);
		maxpool_mask_fwd_plane_kernel<<<pg.grid, pg.threads, 0, pz_stream(stream)>>>(x, y, mask, (int)planes, g);
	} else
		maxpool_mask_fwd_kernel<<<grid_for(total), kThreads, 0, pz_stream(stream)>>>(x, y, mask, total, g);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

int pz_maxpool2d_mask_bwd(const float* dy, const int32_t* mask, float* dx, int64_t planes, int H, int W, int OH, int OW, int fh,
						  int fw, int sh, int sw, int ph, int pw, void* stream)
{
	PoolGeo g{H, W, OH, OW, fh, fw, sh, sw, ph, pw};
	int st = check_geo(planes, g);
	if (st != PZ_OK) return st;
	const int64_t total = planes * H * W;
	PzProfScope prof(PZ_PROF_POOL, pz_stream(stream), 0.0, 4.0 * planes * ((double)H * W + 2.0 * OH * OW));
	if (H * W >= kMinPlaneElems && planes < (1ll << 31)) {
		const PlaneGrid pg = plane_grid(H * W, planes);
		maxpool_mask_bwd_plane_kernel<<<pg.grid, pg.threads, 0, pz_stream(stream)>>>(dy, mask, dx, (int)planes, g);
	} else
		maxpool_mask_bwd_kernel<<<grid_for(total), kThreads, 0, pz_stream(stream)>>>(dy, mask, dx, total, g);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

int pz_maxunpool2d_fwd(const float* x, const int32_t* mask, float* y, int64_t planes, int inHW, int outHW, void* stream)
{
	PZ_REQUIRE(planes > 0 && inHW > 0 && outHW > 0, "maxunpool2d: empty tensor");
	const int64_t total = planes * inHW;
	maxunpool_fwd_kernel<<<grid_for(total), kThreads, 0, pz_stream(stream)>>>(x, mask, y, total, inHW, outHW);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

int pz_maxunpool2d_bwd(const float* dy, const int32_t* mask, float* dx, int64_t planes, int inHW, int outHW, void* stream)
{
	PZ_REQUIRE(planes > 0 && inHW > 0 && outHW > 0, "maxunpool2d: empty tensor");
	const int64_t total = planes * inHW;
	maxunpool_bwd_kernel<<<grid_for(total), kThreads, 0, pz_stream(stream)>>>(dy, mask, dx, total, inHW, outHW);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

}  // extern "C"
