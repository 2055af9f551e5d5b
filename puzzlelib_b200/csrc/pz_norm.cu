// pz_norm.cu -- batch normalisation forward (train / inference) and backward, HBM-bandwidth bound.
//
// Replaces cudnnBatchNormalizationForwardTraining / ForwardInference / Backward as called from
// CuDnn_Context_batchNormNd / _batchNormNdBackward (reference Cuda/Source/Libs/CuDnnNorm.c:31-71,158-194).
//
// B200 design.  An NCHW tensor is the matrix [N rows][C*S columns]: row n is contiguous, column j belongs to channel
// j / S.  A thread owns VEC consecutive columns (one 128-bit vector) and walks over rows, so
//   * every load / store is a fully coalesced, 16-byte aligned vector whatever the plane size (7x7 planes are as
//     efficient as 112x112 ones -- no per-plane head / tail handling),
//   * the channels of a thread's columns never change: per-channel coefficients live in registers for the whole walk,
//   * the row loop is unrolled, giving 8 independent 128-bit loads in flight per thread.
// Two kernels per pass: `stats` (column sums over its rows -> shared-memory per-channel bins -> fp32 red.add into a
// [C][2] buffer) and `apply` (normalise / input gradient, pure streaming).  Both are persistent: the work items
// (column block x group of 8 rows) are dealt out evenly to SMs x 8 CTAs, consecutive items of a CTA share the column
// block, so registers carry the partial sums / coefficients across them (measured on B200: splitting a tensor into
// L2-sized channel slabs to save the second DRAM read costs more in launch ramps than it saves; PZ_BN_SLAB_MB).
// Statistics are accumulated around a per-channel pivot (the channel's first element) so that E[d^2] - E[d]^2 does
// not cancel when |mean| >> std.
#include "pz_common.h"
#pragma nv_diag_suppress 128      // "loop is not reachable": the compile-time BULK branches of the cluster kernels return early

#include <cstdlib>
#include <initializer_list>
#include <map>
#include <utility>
#include <mutex>

namespace {

constexpr int kThreads = 256;
constexpr int kRowUnroll = 8;
// tensor bytes per slab that the apply kernel re-reads from L2 (PZ_BN_SLAB_MB overrides, for tuning)
double slab_bytes()
{
	static const double v = [] { const char* e = getenv("PZ_BN_SLAB_MB"); return (e ? atof(e) : 1e9) * 1024 * 1024; }();
	return v;
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T, int VEC> struct alignas(sizeof(T) * VEC) Pack { T v[VEC]; };

struct FastDiv32 { uint32_t d, m, sh; };     // d == 1: identity

inline FastDiv32 make_fastdiv32(uint32_t d)
{
	FastDiv32 f{d, 0, 0};
	if (d <= 1) return f;
	uint32_t s = 0;
	while ((1ull << s) < d) s++;
	const uint32_t p = 31 + s;
	f.m = (uint32_t)(((1ull << p) + d - 1) / d);
	f.sh = p - 32;
	return f;
}
// exact for n < 2^31
__device__ __forceinline__ uint32_t fdiv32(uint32_t n, const FastDiv32& f) { return f.d == 1 ? n : (__umulhi(n, f.m) >> f.sh); }

__device__ __forceinline__ float warp_sum(float v)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// geometry of one launch: columns [col0, col0 + ncols) of the [N][C*S] matrix, rows split over gridDim.y
struct Slab {
	int64_t row_stride;      // C * S elements
	int N, rows_per_block;
	uint32_t col0, ncols;    // columns of this slab (col0 is a multiple of VEC and of S)
	FastDiv32 sdiv;          // S
	uint32_t S;
	int row_groups, items, items_per_cta;   // persistent schedule: item = col_block * row_groups + row_group
};

// the CTA's contiguous range of work items
struct ItemRange { int begin, end; };
__device__ __forceinline__ ItemRange my_items(const Slab& g)
{
	ItemRange r;
	r.begin = blockIdx.x * g.items_per_cta;
	r.end = min(g.items, r.begin + g.items_per_cta);
	return r;
}

// adds the VEC per-column partial sums (a[e], b[e]) of a thread into the per-channel bins of the CTA and then into the global
// [C][2] buffer.  ch[e] is non-decreasing in e and across the threads of the CTA.
template <int VEC>
__device__ __forceinline__ void reduce_to_channels(const float (&a)[VEC], const float (&b)[VEC], const int (&ch)[VEC], bool active,
												   float* bins /* smem [nbins][2] */, int nbins, int ch_base, float* sums)
{
	for (int i = threadIdx.x; i < nbins * 2; i += kThreads) bins[i] = 0.0f;
	__syncthreads();
	// fast path: the whole warp works on one channel -> shuffle reduction, one shared-memory atomic per warp
	const int first = __shfl_sync(0xffffffffu, active ? ch[0] : -1, 0);
	const bool uniform = __all_sync(0xffffffffu, active && ch[0] == first && ch[VEC - 1] == first);
	if (uniform) {
		float sa = 0.0f, sb = 0.0f;
		#pragma unroll
		for (int e = 0; e < VEC; e++) { sa += a[e]; sb += b[e]; }
		sa = warp_sum(sa);
		sb = warp_sum(sb);
		if ((threadIdx.x & 31) == 0) {
			atomicAdd(&bins[(first - ch_base) * 2], sa);
			atomicAdd(&bins[(first - ch_base) * 2 + 1], sb);
		}
	} else if (active) {
		float sa = a[0], sb = b[0];
		#pragma unroll
		for (int e = 1; e < VEC; e++) {
			if (ch[e] != ch[e - 1]) {
				atomicAdd(&bins[(ch[e - 1] - ch_base) * 2], sa);
				atomicAdd(&bins[(ch[e - 1] - ch_base) * 2 + 1], sb);
				sa = 0.0f;
				sb = 0.0f;
			}
			sa += a[e];
			sb += b[e];
		}
		atomicAdd(&bins[(ch[VEC - 1] - ch_base) * 2], sa);
		atomicAdd(&bins[(ch[VEC - 1] - ch_base) * 2 + 1], sb);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < nbins * 2; i += kThreads) {
		const float v = bins[i];
		if (v != 0.0f) atomicAdd(&sums[(size_t)ch_base * 2 + i], v);
	}
}

// number of per-channel bins a CTA of kThreads * VEC consecutive columns can touch
__host__ __device__ inline int bins_per_block(int vec, uint32_t S) { return (int)((uint32_t)(kThreads * vec - 1) / S) + 2; }

// ------------------------------------------------------------------------------------------ forward statistics
// sums[c] = { sum(x - pivot_c), sum((x - pivot_c)^2) } over the rows of this block
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads) bn_stats_kernel(const T* __restrict__ x, Slab g, float* __restrict__ sums, float* __restrict__ pivots)
{
	extern __shared__ float bins[];
	const ItemRange ir = my_items(g);
	int cur = -1, ch_base = 0;
	uint32_t j = 0;
	bool active = false;
	int ch[VEC];
	float pivot[VEC], s1[VEC], s2[VEC];
	for (int it = ir.begin; it < ir.end; it++) {
		const int cb = it / g.row_groups, rg = it - cb * g.row_groups;
		if (cb != cur) {
			if (cur >= 0) reduce_to_channels<VEC>(s1, s2, ch, active, bins, bins_per_block(VEC, g.S), ch_base, sums);
			cur = cb;
			j = g.col0 + ((uint32_t)cb * kThreads + threadIdx.x) * VEC;
			active = j < g.col0 + g.ncols;
			ch_base = (int)fdiv32(g.col0 + (uint32_t)cb * kThreads * VEC, g.sdiv);
			#pragma unroll
			for (int e = 0; e < VEC; e++) {
				ch[e] = (int)fdiv32(active ? j + e : g.col0, g.sdiv);
				pivot[e] = to_f<T>(x[(size_t)ch[e] * g.S]);
				s1[e] = 0.0f;
				s2[e] = 0.0f;
				// the apply kernel takes the pivot from here, not from x: with y == x another CTA may have normalised that element already
				if (active && rg == 0 && j + e == (uint32_t)ch[e] * g.S) pivots[ch[e]] = pivot[e];
			}
		}
		if (!active) continue;
		const int r0 = rg * kRowUnroll, r1 = min(g.N, r0 + kRowUnroll);
		const T* p = x + (size_t)r0 * g.row_stride + j;
		if (r1 - r0 == kRowUnroll) {
			Pack<T, VEC> v[kRowUnroll];
			#pragma unroll
			for (int u = 0; u < kRowUnroll; u++) v[u] = *reinterpret_cast<const Pack<T, VEC>*>(p + (size_t)u * g.row_stride);
			#pragma unroll
			for (int u = 0; u < kRowUnroll; u++) {
				#pragma unroll
				for (int e = 0; e < VEC; e++) { const float d = to_f<T>(v[u].v[e]) - pivot[e]; s1[e] += d; s2[e] = fmaf(d, d, s2[e]); }
			}
		} else {
			for (int r = r0; r < r1; r++) {
				const Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(p);
				#pragma unroll
				for (int e = 0; e < VEC; e++) { const float d = to_f<T>(v.v[e]) - pivot[e]; s1[e] += d; s2[e] = fmaf(d, d, s2[e]); }
				p += g.row_stride;
			}
		}
	}
	if (cur >= 0) reduce_to_channels<VEC>(s1, s2, ch, active, bins, bins_per_block(VEC, g.S), ch_base, sums);
}

// ------------------------------------------------------------------------------------------ forward apply
// TRAIN: coefficients from the accumulated sums (+ writes the saved / running statistics once per channel);
// otherwise from the given mean / var (inference).
template <typename T, int VEC, bool TRAIN>
__global__ void __launch_bounds__(kThreads) bn_apply_kernel(const T* __restrict__ x, T* __restrict__ y, Slab g,
															const float* __restrict__ sums, const float* __restrict__ pivots,
															const float* __restrict__ scale,
															const float* __restrict__ bias, float* mean_io, float* var_io, float* save_mean,
															float* save_invvar, float eps, float factor, float count, float* zero_ptr,
															int zero_n)
{
	// clear the accumulators the NEXT batch-norm pass will use (two buffers alternate, so no memset launch is needed)
	for (int i = blockIdx.x * kThreads + threadIdx.x; i < zero_n; i += gridDim.x * kThreads) zero_ptr[i] = 0.0f;
	const ItemRange ir = my_items(g);
	int cur = -1;
	uint32_t j = 0;
	bool active = false;
	float a[VEC], b[VEC], m[VEC];      // y = (x - m) * a + b
	for (int it = ir.begin; it < ir.end; it++) {
		const int cb = it / g.row_groups, rg = it - cb * g.row_groups;
		if (cb != cur) {
			cur = cb;
			j = g.col0 + ((uint32_t)cb * kThreads + threadIdx.x) * VEC;
			active = j < g.col0 + g.ncols;
			if (active) {
				#pragma unroll
				for (int e = 0; e < VEC; e++) {
					const int c = (int)fdiv32(j + e, g.sdiv);
					float mean, invstd;
					if (TRAIN) {
						const float pivot = pivots[c];
						const float dmean = sums[2 * c] / count;
						mean = pivot + dmean;
						const float var = fmaxf(sums[2 * c + 1] / count - dmean * dmean, 0.0f);      // biased, used for normalisation
						invstd = 1.0f / sqrtf(var + eps);
						if (rg == 0 && j + e == (uint32_t)c * g.S) {
							save_mean[c] = mean;
							save_invvar[c] = invstd;
							// cuDNN keeps the UNBIASED variance in the running estimate
							const float uvar = count > 1.0f ? var * (count / (count - 1.0f)) : var;
							mean_io[c] = (1.0f - factor) * mean_io[c] + factor * mean;
							var_io[c] = (1.0f - factor) * var_io[c] + factor * uvar;
						}
					} else {
						mean = mean_io[c];
						invstd = 1.0f / sqrtf(var_io[c] + eps);
					}
					a[e] = scale[c] * invstd;
					b[e] = bias[c];
					m[e] = mean;
				}
			}
		}
		if (!active) continue;
		const int r0 = rg * kRowUnroll, r1 = min(g.N, r0 + kRowUnroll);
		const T* p = x + (size_t)r0 * g.row_stride + j;
		T* q = y + (size_t)r0 * g.row_stride + j;
		if (r1 - r0 == kRowUnroll) {
			Pack<T, VEC> v[kRowUnroll];
			#pragma unroll
			for (int u = 0; u < kRowUnroll; u++) v[u] = *reinterpret_cast<const Pack<T, VEC>*>(p + (size_t)u * g.row_stride);
			#pragma unroll
			for (int u = 0; u < kRowUnroll; u++) {
				#pragma unroll
				for (int e = 0; e < VEC; e++) v[u].v[e] = from_f<T>(fmaf(to_f<T>(v[u].v[e]) - m[e], a[e], b[e]));
				*reinterpret_cast<Pack<T, VEC>*>(q + (size_t)u * g.row_stride) = v[u];
			}
		} else {
			for (int r = r0; r < r1; r++) {
				Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(p);
				#pragma unroll
				for (int e = 0; e < VEC; e++) v.v[e] = from_f<T>(fmaf(to_f<T>(v.v[e]) - m[e], a[e], b[e]));
				*reinterpret_cast<Pack<T, VEC>*>(q) = v;
				p += g.row_stride;
				q += g.row_stride;
			}
		}
	}
}

// ------------------------------------------------------------------------------------------ backward statistics
// sums[c] = { sum(dy), sum(dy * (x - mean_c)) }
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads) bn_bwd_stats_kernel(const T* __restrict__ x, const T* __restrict__ dy, Slab g,
																const float* __restrict__ save_mean, float* __restrict__ sums)
{
	extern __shared__ float bins[];
	constexpr int UNR = kRowUnroll / 2;          // two tensors are read per row: 2 x 4 loads in flight
	const ItemRange ir = my_items(g);
	int cur = -1, ch_base = 0;
	uint32_t j = 0;
	bool active = false;
	int ch[VEC];
	float mean[VEC], s1[VEC], s2[VEC];
	for (int it = ir.begin; it < ir.end; it++) {
		const int cb = it / g.row_groups, rg = it - cb * g.row_groups;
		if (cb != cur) {
			if (cur >= 0) reduce_to_channels<VEC>(s1, s2, ch, active, bins, bins_per_block(VEC, g.S), ch_base, sums);
			cur = cb;
			j = g.col0 + ((uint32_t)cb * kThreads + threadIdx.x) * VEC;
			active = j < g.col0 + g.ncols;
			ch_base = (int)fdiv32(g.col0 + (uint32_t)cb * kThreads * VEC, g.sdiv);
			#pragma unroll
			for (int e = 0; e < VEC; e++) {
				ch[e] = (int)fdiv32(active ? j + e : g.col0, g.sdiv);
				mean[e] = save_mean[ch[e]];
				s1[e] = 0.0f;
				s2[e] = 0.0f;
			}
		}
		if (!active) continue;
		const int r0 = rg * kRowUnroll, r1 = min(g.N, r0 + kRowUnroll);
		const size_t off = (size_t)r0 * g.row_stride + j;
		const T* p = x + off;
		const T* q = dy + off;
		int r = r0;
		for (; r + UNR <= r1; r += UNR) {
			Pack<T, VEC> v[UNR], w[UNR];
			#pragma unroll
			for (int u = 0; u < UNR; u++) {
				v[u] = *reinterpret_cast<const Pack<T, VEC>*>(p + (size_t)u * g.row_stride);
				w[u] = *reinterpret_cast<const Pack<T, VEC>*>(q + (size_t)u * g.row_stride);
			}
			#pragma unroll
			for (int u = 0; u < UNR; u++) {
				#pragma unroll
				for (int e = 0; e < VEC; e++) {
					const float gv = to_f<T>(w[u].v[e]);
					s1[e] += gv;
					s2[e] = fmaf(gv, to_f<T>(v[u].v[e]) - mean[e], s2[e]);
				}
			}
			p += (size_t)UNR * g.row_stride;
			q += (size_t)UNR * g.row_stride;
		}
		for (; r < r1; r++) {
			const Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(p);
			const Pack<T, VEC> w = *reinterpret_cast<const Pack<T, VEC>*>(q);
			#pragma unroll
			for (int e = 0; e < VEC; e++) {
				const float gv = to_f<T>(w.v[e]);
				s1[e] += gv;
				s2[e] = fmaf(gv, to_f<T>(v.v[e]) - mean[e], s2[e]);
			}
			p += g.row_stride;
			q += g.row_stride;
		}
	}
	if (cur >= 0) reduce_to_channels<VEC>(s1, s2, ch, active, bins, bins_per_block(VEC, g.S), ch_base, sums);
}

// ------------------------------------------------------------------------------------------ backward apply
// Parameter-gradient accumulation riding on the backward pass (pz_bn_bwd_acc): acc = alpha * grad + beta * acc per channel, with
// the arithmetic of the elementwise kernel it replaces (Axpby in pz_elementwise.cu, built with --use_fast_math: y * beta rounded
// and flushed, then one fused multiply-add, flushed) -- the results are the same bits as two `addKer` launches after the pass.
struct BnAcc {
	float* scale_acc;
	float* bias_acc;
	float scale_alpha, scale_beta, bias_alpha, bias_beta;
};
__device__ __forceinline__ float axpby_ftz(float x, float alpha, float y, float beta)
{
	float t, r;
	asm("mul.rn.ftz.f32 %0, %1, %2;" : "=f"(t) : "f"(y), "f"(beta));
	asm("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(x), "f"(alpha), "f"(t));
	return r;
}
__device__ __forceinline__ void bn_param_grads(const BnAcc& a, int c, float dsc, float dbi, float* dscale, float* dbias)
{
	dscale[c] = dsc;
	dbias[c] = dbi;
	if (a.scale_acc) a.scale_acc[c] = axpby_ftz(dsc, a.scale_alpha, a.scale_acc[c], a.scale_beta);
	if (a.bias_acc) a.bias_acc[c] = axpby_ftz(dbi, a.bias_alpha, a.bias_acc[c], a.bias_beta);
}

// dx = c1*dy - c2 - (x - mean)*c3 with c1 = scale*invstd, c2 = c1*sum(dy)/m, c3 = c1*invstd^2*sum(dy*(x-mean))/m
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads) bn_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
																Slab g, const float* __restrict__ sums, const float* __restrict__ scale,
																const float* __restrict__ save_mean, const float* __restrict__ save_invvar,
																float* dscale, float* dbias, float count, float* zero_ptr, int zero_n, BnAcc acc)
{
	constexpr int UNR = kRowUnroll / 2;
	for (int i = blockIdx.x * kThreads + threadIdx.x; i < zero_n; i += gridDim.x * kThreads) zero_ptr[i] = 0.0f;
	const ItemRange ir = my_items(g);
	int cur = -1;
	uint32_t j = 0;
	bool active = false;
	float c1[VEC], c2[VEC], c3[VEC], mean[VEC];
	for (int it = ir.begin; it < ir.end; it++) {
		const int cb = it / g.row_groups, rg = it - cb * g.row_groups;
		if (cb != cur) {
			cur = cb;
			j = g.col0 + ((uint32_t)cb * kThreads + threadIdx.x) * VEC;
			active = j < g.col0 + g.ncols;
			if (active) {
				#pragma unroll
				for (int e = 0; e < VEC; e++) {
					const int c = (int)fdiv32(j + e, g.sdiv);
					const float invstd = save_invvar[c];
					const float tdy = sums[2 * c], tdyx = sums[2 * c + 1];
					const float dsc = tdyx * invstd;        // sum(dy * xhat)
					mean[e] = save_mean[c];
					c1[e] = scale[c] * invstd;
					c2[e] = c1[e] * tdy / count;
					c3[e] = c1[e] * dsc / count * invstd;
					if (rg == 0 && j + e == (uint32_t)c * g.S) bn_param_grads(acc, c, dsc, tdy, dscale, dbias);
				}
			}
		}
		if (!active) continue;
		const int r0 = rg * kRowUnroll, r1 = min(g.N, r0 + kRowUnroll);
		const size_t off = (size_t)r0 * g.row_stride + j;
		const T* p = x + off;
		const T* q = dy + off;
		T* o = dx + off;
		int r = r0;
		for (; r + UNR <= r1; r += UNR) {
			Pack<T, VEC> v[UNR], w[UNR];
			#pragma unroll
			for (int u = 0; u < UNR; u++) {
				v[u] = *reinterpret_cast<const Pack<T, VEC>*>(p + (size_t)u * g.row_stride);
				w[u] = *reinterpret_cast<const Pack<T, VEC>*>(q + (size_t)u * g.row_stride);
			}
			#pragma unroll
			for (int u = 0; u < UNR; u++) {
				#pragma unroll
				for (int e = 0; e < VEC; e++)
					w[u].v[e] = from_f<T>(c1[e] * to_f<T>(w[u].v[e]) - c2[e] - (to_f<T>(v[u].v[e]) - mean[e]) * c3[e]);
				*reinterpret_cast<Pack<T, VEC>*>(o + (size_t)u * g.row_stride) = w[u];
			}
			p += (size_t)UNR * g.row_stride;
			q += (size_t)UNR * g.row_stride;
			o += (size_t)UNR * g.row_stride;
		}
		for (; r < r1; r++) {
			const Pack<T, VEC> v = *reinterpret_cast<const Pack<T, VEC>*>(p);
			Pack<T, VEC> w = *reinterpret_cast<const Pack<T, VEC>*>(q);
			#pragma unroll
			for (int e = 0; e < VEC; e++) w.v[e] = from_f<T>(c1[e] * to_f<T>(w.v[e]) - c2[e] - (to_f<T>(v.v[e]) - mean[e]) * c3[e]);
			*reinterpret_cast<Pack<T, VEC>*>(o) = w;
			p += g.row_stride;
			q += g.row_stride;
			o += g.row_stride;
		}
	}
}


// ================================================================================================ cluster kernels
// One thread-block CLUSTER per channel: the CL CTAs of a cluster split the N images of channel c.  Phase 1: a CTA streams its
// planes from HBM, keeps them in SHARED MEMORY and accumulates the channel statistics; the partial sums meet through
// DISTRIBUTED SHARED MEMORY (deterministic: fixed trees, partials added in rank order -- no atomics).  Phase 2: the CTA
// normalises out of shared memory and streams the result to HBM.  The tensor crosses HBM exactly once in each direction: 2E
// bytes forward (x, y), 3E backward (x, dy, dx).  Planes need no alignment: a plane is covered by the 16-byte aligned vectors
// that overlap it; lanes outside the plane are masked on the way in and stored as scalars on the way out.
struct ClusterGeo {
	int N, C, S;
	uint32_t plane_step;         // C * S (elements between the planes of consecutive images)
	int CL, rows_per_cta;
	int SP;                      // vector slots per plane (upper bound)
	FastDiv32 spdiv;
	int stash_slots;             // rows_per_cta * SP
	int stream_stores;           // st.global.cs for the results (PZ_BN_STREAM_STORES)
	int bulk_store;              // BULK kernels: results leave by cp.async.bulk too (PZ_BN_BULK_STORE=1; measured no faster, off by default)
};

constexpr int kSlotUnroll = 8;      // 16-byte loads in flight per thread (the backward kernel reads two tensors: 2 x 4)

__device__ __forceinline__ unsigned cluster_ctarank()
{
	unsigned r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem(const float* local, unsigned rank)
{
	unsigned addr = (unsigned)__cvta_generic_to_shared(local), remote;
	float v;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(rank));
	asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
	return v;
}

// results go out with plain stores by default: the next kernel (the activation, the next convolution's gather) reads them, and the
// 126 MB L2 keeps a good part of a layer's output.  PZ_BN_STREAM_STORES=1 restores the evict-first hint (st.global.cs) of the
// first version for A/B runs.
template <typename T, int VEC>
__device__ __forceinline__ void store_streaming(T* dst, const Pack<T, VEC>& v, int streaming)
{
	const uint4 u = *reinterpret_cast<const uint4*>(&v);
	if (streaming) asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
	else *reinterpret_cast<uint4*>(dst) = u;
}

// the first / last vector of an unaligned plane: only the lanes inside the plane are written.  Out of line on purpose -- inlined
// it gets if-converted into the hot loop and doubles its instruction count.
template <typename T, int VEC>
__device__ __noinline__ void store_partial(T* dst, const Pack<T, VEC>& v, uint32_t mask)
{
	#pragma unroll
	for (int e = 0; e < VEC; e++)
		if (mask >> e & 1u) dst[e] = v.v[e];
}

// ---- bulk asynchronous copies (TMA, non-tensor form) of whole planes into the stash.  One elected warp issues a
// cp.async.bulk per plane -- the 16-byte aligned byte range that covers it, which is exactly the plane's stash layout -- so a CTA
// has ALL of its planes (up to 100 KB) in flight at once without holding a register or a scoreboard per load; completion is counted
// by an mbarrier.  (The per-thread loads of the first version kept ~8 KB per CTA in flight: DRAM utilisation 39 %, ncu
// profiles/r02_bn_fwd_ncu_keys.txt.)
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init1(uint32_t bar)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive1(uint32_t bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait0(uint32_t bar)
{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
		"@!p bra WAIT_%=;\n\t}"
		::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
				 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// warp 0 of the CTA: plane `row` of each of `ntensors` tensors -> its SP-vector region of the matching stash
template <typename T, int VEC>
__device__ __forceinline__ void bulk_load_planes(const ClusterGeo& g, uint32_t plane0, uint32_t nrows, const T* const (&src)[2], uint4* const (&dst)[2],
												 int ntensors, uint32_t bar)
{
	const uint32_t lane = threadIdx.x & 31u, S = (uint32_t)g.S;
	for (uint32_t row = lane; row < nrows; row += 32u) {
		const uint32_t lo = plane0 + row * g.plane_step, head = lo & (uint32_t)(VEC - 1);
		const uint32_t bytes = ((head + S + VEC - 1) / VEC) * 16u;
		mbar_expect(bar, bytes * (uint32_t)ntensors);
		for (int t = 0; t < ntensors; t++) bulk_g2s(smem_addr(dst[t] + row * (uint32_t)g.SP), src[t] + (lo - head), bytes, bar);
	}
	__syncwarp();
	if (lane == 0) mbar_arrive1(bar);
}

// the lanes of a plane's stash region that lie OUTSIDE the plane (before its first element in the head vector, after its last one
// up to the end of the SP-vector region) are overwritten with `fill`, chosen so that they add nothing to the statistics: the hot
// loops then run over the stash as a flat array -- no slot decode, no masks (they were issue-bound on both: ~90 instructions per
// 16-byte vector, profiles/r02_bn_fwd_ncu_keys.txt)
template <typename T, int VEC>
__device__ __forceinline__ void stash_fill_outside(uint4* stash, const ClusterGeo& g, uint32_t plane0, uint32_t nrows, T fill, int nthreads)
{
	T* lanes = reinterpret_cast<T*>(stash);
	const uint32_t S = (uint32_t)g.S, width = (uint32_t)g.SP * VEC;
	for (uint32_t row = threadIdx.x; row < nrows; row += nthreads) {
		const uint32_t head = (plane0 + row * g.plane_step) & (uint32_t)(VEC - 1);
		T* rp = lanes + row * width;
		for (uint32_t e = 0; e < head; e++) rp[e] = fill;
		for (uint32_t e = head + S; e < width; e++) rp[e] = fill;
	}
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
// warp 0: the FULL vectors of every plane leave as one bulk copy per plane (shared -> global); the at most two partial vectors
// per plane are written lane by lane by `store_partial_vectors`.  The warp waits until the copies have read shared memory.
template <typename T, int VEC>
__device__ __forceinline__ void bulk_store_planes(const ClusterGeo& g, uint32_t plane0, uint32_t nrows, T* dst, const uint4* stash)
{
	const uint32_t lane = threadIdx.x & 31u, S = (uint32_t)g.S;
	for (uint32_t row = lane; row < nrows; row += 32u) {
		const uint32_t lo = plane0 + row * g.plane_step, head = lo & (uint32_t)(VEC - 1);
		const uint32_t vlo = head ? 1u : 0u, vhi = (head + S) / VEC;
		if (vhi > vlo) bulk_s2g(dst + (lo - head + vlo * VEC), smem_addr(stash + row * (uint32_t)g.SP + vlo), (vhi - vlo) * 16u);
	}
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_store_wait() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// The per-thread partial sums (a few dozen float terms each) are combined in DOUBLE: a float tree loses ~1e-7 of the pivot-shifted
// sums, which is what separates a mean / variance from the correctly rounded one -- and the reference's own unit tests compare
// batch-norm outputs with numpy at atol 1e-8 (Modules/BatchNorm3D.py:57-59).  Deterministic: fixed trees, partials in rank order.
__device__ __forceinline__ double warp_sum_d(double v)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
template <int THREADS>
__device__ __forceinline__ void block_sum2_d(double& a, double& b, double* red /* [2 * THREADS / 32 + 2] */)
{
	a = warp_sum_d(a);
	b = warp_sum_d(b);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (lane == 0) { red[2 * warp] = a; red[2 * warp + 1] = b; }
	__syncthreads();
	if (warp == 0) {
		double x = lane < THREADS / 32 ? red[2 * lane] : 0.0, y = lane < THREADS / 32 ? red[2 * lane + 1] : 0.0;
		x = warp_sum_d(x);
		y = warp_sum_d(y);
		if (lane == 0) { red[2 * (THREADS / 32)] = x; red[2 * (THREADS / 32) + 1] = y; }
	}
	__syncthreads();
	a = red[2 * (THREADS / 32)];
	b = red[2 * (THREADS / 32) + 1];
}

// deterministic block sum of two values; result valid in every thread
template <int THREADS>
__device__ __forceinline__ void block_sum2(float& a, float& b, float* red /* [2 * THREADS / 32 + 2] */)
{
	a = warp_sum(a);
	b = warp_sum(b);
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (lane == 0) { red[2 * warp] = a; red[2 * warp + 1] = b; }
	__syncthreads();
	if (warp == 0) {
		float x = lane < THREADS / 32 ? red[2 * lane] : 0.0f, y = lane < THREADS / 32 ? red[2 * lane + 1] : 0.0f;
		x = warp_sum(x);
		y = warp_sum(y);
		if (lane == 0) { red[2 * (THREADS / 32)] = x; red[2 * (THREADS / 32) + 1] = y; }
	}
	__syncthreads();
	a = red[2 * (THREADS / 32)];
	b = red[2 * (THREADS / 32) + 1];
}

// Slot decode.  Slot s of a CTA = vector v = s % SP of its plane s / SP; the plane starts at element lo = (n*C + c)*S, the
// vector at e0 = lo - head + v*VEC with head = lo % VEC.  Only vector 0 (when head > 0) and the vector holding the plane's end
// (when (head + S) % VEC > 0) are partial.  The hot loops touch FULL vectors only -- no lane masks, 32-bit index math (the host
// checks N*C*S < 2^31), about 12 integer instructions per 16-byte vector; the at most two partial vectors per plane are done
// by a short masked loop afterwards (the kernels were issue-bound on the general decode: profiles/r02_bn_fwd_ncu_keys.txt).
template <int VEC>
struct SlotGeo {
	uint32_t lo, head, v;
	// full vectors of the plane: vlo <= v < vhi
	__device__ __forceinline__ bool full(uint32_t S) const { return v >= (head > 0u ? 1u : 0u) && (v + 1u) * VEC <= head + S; }
	__device__ __forceinline__ uint32_t e0() const { return lo - head + v * VEC; }
};

template <int VEC>
__device__ __forceinline__ SlotGeo<VEC> slot_decode(const ClusterGeo& g, uint32_t plane0, uint32_t slot)
{
	SlotGeo<VEC> q;
	const uint32_t row = fdiv32(slot, g.spdiv);
	q.v = slot - row * (uint32_t)g.SP;
	q.lo = plane0 + row * g.plane_step;
	q.head = q.lo & (uint32_t)(VEC - 1);
	return q;
}

// partial vector `which` (0: the head, 1: the tail) of plane `row`: vector index and lane mask, 0 when that vector is not partial
template <int VEC>
__device__ __forceinline__ uint32_t partial_vector(const ClusterGeo& g, uint32_t plane0, uint32_t row, int which, uint32_t& v, uint32_t& e0)
{
	const uint32_t lo = plane0 + row * g.plane_step, head = lo & (uint32_t)(VEC - 1), S = (uint32_t)g.S;
	const uint32_t vend = (head + S) / VEC, rem = head + S - vend * VEC;      // the plane ends `rem` lanes into vector vend
	uint32_t mask;
	if (which == 0) {
		if (head == 0u) return 0u;
		v = 0u;
		mask = ~((1u << head) - 1u) & ((1u << VEC) - 1u);
		if (vend == 0u) mask &= (1u << rem) - 1u;                               // plane shorter than one vector
	} else {
		if (rem == 0u || (vend == 0u && head > 0u)) return 0u;                  // aligned end, or already handled as the head
		v = vend;
		mask = (1u << rem) - 1u;
	}
	e0 = lo - head + v * VEC;
	return mask;
}

template <typename T, int VEC>
__device__ __forceinline__ void store_partial_vectors(const ClusterGeo& g, uint32_t plane0, uint32_t nrows, T* dst, const uint4* stash, int nthreads)
{
	using P = Pack<T, VEC>;
	for (uint32_t i = threadIdx.x; i < 2u * nrows; i += nthreads) {
		uint32_t vv, e0;
		const uint32_t mask = partial_vector<VEC>(g, plane0, i >> 1, (int)(i & 1u), vv, e0);
		if (!mask) continue;
		const uint4 raw = stash[(i >> 1) * (uint32_t)g.SP + vv];
		store_partial<T, VEC>(dst + e0, *reinterpret_cast<const P*>(&raw), mask);
	}
}

// the cluster-wide sum of the CTAs' (s1, s2) through distributed shared memory, in rank order
__device__ __forceinline__ double ld_dsmem_d(const double* local, unsigned rank)
{
	unsigned addr = (unsigned)__cvta_generic_to_shared(local), remote;
	double v;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(addr), "r"(rank));
	asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
	return v;
}
__device__ __forceinline__ void cluster_sum2_d(double& s1, double& s2, double* part, int CL)
{
	if (CL <= 1) return;
	if (threadIdx.x == 0) { part[0] = s1; part[1] = s2; }
	cluster_sync_all();
	s1 = 0.0;
	s2 = 0.0;
	for (int r = 0; r < CL; r++) { s1 += ld_dsmem_d(&part[0], (unsigned)r); s2 += ld_dsmem_d(&part[1], (unsigned)r); }
	cluster_sync_all();                 // nobody leaves (or overwrites `part`) while a peer still reads it
}
__device__ __forceinline__ void cluster_sum2(float& s1, float& s2, float* part, int CL)
{
	if (CL <= 1) return;
	if (threadIdx.x == 0) { part[0] = s1; part[1] = s2; }
	cluster_sync_all();
	s1 = 0.0f;
	s2 = 0.0f;
	for (int r = 0; r < CL; r++) { s1 += ld_dsmem(&part[0], (unsigned)r); s2 += ld_dsmem(&part[1], (unsigned)r); }
	cluster_sync_all();                 // nobody leaves (or overwrites `part`) while a peer still reads it
}

template <typename T, int VEC, int THREADS, bool BULK>
__global__ void __launch_bounds__(THREADS) bn_fwd_cluster_kernel(const T* __restrict__ x, T* __restrict__ y, T* __restrict__ z, ClusterGeo g,
																  const float* __restrict__ scale, const float* __restrict__ bias,
																  float* mean_io, float* var_io, float* save_mean, float* save_invvar,
																  float eps, float factor)
{
	extern __shared__ uint4 stash[];                 // [stash_slots]: this CTA's planes, as the aligned vectors that cover them
	__shared__ double red[2 * THREADS / 32 + 2];
	__shared__ double part[2];
	using P = Pack<T, VEC>;
	const unsigned rank = g.CL > 1 ? cluster_ctarank() : 0u;
	const int c = (int)(blockIdx.x / (unsigned)g.CL);
	const int n0 = (int)rank * g.rows_per_cta, n1 = min(g.N, n0 + g.rows_per_cta);
	const int nslots = max(0, n1 - n0) * g.SP;
	const float pivot = to_f<T>(x[(size_t)c * g.S]);

	// ---- phase 1: HBM -> shared memory, sum(x - pivot), sum((x - pivot)^2)
	const uint32_t nrows = (uint32_t)max(0, n1 - n0), S = (uint32_t)g.S;
	const uint32_t plane0 = ((uint32_t)n0 * (uint32_t)g.C + (uint32_t)c) * S;
	float s1 = 0.0f, s2 = 0.0f;
	if (BULK) {
		__shared__ uint64_t bar;
		const uint32_t barp = smem_addr(&bar);
		if (threadIdx.x == 0) mbar_init1(barp);
		__syncthreads();
		if (threadIdx.x < 32) {
			const T* const src[2] = {x, x};
			uint4* const dst[2] = {stash, stash};
			bulk_load_planes<T, VEC>(g, plane0, nrows, src, dst, 1, barp);
		}
		mbar_wait0(barp);
		stash_fill_outside<T, VEC>(stash, g, plane0, nrows, from_f<T>(pivot), THREADS);      // x - pivot = 0 there
		__syncthreads();
		for (uint32_t slot = threadIdx.x; slot < (uint32_t)nslots; slot += THREADS) {
			const uint4 raw = stash[slot];
			const P v = *reinterpret_cast<const P*>(&raw);
			#pragma unroll
			for (int e = 0; e < VEC; e++) { const float d = to_f<T>(v.v[e]) - pivot; s1 += d; s2 = fmaf(d, d, s2); }
		}
	} else {
	for (uint32_t base = threadIdx.x; base < (uint32_t)nslots; base += THREADS * kSlotUnroll) {
		P v[kSlotUnroll];
		bool ok[kSlotUnroll];
		#pragma unroll
		for (int u = 0; u < kSlotUnroll; u++) {
			const uint32_t slot = base + u * THREADS;
			const SlotGeo<VEC> q = slot_decode<VEC>(g, plane0, slot);
			ok[u] = slot < (uint32_t)nslots && q.full(S);
			if (ok[u]) v[u] = *reinterpret_cast<const P*>(x + q.e0());
		}
		#pragma unroll
		for (int u = 0; u < kSlotUnroll; u++) {
			if (!ok[u]) continue;
			stash[base + u * THREADS] = *reinterpret_cast<const uint4*>(&v[u]);
			#pragma unroll
			for (int e = 0; e < VEC; e++) { const float d = to_f<T>(v[u].v[e]) - pivot; s1 += d; s2 = fmaf(d, d, s2); }
		}
	}
	// the (at most two) partial vectors of every plane
	for (uint32_t i = threadIdx.x; i < 2u * nrows; i += THREADS) {
		uint32_t vv, e0;
		const uint32_t mask = partial_vector<VEC>(g, plane0, i >> 1, (int)(i & 1u), vv, e0);
		if (!mask) continue;
		const P pv = *reinterpret_cast<const P*>(x + e0);
		stash[(i >> 1) * (uint32_t)g.SP + vv] = *reinterpret_cast<const uint4*>(&pv);
		#pragma unroll
		for (int e = 0; e < VEC; e++)
			if (mask >> e & 1u) { const float d = to_f<T>(pv.v[e]) - pivot; s1 += d; s2 = fmaf(d, d, s2); }
	}
	}
	double d1 = s1, d2 = s2;
	block_sum2_d<THREADS>(d1, d2, red);
	cluster_sum2_d(d1, d2, part, g.CL);

	const float count = (float)g.N * (float)g.S;
	const double dcount = (double)g.N * (double)g.S, ddmean = d1 / dcount;
	const float mean = (float)((double)pivot + ddmean);
	const float var = (float)fmax(d2 / dcount - ddmean * ddmean, 0.0);      // biased: used for normalisation
	const float invstd = (float)(1.0 / sqrt(fmax(d2 / dcount - ddmean * ddmean, 0.0) + (double)eps));
	if (rank == 0 && threadIdx.x == 0) {
		save_mean[c] = mean;
		save_invvar[c] = invstd;
		// cuDNN keeps the UNBIASED variance in the running estimate (pinned by tests/golden/ref_cuda_ops.npz bn*/runvar)
		const float uvar = count > 1.0f ? var * (count / (count - 1.0f)) : var;
		mean_io[c] = (1.0f - factor) * mean_io[c] + factor * mean;
		var_io[c] = (1.0f - factor) * var_io[c] + factor * uvar;
	}
	const float a = scale[c] * invstd, b = bias[c];       // y = (x - mean) * a + b: x - mean is exact where y is small

	// ---- phase 2: shared memory -> y = a * x + b -> HBM
	__syncthreads();                              // partial vectors were stashed by other threads than the ones that read them
	if (BULK && g.bulk_store) {
		// normalise in place, then the planes leave through the copy engine
		for (uint32_t slot = threadIdx.x; slot < (uint32_t)nslots; slot += THREADS) {
			const uint4 raw = stash[slot];
			P v = *reinterpret_cast<const P*>(&raw);
			#pragma unroll
			for (int e = 0; e < VEC; e++) v.v[e] = from_f<T>(fmaf(to_f<T>(v.v[e]) - mean, a, b));
			stash[slot] = *reinterpret_cast<const uint4*>(&v);
		}
		fence_proxy_async();
		__syncthreads();
		if (threadIdx.x < 32) bulk_store_planes<T, VEC>(g, plane0, nrows, y, stash);
		store_partial_vectors<T, VEC>(g, plane0, nrows, y, stash, THREADS);
		if (threadIdx.x < 32) bulk_store_wait();
		return;
	}
	for (uint32_t slot = threadIdx.x; slot < (uint32_t)nslots; slot += THREADS) {
		const SlotGeo<VEC> q = slot_decode<VEC>(g, plane0, slot);
		if (!q.full(S)) continue;
		const uint4 raw = stash[slot];
		P v = *reinterpret_cast<const P*>(&raw);
		#pragma unroll
		for (int e = 0; e < VEC; e++) v.v[e] = from_f<T>(fmaf(to_f<T>(v.v[e]) - mean, a, b));
		store_streaming<T, VEC>(y + q.e0(), v, g.stream_stores);
		if (z != nullptr) {
			// the ReLU the next module applies to y (pz_bn_fwd_train_relu), from the ROUNDED y like the separate kernel: y * (y > 0)
			#pragma unroll
			for (int e = 0; e < VEC; e++) { const float f = to_f<T>(v.v[e]); v.v[e] = from_f<T>(f * (f > 0.0f ? 1.0f : 0.0f)); }
			store_streaming<T, VEC>(z + q.e0(), v, 0);
		}
	}
	for (uint32_t i = threadIdx.x; i < 2u * nrows; i += THREADS) {
		uint32_t vv, e0;
		const uint32_t mask = partial_vector<VEC>(g, plane0, i >> 1, (int)(i & 1u), vv, e0);
		if (!mask) continue;
		const uint4 raw = stash[(i >> 1) * (uint32_t)g.SP + vv];
		P v = *reinterpret_cast<const P*>(&raw);
		#pragma unroll
		for (int e = 0; e < VEC; e++) v.v[e] = from_f<T>(fmaf(to_f<T>(v.v[e]) - mean, a, b));
		store_partial<T, VEC>(y + e0, v, mask);
		if (z != nullptr) {
			#pragma unroll
			for (int e = 0; e < VEC; e++) { const float f = to_f<T>(v.v[e]); v.v[e] = from_f<T>(f * (f > 0.0f ? 1.0f : 0.0f)); }
			store_partial<T, VEC>(z + e0, v, mask);
		}
	}
}

// ---- persistent, double-buffered forward pass.  One cluster of CL CTAs (one CTA per SM, two stash buffers) walks over the channels
// c = cluster, cluster + nclusters, ...: while the statistics / cluster reduction / normalisation / store of channel i run, the planes
// of channel i + 1 are already in flight into the other buffer (cp.async.bulk both ways), so an SM always has a channel's worth
// of loads outstanding and the load -> reduce -> barrier -> store life cycle of the one-shot kernel no longer idles the memory system.
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity)
{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"WAITP_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@!p bra WAITP_%=;\n\t}"
		::"r"(bar), "r"(parity) : "memory");
}

template <typename T, int VEC, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) bn_fwd_persistent_kernel(const T* __restrict__ x, T* __restrict__ y, ClusterGeo g, int nclusters,
																		const float* __restrict__ scale, const float* __restrict__ bias,
																		float* mean_io, float* var_io, float* save_mean, float* save_invvar,
																		float eps, float factor)
{
	extern __shared__ uint4 stash2[];                // [2][stash_slots]
	__shared__ double red[2 * THREADS / 32 + 2];
	__shared__ double part[2];
	__shared__ uint64_t bars[2];
	using P = Pack<T, VEC>;
	const unsigned rank = g.CL > 1 ? cluster_ctarank() : 0u;
	const int cluster = (int)(blockIdx.x / (unsigned)g.CL);
	const int n0 = (int)rank * g.rows_per_cta, n1 = min(g.N, n0 + g.rows_per_cta);
	const uint32_t nrows = (uint32_t)max(0, n1 - n0), S = (uint32_t)g.S;
	const uint32_t nslots = nrows * (uint32_t)g.SP;
	const float count = (float)g.N * (float)g.S;

	if (threadIdx.x == 0) { mbar_init1(smem_addr(&bars[0])); mbar_init1(smem_addr(&bars[1])); }
	__syncthreads();

	auto plane0_of = [&](int c) { return ((uint32_t)n0 * (uint32_t)g.C + (uint32_t)c) * S; };
	auto issue = [&](int c, int buf) {           // warp 0
		const T* const src[2] = {x, x};
		uint4* const dst[2] = {stash2 + (size_t)buf * g.stash_slots, stash2};
		bulk_load_planes<T, VEC>(g, plane0_of(c), nrows, src, dst, 1, smem_addr(&bars[buf]));
	};

	int c = cluster;
	if (c < g.C && threadIdx.x < 32) issue(c, 0);
	float pivot = c < g.C ? to_f<T>(x[(size_t)c * g.S]) : 0.0f;
	for (int it = 0; c < g.C; it++, c += nclusters) {
		const int buf = it & 1, cnext = c + nclusters;
		uint4* stash = stash2 + (size_t)buf * g.stash_slots;
		const uint32_t plane0 = plane0_of(c);
		float pivot_next = 0.0f;
		if (cnext < g.C) {
			// the other buffer is free once the bulk stores of the previous channel have read it (warp 0 issued them)
			if (threadIdx.x < 32) {
				bulk_store_wait();
				issue(cnext, buf ^ 1);
			}
			pivot_next = to_f<T>(x[(size_t)cnext * g.S]);
		}
		mbar_wait_parity(smem_addr(&bars[buf]), (uint32_t)(it >> 1) & 1u);

		stash_fill_outside<T, VEC>(stash, g, plane0, nrows, from_f<T>(pivot), THREADS);
		__syncthreads();
		float s1 = 0.0f, s2 = 0.0f;
		for (uint32_t slot = threadIdx.x; slot < nslots; slot += THREADS) {
			const uint4 raw = stash[slot];
			const P v = *reinterpret_cast<const P*>(&raw);
			#pragma unroll
			for (int e = 0; e < VEC; e++) { const float d = to_f<T>(v.v[e]) - pivot; s1 += d; s2 = fmaf(d, d, s2); }
		}
		double d1 = s1, d2 = s2;
		block_sum2_d<THREADS>(d1, d2, red);
		cluster_sum2_d(d1, d2, part, g.CL);

		const double dcount = (double)g.N * (double)g.S, ddmean = d1 / dcount;
		const float mean = (float)((double)pivot + ddmean);
		const float var = (float)fmax(d2 / dcount - ddmean * ddmean, 0.0);
		const float invstd = (float)(1.0 / sqrt(fmax(d2 / dcount - ddmean * ddmean, 0.0) + (double)eps));
		if (rank == 0 && threadIdx.x == 0) {
			save_mean[c] = mean;
			save_invvar[c] = invstd;
			const float uvar = count > 1.0f ? var * (count / (count - 1.0f)) : var;
			mean_io[c] = (1.0f - factor) * mean_io[c] + factor * mean;
			var_io[c] = (1.0f - factor) * var_io[c] + factor * uvar;
		}
		const float a = scale[c] * invstd, b = bias[c];       // y = (x - mean) * a + b: x - mean is exact where y is small

		for (uint32_t slot = threadIdx.x; slot < nslots; slot += THREADS) {
			const uint4 raw = stash[slot];
			P v = *reinterpret_cast<const P*>(&raw);
			#pragma unroll
			for (int e = 0; e < VEC; e++) v.v[e] = from_f<T>(fmaf(to_f<T>(v.v[e]) - mean, a, b));
			stash[slot] = *reinterpret_cast<const uint4*>(&v);
		}
		fence_proxy_async();
		__syncthreads();
		if (threadIdx.x < 32) bulk_store_planes<T, VEC>(g, plane0, nrows, y, stash);
		store_partial_vectors<T, VEC>(g, plane0, nrows, y, stash, THREADS);
		__syncthreads();                    // `red` / the partial-vector reads of this channel are done before the next one starts
		pivot = pivot_next;
	}
	if (threadIdx.x < 32) bulk_store_wait();
}

template <typename T, int VEC, int THREADS, bool BULK>
__global__ void __launch_bounds__(THREADS) bn_bwd_cluster_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
																  ClusterGeo g, const float* __restrict__ scale,
																  const float* __restrict__ save_mean, const float* __restrict__ save_invvar,
																  float* dscale, float* dbias, BnAcc acc)
{
	extern __shared__ uint4 stash[];                 // [2][stash_slots]: x planes, then dy planes
	__shared__ double red[2 * THREADS / 32 + 2];
	__shared__ double part[2];
	using P = Pack<T, VEC>;
	const unsigned rank = g.CL > 1 ? cluster_ctarank() : 0u;
	const int c = (int)(blockIdx.x / (unsigned)g.CL);
	const int n0 = (int)rank * g.rows_per_cta, n1 = min(g.N, n0 + g.rows_per_cta);
	const int nslots = max(0, n1 - n0) * g.SP;
	const float mean = save_mean[c], invstd = save_invvar[c];
	constexpr int UNR = kSlotUnroll / 2;              // two tensors per slot
	uint4* stash_dy = stash + g.stash_slots;

	// ---- phase 1: HBM -> shared memory, sum(dy), sum(dy * (x - mean))
	const uint32_t nrows = (uint32_t)max(0, n1 - n0), S = (uint32_t)g.S;
	const uint32_t plane0 = ((uint32_t)n0 * (uint32_t)g.C + (uint32_t)c) * S;
	float s1 = 0.0f, s2 = 0.0f;
	if (BULK) {
		__shared__ uint64_t bar;
		const uint32_t barp = smem_addr(&bar);
		if (threadIdx.x == 0) mbar_init1(barp);
		__syncthreads();
		if (threadIdx.x < 32) {
			const T* const src[2] = {x, dy};
			uint4* const dst[2] = {stash, stash_dy};
			bulk_load_planes<T, VEC>(g, plane0, nrows, src, dst, 2, barp);
		}
		mbar_wait0(barp);
		stash_fill_outside<T, VEC>(stash_dy, g, plane0, nrows, from_f<T>(0.0f), THREADS);    // no gradient outside the plane ...
		stash_fill_outside<T, VEC>(stash, g, plane0, nrows, from_f<T>(mean), THREADS);       // ... times a finite (x - mean)
		__syncthreads();
		for (uint32_t slot = threadIdx.x; slot < (uint32_t)nslots; slot += THREADS) {
			const uint4 rx = stash[slot], rg = stash_dy[slot];
			const P v = *reinterpret_cast<const P*>(&rx), w = *reinterpret_cast<const P*>(&rg);
			#pragma unroll
			for (int e = 0; e < VEC; e++) {
				const float gv = to_f<T>(w.v[e]);
				s1 += gv;
				s2 = fmaf(gv, to_f<T>(v.v[e]) - mean, s2);
			}
		}
	} else {
	for (uint32_t base = threadIdx.x; base < (uint32_t)nslots; base += THREADS * UNR) {
		P v[UNR], w[UNR];
		bool ok[UNR];
		#pragma unroll
		for (int u = 0; u < UNR; u++) {
			const uint32_t slot = base + u * THREADS;
			const SlotGeo<VEC> q = slot_decode<VEC>(g, plane0, slot);
			ok[u] = slot < (uint32_t)nslots && q.full(S);
			if (ok[u]) {
				v[u] = *reinterpret_cast<const P*>(x + q.e0());
				w[u] = *reinterpret_cast<const P*>(dy + q.e0());
			}
		}
		#pragma unroll
		for (int u = 0; u < UNR; u++) {
			if (!ok[u]) continue;
			stash[base + u * THREADS] = *reinterpret_cast<const uint4*>(&v[u]);
			stash_dy[base + u * THREADS] = *reinterpret_cast<const uint4*>(&w[u]);
			#pragma unroll
			for (int e = 0; e < VEC; e++) {
				const float gv = to_f<T>(w[u].v[e]);
				s1 += gv;
				s2 = fmaf(gv, to_f<T>(v[u].v[e]) - mean, s2);
			}
		}
	}
	for (uint32_t i = threadIdx.x; i < 2u * nrows; i += THREADS) {
		uint32_t vv, e0;
		const uint32_t mask = partial_vector<VEC>(g, plane0, i >> 1, (int)(i & 1u), vv, e0);
		if (!mask) continue;
		const P pv = *reinterpret_cast<const P*>(x + e0), pw = *reinterpret_cast<const P*>(dy + e0);
		stash[(i >> 1) * (uint32_t)g.SP + vv] = *reinterpret_cast<const uint4*>(&pv);
		stash_dy[(i >> 1) * (uint32_t)g.SP + vv] = *reinterpret_cast<const uint4*>(&pw);
		#pragma unroll
		for (int e = 0; e < VEC; e++)
			if (mask >> e & 1u) {
				const float gv = to_f<T>(pw.v[e]);
				s1 += gv;
				s2 = fmaf(gv, to_f<T>(pv.v[e]) - mean, s2);
			}
	}
	}
	double d1 = s1, d2 = s2;
	block_sum2_d<THREADS>(d1, d2, red);
	cluster_sum2_d(d1, d2, part, g.CL);
	s1 = (float)d1;
	s2 = (float)d2;

	const float count = (float)g.N * (float)g.S;
	const float dsc = s2 * invstd;               // sum(dy * xhat)
	if (rank == 0 && threadIdx.x == 0) bn_param_grads(acc, c, dsc, s1, dscale, dbias);
	const float c1 = scale[c] * invstd, c2 = c1 * s1 / count, c3 = c1 * dsc / count * invstd;

	// ---- phase 2: shared memory -> dx = c1*dy - c2 - (x - mean)*c3 -> HBM
	__syncthreads();
	if (BULK && g.bulk_store) {
		for (uint32_t slot = threadIdx.x; slot < (uint32_t)nslots; slot += THREADS) {
			const uint4 rx = stash[slot], rg = stash_dy[slot];
			const P v = *reinterpret_cast<const P*>(&rx);
			P w = *reinterpret_cast<const P*>(&rg);
			#pragma unroll
			for (int e = 0; e < VEC; e++) w.v[e] = from_f<T>(c1 * to_f<T>(w.v[e]) - c2 - (to_f<T>(v.v[e]) - mean) * c3);
			stash_dy[slot] = *reinterpret_cast<const uint4*>(&w);
		}
		fence_proxy_async();
		__syncthreads();
		if (threadIdx.x < 32) bulk_store_planes<T, VEC>(g, plane0, nrows, dx, stash_dy);
		store_partial_vectors<T, VEC>(g, plane0, nrows, dx, stash_dy, THREADS);
		if (threadIdx.x < 32) bulk_store_wait();
		return;
	}
	for (uint32_t slot = threadIdx.x; slot < (uint32_t)nslots; slot += THREADS) {
		const SlotGeo<VEC> q = slot_decode<VEC>(g, plane0, slot);
		if (!q.full(S)) continue;
		const uint4 rx = stash[slot], rg = stash_dy[slot];
		const P v = *reinterpret_cast<const P*>(&rx);
		P w = *reinterpret_cast<const P*>(&rg);
		#pragma unroll
		for (int e = 0; e < VEC; e++) w.v[e] = from_f<T>(c1 * to_f<T>(w.v[e]) - c2 - (to_f<T>(v.v[e]) - mean) * c3);
		store_streaming<T, VEC>(dx + q.e0(), w, g.stream_stores);
	}
	for (uint32_t i = threadIdx.x; i < 2u * nrows; i += THREADS) {
		uint32_t vv, e0;
		const uint32_t mask = partial_vector<VEC>(g, plane0, i >> 1, (int)(i & 1u), vv, e0);
		if (!mask) continue;
		const uint4 rx = stash[(i >> 1) * (uint32_t)g.SP + vv], rg = stash_dy[(i >> 1) * (uint32_t)g.SP + vv];
		const P v = *reinterpret_cast<const P*>(&rx);
		P w = *reinterpret_cast<const P*>(&rg);
		#pragma unroll
		for (int e = 0; e < VEC; e++) w.v[e] = from_f<T>(c1 * to_f<T>(w.v[e]) - c2 - (to_f<T>(v.v[e]) - mean) * c3);
		store_partial<T, VEC>(dx + e0, w, mask);
	}
}

// ---- host side of the cluster kernels
struct ClusterPlan {
	ClusterGeo g;
	int threads;             // 256 or 512
	size_t smem;             // the stash
	bool ok;
	bool bulk;               // planes arrive by cp.async.bulk (the tensor must end on a 16-byte boundary: the last plane's cover is read whole)
	int grid_override;       // persistent kernels: number of CTAs (0: one cluster per channel)
};

int env_int(const char* name, int dflt)
{
	const char* e = getenv(name);
	return e ? atoi(e) : dflt;
}

constexpr size_t kStashTwoPerSm = 100 * 1024, kStashMax = 200 * 1024;

// `tensors`: 1 forward (x), 2 backward (x, dy) kept in shared memory between the phases
ClusterPlan make_cluster_plan(std::initializer_list<const void*> ptrs, int64_t N, int64_t C, int64_t S, size_t es, int tensors)
{
	static const int force_threads = env_int("PZ_BN_THREADS", 0), force_cl = env_int("PZ_BN_CL", 0);
	static const int max_cl = env_int("PZ_BN_MAX_CL", 8);              // 16 needs the non-portable cluster size opt-in
	static const size_t stash_target = (size_t)env_int("PZ_BN_STASH_KB", 100) * 1024;
	ClusterPlan p{};
	const int vec = (int)(16 / es);
	p.ok = getenv("PZ_BN_NO_CLUSTER") == nullptr && N * C * S < (1ll << 31) - 64 && C < (1 << 24);      // 32-bit element offsets
	for (const void* q : ptrs) p.ok = p.ok && ((uintptr_t)q % 16 == 0);
	if (!p.ok) return p;

	ClusterGeo& g = p.g;
	g.N = (int)N; g.C = (int)C; g.S = (int)S;
	g.plane_step = (uint32_t)(C * S);
	g.SP = (int)((S + 2 * (vec - 1)) / vec);           // aligned vectors that can overlap a plane starting anywhere
	if ((S % vec) == 0) g.SP = (int)(S / vec);         // ... every plane is aligned when S is a multiple of the vector width
	g.spdiv = make_fastdiv32((uint32_t)g.SP);

	// smallest cluster whose CTAs can keep their planes in shared memory with two CTAs per SM (a bigger cluster when the machine
	// would not be filled); failing that the two-kernel path
	const int sms = pz_num_sms();
	int CL = 0;
	for (int cl = 1; cl <= max_cl; cl *= 2) {
		if (cl > 1 && cl > N) break;
		const size_t bytes = (size_t)pz_cdiv(N, cl) * g.SP * 16 * tensors;
		if (bytes <= stash_target && (C * cl >= 2 * sms || cl == max_cl || cl * 2 > N)) { CL = cl; break; }
	}
	// (a stash that leaves room for ONE CTA per SM only -- planes above ~100 KB per CTA even in a cluster of 8 -- measured slower
	// than the two-kernel path: profiles/r02_bn.md)
	if (force_cl > 0 && force_cl <= 16 && (force_cl & (force_cl - 1)) == 0 && force_cl <= (N > 1 ? N : 1)) CL = force_cl;
	if (CL == 0) { p.ok = false; return p; }

	g.CL = CL;
	g.rows_per_cta = (int)pz_cdiv(N, CL);
	const long long slots = (long long)g.rows_per_cta * g.SP;
	g.stash_slots = (int)slots;
	p.smem = (size_t)slots * 16 * tensors;
	if (p.smem > kStashMax) { p.ok = false; return p; }
	p.threads = slots >= 2048 ? 512 : 256;
	if (force_threads == 256 || force_threads == 512) p.threads = force_threads;
	static const bool bulk_on = env_int("PZ_BN_BULK", 1) != 0, bulk_store_on = env_int("PZ_BN_BULK_STORE", 0) != 0;
	p.bulk = bulk_on && (N * C * S) % vec == 0;
	g.bulk_store = bulk_store_on ? 1 : 0;
	static const bool stream_on = env_int("PZ_BN_STREAM_STORES", 0) != 0;
	g.stream_stores = stream_on ? 1 : 0;
	return p;
}

// how many clusters of this shape the device holds at the same time (the GPCs bound it: a cluster lives inside one GPC), cached per
// (kernel, cluster size, shared memory)
template <typename K>
int resident_clusters(K kernel, const ClusterPlan& p)
{
	static std::mutex mu;
	static std::map<std::pair<int, size_t>, int> cache;
	std::lock_guard<std::mutex> lock(mu);
	const auto key = std::make_pair(p.g.CL, p.smem);
	auto it = cache.find(key);
	if (it != cache.end()) return it->second;
	int n = 0;
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3((unsigned)(pz_num_sms() / p.g.CL * p.g.CL));
	cfg.blockDim = dim3((unsigned)p.threads);
	cfg.dynamicSmemBytes = p.smem;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = (unsigned)p.g.CL;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kStashMax + 24 * 1024)) != cudaSuccess ||
		cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) {
		cudaGetLastError();
		n = 0;
	}
	cache[key] = n;
	return n;
}

template <typename K, typename... Args>
int launch_cluster(K kernel, const ClusterPlan& p, cudaStream_t s, Args... args)
{
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3((unsigned)(p.grid_override > 0 ? p.grid_override : p.g.C * p.g.CL));
	cfg.blockDim = dim3((unsigned)p.threads);
	cfg.dynamicSmemBytes = p.smem;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = (unsigned)p.g.CL;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	if (p.smem > 48 * 1024) {
		// (idempotent; a per-kernel flag would need one static per instantiation)
		PZ_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kStashMax + 24 * 1024)));
	}
	if (p.g.CL > 8) PZ_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
	PZ_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
	pz_count_launch(1);
	return PZ_OK;
}

// ------------------------------------------------------------------------------------------ host side
// Two [C][2] fp32 accumulator buffers, library-owned and used in stream order.  A pass accumulates into one of them and its
// apply kernel clears what the previous pass left in the other, which the next pass will use.
struct SumsPair {
	float* pivots;       // [C] per-channel pivots, written by the statistics kernel (behind the two sums of every channel)
	float* use;          // all zero over [0, C)
	float* clear;        // to be cleared by this pass over [0, clear_n)
	int clear_n;
};

float* g_sums[2] = {nullptr, nullptr};
int64_t g_sums_cap = 0, g_sums_dirty[2] = {0, 0};
int g_sums_cur = 0;
std::mutex g_sums_mu;

bool acquire_sums(int64_t C, SumsPair& sp)
{
	float* (&buf)[2] = g_sums;
	int64_t& cap = g_sums_cap;
	int64_t (&dirty)[2] = g_sums_dirty;
	int& cur = g_sums_cur;
	std::lock_guard<std::mutex> lock(g_sums_mu);
	if (C > cap) {
		// growth retires the old buffers instead of freeing them: launches (or captured graphs) that hold them stay valid
		const int64_t want = C < 65536 ? 65536 : C + C / 2;
		for (int i = 0; i < 2; i++) {
			buf[i] = nullptr;
			if (cudaMalloc((void**)&buf[i], (size_t)want * 3 * sizeof(float)) != cudaSuccess) { cap = 0; cudaGetLastError(); return false; }
			cudaMemset(buf[i], 0, (size_t)want * 3 * sizeof(float));
			dirty[i] = 0;
		}
		cap = want;
	}
	const int other = 1 - cur;
	sp.use = buf[cur];
	sp.pivots = buf[cur] + 2 * cap;
	sp.clear = buf[other];
	sp.clear_n = (int)(dirty[other] * 2);
	dirty[cur] = C;
	dirty[other] = 0;
	cur = other;
	return true;
}

struct Plan {
	int vec;                 // columns per thread
	int64_t slab_channels;   // channels per slab (all of them when the tensor fits)
	int rows_per_block, row_blocks;
};

template <typename T>
Plan make_plan(std::initializer_list<const void*> ptrs, int64_t N, int64_t C, int64_t S, int tensors)
{
	Plan p;
	constexpr int V = 16 / sizeof(T);
	bool aligned = (C * S) % V == 0;
	for (const void* q : ptrs) aligned = aligned && ((uintptr_t)q % 16 == 0);
	p.vec = aligned ? V : 1;

	// slabs: whole channels, column start a multiple of the vector width, working set <= kSlabBytes
	const double chan_bytes = (double)N * S * sizeof(T) * tensors;
	int64_t sc = (int64_t)(slab_bytes() / chan_bytes);
	if (sc >= C) sc = C;
	else {
		int64_t step = 1;                         // smallest channel count whose column count is a multiple of vec
		while ((step * S) % p.vec != 0) step *= 2;
		sc = sc / step * step;
		if (sc < step) sc = step;
		if (sc > C) sc = C;
	}
	p.slab_channels = sc;

	p.rows_per_block = kRowUnroll;
	p.row_blocks = (int)pz_cdiv(N, kRowUnroll);
	return p;
}

Slab make_slab(const Plan& p, int64_t N, int64_t C, int64_t S, int64_t c0, int64_t c1)
{
	Slab g;
	g.row_stride = C * S;
	g.N = (int)N;
	g.rows_per_block = p.rows_per_block;
	g.col0 = (uint32_t)(c0 * S);
	g.ncols = (uint32_t)((c1 - c0) * S);
	g.sdiv = make_fastdiv32((uint32_t)S);
	g.S = (uint32_t)S;
	g.row_groups = p.row_blocks;
	const int64_t col_blocks = pz_cdiv(g.ncols, (int64_t)kThreads * p.vec);
	g.items = (int)(col_blocks * g.row_groups);
	const int64_t ctas = (int64_t)pz_num_sms() * (2048 / kThreads);
	g.items_per_cta = (int)pz_cdiv(g.items, ctas);
	return g;
}

#define PZ_BN_LAUNCH(VECV, KERNEL, SMEM, ...)                                                                   \
	do {                                                                                                        \
		const dim3 grid((unsigned)pz_cdiv(g.items, g.items_per_cta));                                           \
		KERNEL<<<grid, kThreads, SMEM, s>>>(__VA_ARGS__);                                                       \
		pz_count_launch(1);                                                                                     \
	} while (0)

template <typename T>
int fwd_train(const void* x, void* y, int64_t N, int64_t C, int64_t S, const float* scale, const float* bias, float* rm,
			  float* rv, float* sm, float* siv, double eps, double factor, void* stream, void* z = nullptr, bool* fused = nullptr)
{
	constexpr int V = 16 / sizeof(T);
	cudaStream_t s = pz_stream(stream);
	// one cluster per channel, the tensor crosses HBM once (see "cluster kernels").  In-place (y == x) is safe here: a CTA
	// only rewrites planes it alone reads, and the pivot element is read by every CTA of the cluster before its barrier
	const ClusterPlan cp = make_cluster_plan({x, y}, N, C, S, sizeof(T), 1);
	const bool fold = cp.ok && z != nullptr && !cp.g.bulk_store && (uintptr_t)z % 16 == 0;      // the ReLU output rides on the store pass
	PzProfScope prof(PZ_PROF_BN_FWD, s, 0.0, (fold ? 3.0 : 2.0) * (double)N * C * S * sizeof(T));
	{
		if (cp.ok) {
			// persistent variant: two stash buffers per CTA, one CTA per SM, as many clusters as the GPU can hold at once
			// (measured SLOWER than the one-shot kernel on every ResNet-50 shape -- 64 x 256 x 55 x 55: 0.149 vs 0.129 ms, and much slower on
			// small maps where a cluster's serial chain per channel dominates: profiles/r02_conv_epilogue_gather_experiments.md section 6;
			// PZ_BN_PERSISTENT=1 enables it for experiments)
			static const bool persistent_on = env_int("PZ_BN_PERSISTENT", 0) != 0;
			if (persistent_on && !fold && z == nullptr && cp.bulk && 2 * cp.smem <= 220 * 1024) {
				ClusterPlan pp = cp;
				pp.smem = 2 * cp.smem;
				pp.threads = 512;
				const int resident = resident_clusters(bn_fwd_persistent_kernel<T, V, 512>, pp);
				const int nclusters = resident < (int)C ? resident : (int)C;
				if (nclusters >= 1 && C >= 2 * nclusters) {
					pp.grid_override = nclusters * cp.g.CL;
					return launch_cluster(bn_fwd_persistent_kernel<T, V, 512>, pp, s, (const T*)x, (T*)y, cp.g, nclusters, scale, bias, rm, rv, sm, siv,
										  (float)eps, (float)factor);
				}
			}
			// the fused ReLU output rides on the plain-store phase 2 of the cluster kernel (16-byte aligned like x and y)
			T* zz = fold ? (T*)z : nullptr;
			if (fused) *fused = zz != nullptr;
#define PZ_BN_FWD_CLUSTER(TH, BK) launch_cluster(bn_fwd_cluster_kernel<T, V, TH, BK>, cp, s, (const T*)x, (T*)y, zz, cp.g, scale, bias, rm, rv, sm, siv, (float)eps, (float)factor)
			if (cp.threads == 512) return cp.bulk ? PZ_BN_FWD_CLUSTER(512, true) : PZ_BN_FWD_CLUSTER(512, false);
			return cp.bulk ? PZ_BN_FWD_CLUSTER(256, true) : PZ_BN_FWD_CLUSTER(256, false);
#undef PZ_BN_FWD_CLUSTER
		}
	}
	SumsPair sp;
	if (!acquire_sums(C, sp)) { pz_set_error(PZ_ERR_MEMORY, "batchnorm: cannot allocate the statistics buffer"); return PZ_ERR_MEMORY; }
	float* sums = sp.use;
	const Plan plan = make_plan<T>({x, y}, N, C, S, 1);
	const float count = (float)(N * S);
	for (int64_t c0 = 0; c0 < C; c0 += plan.slab_channels) {
		const int64_t c1 = c0 + plan.slab_channels < C ? c0 + plan.slab_channels : C;
		const Slab g = make_slab(plan, N, C, S, c0, c1);
		if (plan.vec == V) {
			PZ_BN_LAUNCH(V, (bn_stats_kernel<T, V>), bins_per_block(V, g.S) * 2 * sizeof(float), (const T*)x, g, sums, sp.pivots);
			PZ_BN_LAUNCH(V, (bn_apply_kernel<T, V, true>), 0, (const T*)x, (T*)y, g, sums, sp.pivots, scale, bias, rm, rv, sm, siv, (float)eps,
						 (float)factor, count, sp.clear, c0 == 0 ? sp.clear_n : 0);
		} else {
			PZ_BN_LAUNCH(1, (bn_stats_kernel<T, 1>), bins_per_block(1, g.S) * 2 * sizeof(float), (const T*)x, g, sums, sp.pivots);
			PZ_BN_LAUNCH(1, (bn_apply_kernel<T, 1, true>), 0, (const T*)x, (T*)y, g, sums, sp.pivots, scale, bias, rm, rv, sm, siv, (float)eps,
						 (float)factor, count, sp.clear, c0 == 0 ? sp.clear_n : 0);
		}
	}
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

template <typename T>
int fwd_infer(const void* x, void* y, int64_t N, int64_t C, int64_t S, const float* scale, const float* bias, const float* mean,
			  const float* var, double eps, void* stream)
{
	constexpr int V = 16 / sizeof(T);
	cudaStream_t s = pz_stream(stream);
	Plan plan = make_plan<T>({x, y}, N, C, S, 1);
	plan.slab_channels = C;                       // nothing is re-read: one launch over the whole tensor
	const Slab g = make_slab(plan, N, C, S, 0, C);
	PzProfScope prof(PZ_PROF_BN_FWD, s, 0.0, 2.0 * (double)N * C * S * sizeof(T));
	if (plan.vec == V)
		PZ_BN_LAUNCH(V, (bn_apply_kernel<T, V, false>), 0, (const T*)x, (T*)y, g, nullptr, nullptr, scale, bias, (float*)mean, (float*)var, nullptr,
					 nullptr, (float)eps, 0.0f, 1.0f, nullptr, 0);
	else
		PZ_BN_LAUNCH(1, (bn_apply_kernel<T, 1, false>), 0, (const T*)x, (T*)y, g, nullptr, nullptr, scale, bias, (float*)mean, (float*)var, nullptr,
					 nullptr, (float)eps, 0.0f, 1.0f, nullptr, 0);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

template <typename T>
int bwd(const void* x, const void* dy, void* dx, int64_t N, int64_t C, int64_t S, const float* scale, const float* sm,
		const float* siv, float* dscale, float* dbias, const BnAcc& acc, void* stream)
{
	constexpr int V = 16 / sizeof(T);
	cudaStream_t s = pz_stream(stream);
	PzProfScope prof(PZ_PROF_BN_BWD, s, 0.0, 3.0 * (double)N * C * S * sizeof(T));
	{
		const ClusterPlan cp = make_cluster_plan({x, dy, dx}, N, C, S, sizeof(T), 2);
		if (cp.ok) {
#define PZ_BN_BWD_CLUSTER(TH, BK) launch_cluster(bn_bwd_cluster_kernel<T, V, TH, BK>, cp, s, (const T*)x, (const T*)dy, (T*)dx, cp.g, scale, sm, siv, dscale, dbias, acc)
			if (cp.threads == 512) return cp.bulk ? PZ_BN_BWD_CLUSTER(512, true) : PZ_BN_BWD_CLUSTER(512, false);
			return cp.bulk ? PZ_BN_BWD_CLUSTER(256, true) : PZ_BN_BWD_CLUSTER(256, false);
#undef PZ_BN_BWD_CLUSTER
		}
	}
	SumsPair sp;
	if (!acquire_sums(C, sp)) { pz_set_error(PZ_ERR_MEMORY, "batchnorm: cannot allocate the statistics buffer"); return PZ_ERR_MEMORY; }
	float* sums = sp.use;
	const Plan plan = make_plan<T>({x, dy, dx}, N, C, S, 2);
	const float count = (float)(N * S);
	for (int64_t c0 = 0; c0 < C; c0 += plan.slab_channels) {
		const int64_t c1 = c0 + plan.slab_channels < C ? c0 + plan.slab_channels : C;
		const Slab g = make_slab(plan, N, C, S, c0, c1);
		if (plan.vec == V) {
			PZ_BN_LAUNCH(V, (bn_bwd_stats_kernel<T, V>), bins_per_block(V, g.S) * 2 * sizeof(float), (const T*)x, (const T*)dy, g, sm, sums);
			PZ_BN_LAUNCH(V, (bn_bwd_apply_kernel<T, V>), 0, (const T*)x, (const T*)dy, (T*)dx, g, sums, scale, sm, siv, dscale, dbias, count, sp.clear, c0 == 0 ? sp.clear_n : 0, acc);
		} else {
			PZ_BN_LAUNCH(1, (bn_bwd_stats_kernel<T, 1>), bins_per_block(1, g.S) * 2 * sizeof(float), (const T*)x, (const T*)dy, g, sm, sums);
			PZ_BN_LAUNCH(1, (bn_bwd_apply_kernel<T, 1>), 0, (const T*)x, (const T*)dy, (T*)dx, g, sums, scale, sm, siv, dscale, dbias, count, sp.clear, c0 == 0 ? sp.clear_n : 0, acc);
		}
	}
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

int check_dims(int64_t N, int64_t C, int64_t S)
{
	PZ_REQUIRE(N > 0 && C > 0 && S > 0, "batchnorm: empty tensor");
	PZ_REQUIRE(N < (1ll << 31) && S < (1ll << 31) && C * S < (1ll << 31) - 4096 && pz_cdiv(C * S, kThreads) * pz_cdiv(N, kRowUnroll) < (1ll << 31),
			   "batchnorm: tensor too large (N=%lld C=%lld S=%lld)", (long long)N, (long long)C, (long long)S);
	return PZ_OK;
}

}  // namespace

// A captured graph replays the same accumulator choreography every time, so it must start from a known state: both buffers
// cleared (as the first nodes of the graph) and the alternation reset.
void pz_norm_graph_begin(cudaStream_t stream)
{
	std::lock_guard<std::mutex> lock(g_sums_mu);
	for (int i = 0; i < 2; i++) {
		if (g_sums[i]) cudaMemsetAsync(g_sums[i], 0, (size_t)g_sums_cap * 2 * sizeof(float), stream);
		g_sums_dirty[i] = 0;
	}
	g_sums_cur = 0;
}

extern "C" {

int pz_bn_fwd_train(int dtype, const void* x, void* y, int64_t N, int64_t C, int64_t S, const float* scale, const float* bias,
					float* running_mean, float* running_var, float* save_mean, float* save_invvar, double eps, double factor,
					void* stream)
{
	int st = check_dims(N, C, S);
	if (st != PZ_OK) return st;
	PZ_DISPATCH_FLOAT(dtype, fwd_train<T>(x, y, N, C, S, scale, bias, running_mean, running_var, save_mean, save_invvar, eps, factor, stream));
}

int pz_bn_fwd_train_relu(int dtype, const void* x, void* y, void* z, int64_t N, int64_t C, int64_t S, const float* scale, const float* bias,
						 float* running_mean, float* running_var, float* save_mean, float* save_invvar, double eps, double factor,
						 void* stream)
{
	int st = check_dims(N, C, S);
	if (st != PZ_OK) return st;
	bool fused = false;
	switch (dtype) {
		case PZ_F32: st = fwd_train<float>(x, y, N, C, S, scale, bias, running_mean, running_var, save_mean, save_invvar, eps, factor, stream, z, &fused); break;
		case PZ_F16: st = fwd_train<__half>(x, y, N, C, S, scale, bias, running_mean, running_var, save_mean, save_invvar, eps, factor, stream, z, &fused); break;
		case PZ_BF16: st = fwd_train<__nv_bfloat16>(x, y, N, C, S, scale, bias, running_mean, running_var, save_mean, save_invvar, eps, factor, stream, z, &fused); break;
		default: pz_set_error(PZ_ERR_UNSUPPORTED, "unsupported dtype %d", dtype); return PZ_ERR_UNSUPPORTED;
	}
	if (st != PZ_OK || fused) return st;
	return pz_act_fwd(PZ_ACT_RELU, dtype, z, y, N * C * S, 0.0f, 0.0f, stream);      // paths without the fused store: the plain ReLU kernel
}

int pz_bn_fwd_infer(int dtype, const void* x, void* y, int64_t N, int64_t C, int64_t S, const float* scale, const float* bias,
					const float* mean, const float* var, double eps, void* stream)
{
	int st = check_dims(N, C, S);
	if (st != PZ_OK) return st;
	PZ_DISPATCH_FLOAT(dtype, fwd_infer<T>(x, y, N, C, S, scale, bias, mean, var, eps, stream));
}

int pz_bn_bwd(int dtype, const void* x, const void* dy, void* dx, int64_t N, int64_t C, int64_t S, const float* scale,
			  const float* save_mean, const float* save_invvar, float* dscale, float* dbias, void* stream)
{
	int st = check_dims(N, C, S);
	if (st != PZ_OK) return st;
	const BnAcc none{nullptr, nullptr, 0.0f, 0.0f, 0.0f, 0.0f};
	PZ_DISPATCH_FLOAT(dtype, bwd<T>(x, dy, dx, N, C, S, scale, save_mean, save_invvar, dscale, dbias, none, stream));
}

// pz_bn_bwd followed by  scale_acc = scale_alpha * dscale + scale_beta * scale_acc  and the same for the bias (either may be NULL):
// what BatchNormND.accGradParams (Modules/BatchNormND.py:74-83) does with two addVectorToVector launches right after the pass.
// The accumulators must not overlap any other tensor of the call.
int pz_bn_bwd_acc(int dtype, const void* x, const void* dy, void* dx, int64_t N, int64_t C, int64_t S, const float* scale,
				  const float* save_mean, const float* save_invvar, float* dscale, float* dbias, float* scale_acc, float scale_alpha,
				  float scale_beta, float* bias_acc, float bias_alpha, float bias_beta, void* stream)
{
	int st = check_dims(N, C, S);
	if (st != PZ_OK) return st;
	const BnAcc acc{scale_acc, bias_acc, scale_alpha, scale_beta, bias_alpha, bias_beta};
	PZ_DISPATCH_FLOAT(dtype, bwd<T>(x, dy, dx, N, C, S, scale, save_mean, save_invvar, dscale, dbias, acc, stream));
}

}  // extern "C"
