// pz_norm.cu -- batch normalisation forward (train / inference) and backward, HBM-bandwidth bound.
//
// Replaces cudnnBatchNormalizationForwardTraining / ForwardInference / Backward as called from
// CuDnn_Context_batchNormNd / _batchNormNdBackward (reference Cuda/Source/Libs/CuDnnNorm.c:31-71,158-194).
//
// B200 design: one thread-block CLUSTER per channel.  The N planes of a channel are dealt out to the CTAs
// of the cluster; each CTA reduces its planes with 128-bit loads + warp shuffles, the per-CTA partials are
// combined through distributed shared memory (every CTA reads all partials in rank order, so all CTAs hold
// the same bits and no global atomics / second launch are needed), and the same CTA immediately re-reads
// its planes - still resident in the 126 MB L2 - to write the normalised output.  DRAM traffic is therefore
// the algorithmic one: read x once + write y once (forward), read x, dy once + write dx once (backward).
#include "pz_common.h"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 512;
constexpr int kMaxCluster = 8;

template <typename T> struct VecOf;
template <> struct VecOf<float> { static constexpr int N = 4; };
template <> struct VecOf<__half> { static constexpr int N = 8; };
template <> struct VecOf<__nv_bfloat16> { static constexpr int N = 8; };

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T> struct alignas(16) Pack { T v[16 / sizeof(T)]; };

// Visit every element of a plane of S elements with a group of `gsize` threads (rank `gi` in the group).
// body(i, vec) is called for 16-byte aligned packs, tail(i) for the unaligned head / tail elements.
template <typename T, typename FV, typename FS>
__device__ __forceinline__ void plane_visit(const T* plane, int64_t S, int gi, int gsize, bool vec_ok, FV body, FS tail)
{
	constexpr int V = VecOf<T>::N;
	int64_t head = 0, nvec = 0;
	if (vec_ok) {
		uintptr_t mis = ((uintptr_t)plane % 16) / sizeof(T);
		head = mis ? (int64_t)(V - mis) : 0;
		if (head > S) head = S;
		nvec = (S - head) / V;
	}
	#pragma unroll 4
	for (int64_t v = gi; v < nvec; v += gsize) body(head + v * V);      // unrolled: 4 independent 128-bit loads in flight
	const int64_t tailstart = head + nvec * V;
	const int64_t nscalar = head + (S - tailstart);
	for (int64_t s = gi; s < nscalar; s += gsize) tail(s < head ? s : tailstart + (s - head));
}

// ---- Chan et al. parallel (count, mean, M2) combination
struct Moments { float n, mean, m2; };

__device__ __forceinline__ Moments combine(const Moments& a, const Moments& b)
{
	Moments r;
	r.n = a.n + b.n;
	if (r.n == 0.0f) { r.mean = 0.0f; r.m2 = 0.0f; return r; }
	float delta = b.mean - a.mean;
	float frac = b.n / r.n;
	r.mean = a.mean + delta * frac;
	r.m2 = a.m2 + b.m2 + delta * delta * a.n * frac;
	return r;
}

__device__ __forceinline__ Moments warp_combine(Moments m)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		Moments other;
		other.n = __shfl_xor_sync(0xffffffffu, m.n, o);
		other.mean = __shfl_xor_sync(0xffffffffu, m.mean, o);
		other.m2 = __shfl_xor_sync(0xffffffffu, m.m2, o);
		// order the pair by lane so that both lanes compute bit-identical results
		bool lower = (threadIdx.x & o) == 0;
		m = lower ? combine(m, other) : combine(other, m);
	}
	return m;
}

__device__ __forceinline__ float warp_sum(float v)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

struct PlaneSplit {
	int64_t first, count;  // this CTA handles planes n = first, first + 1, ... (count of them) of its channel
};

__device__ __forceinline__ PlaneSplit split_planes(int64_t N, unsigned rank, unsigned csize)
{
	int64_t base = N / csize, rem = N % csize;
	PlaneSplit p;
	p.first = rank * base + (rank < rem ? rank : rem);
	p.count = base + (rank < rem ? 1 : 0);
	return p;
}

// ------------------------------------------------------------------------------------------ forward, training
template <typename T>
__global__ void __launch_bounds__(kThreads) bn_fwd_train_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t N, int64_t C,
																int64_t S, const float* __restrict__ scale,
																const float* __restrict__ bias, float* running_mean,
																float* running_var, float* save_mean, float* save_invvar,
																float eps, float factor, int vec_ok, int warp_planes)
{
	cg::cluster_group cluster = cg::this_cluster();
	const unsigned rank = cluster.block_rank(), csize = cluster.num_blocks();
	const int64_t c = blockIdx.y;
	const PlaneSplit ps = split_planes(N, rank, csize);

	const int gsize = warp_planes ? 32 : kThreads;
	const int gi = warp_planes ? (threadIdx.x & 31) : threadIdx.x;
	const int group = warp_planes ? (threadIdx.x >> 5) : 0;
	const int ngroups = warp_planes ? kThreads / 32 : 1;

	__shared__ Moments warp_part[kThreads / 32];
	__shared__ Moments cta_part;       // read by the other CTAs of the cluster through DSMEM
	__shared__ float coef[2];

	// pass 1: sums of d = x - pivot and d^2 with a per-channel pivot (the channel's first element, identical for
	// every thread of the cluster), so that E[d^2] - E[d]^2 does not cancel when |mean| >> std; plain sums combine
	// associatively in a fixed order -> bit-identical statistics in every CTA of the cluster
	const float pivot = to_f<T>(x[c * S]);
	float s1 = 0.0f, s2 = 0.0f;
	for (int64_t pl = group; pl < ps.count; pl += ngroups) {
		const T* plane = x + ((ps.first + pl) * C + c) * S;
		plane_visit<T>(plane, S, gi, gsize, vec_ok,
			[&](int64_t i) {
				Pack<T> p = *reinterpret_cast<const Pack<T>*>(plane + i);
				#pragma unroll
				for (int e = 0; e < VecOf<T>::N; e++) { float d = to_f<T>(p.v[e]) - pivot; s1 += d; s2 = fmaf(d, d, s2); }
			},
			[&](int64_t i) { float d = to_f<T>(plane[i]) - pivot; s1 += d; s2 = fmaf(d, d, s2); });
	}
	s1 = warp_sum(s1);
	s2 = warp_sum(s2);
	if ((threadIdx.x & 31) == 0) { warp_part[threadIdx.x >> 5].mean = s1; warp_part[threadIdx.x >> 5].m2 = s2; }
	__syncthreads();
	if (threadIdx.x == 0) {
		float a = 0.0f, b = 0.0f;
		for (int w = 0; w < kThreads / 32; w++) { a += warp_part[w].mean; b += warp_part[w].m2; }
		cta_part.mean = a;
		cta_part.m2 = b;
	}
	cluster.sync();
	if (threadIdx.x == 0) {
		float t1 = 0.0f, t2 = 0.0f;
		for (unsigned r = 0; r < csize; r++) {
			const Moments* remote = cluster.map_shared_rank(&cta_part, r);
			t1 += remote->mean;
			t2 += remote->m2;
		}
		const float cnt = (float)(N * S);
		const float dmean = t1 / cnt;
		const float mean = pivot + dmean;
		const float var = fmaxf(t2 / cnt - dmean * dmean, 0.0f);            // biased, used for normalisation
		const float invstd = 1.0f / sqrtf(var + eps);
		const float a = scale[c] * invstd;
		coef[0] = a;
		coef[1] = bias[c] - mean * a;
		if (rank == 0) {
			save_mean[c] = mean;
			save_invvar[c] = invstd;
			// cuDNN keeps the UNBIASED variance in the running estimate
			const float uvar = cnt > 1.0f ? var * (cnt / (cnt - 1.0f)) : var;
			running_mean[c] = (1.0f - factor) * running_mean[c] + factor * mean;
			running_var[c] = (1.0f - factor) * running_var[c] + factor * uvar;
		}
	}
	cluster.sync();   // also keeps every cta_part alive until all remote reads are done
	const float a = coef[0], b = coef[1];

	// pass 2: the planes were just read by this CTA and are still in L2
	for (int64_t pl = group; pl < ps.count; pl += ngroups) {
		const int64_t off = ((ps.first + pl) * C + c) * S;
		const T* plane = x + off;
		T* out = y + off;
		plane_visit<T>(plane, S, gi, gsize, vec_ok,
			[&](int64_t i) {
				Pack<T> p = *reinterpret_cast<const Pack<T>*>(plane + i);
				#pragma unroll
				for (int e = 0; e < VecOf<T>::N; e++) p.v[e] = from_f<T>(fmaf(to_f<T>(p.v[e]), a, b));
				*reinterpret_cast<Pack<T>*>(out + i) = p;
			},
			[&](int64_t i) { out[i] = from_f<T>(fmaf(to_f<T>(plane[i]), a, b)); });
	}
}

// ------------------------------------------------------------------------------------------ forward, inference
template <typename T>
__global__ void __launch_bounds__(256) bn_fwd_infer_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t planes, int64_t C,
														   int64_t S, const float* __restrict__ scale,
														   const float* __restrict__ bias, const float* __restrict__ mean,
														   const float* __restrict__ var, float eps, int vec_ok)
{
	for (int64_t pl = blockIdx.x; pl < planes; pl += gridDim.x) {
		const int64_t c = pl % C;
		const float a = scale[c] / sqrtf(var[c] + eps);
		const float b = bias[c] - mean[c] * a;
		const T* plane = x + pl * S;
		T* out = y + pl * S;
		plane_visit<T>(plane, S, threadIdx.x, 256, vec_ok,
			[&](int64_t i) {
				Pack<T> p = *reinterpret_cast<const Pack<T>*>(plane + i);
				#pragma unroll
				for (int e = 0; e < VecOf<T>::N; e++) p.v[e] = from_f<T>(fmaf(to_f<T>(p.v[e]), a, b));
				*reinterpret_cast<Pack<T>*>(out + i) = p;
			},
			[&](int64_t i) { out[i] = from_f<T>(fmaf(to_f<T>(plane[i]), a, b)); });
	}
}

// ------------------------------------------------------------------------------------------ backward
template <typename T>
__global__ void __launch_bounds__(kThreads) bn_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx,
														  int64_t N, int64_t C, int64_t S, const float* __restrict__ scale,
														  const float* __restrict__ save_mean,
														  const float* __restrict__ save_invvar, float* dscale, float* dbias,
														  int vec_ok, int warp_planes)
{
	cg::cluster_group cluster = cg::this_cluster();
	const unsigned rank = cluster.block_rank(), csize = cluster.num_blocks();
	const int64_t c = blockIdx.y;
	const PlaneSplit ps = split_planes(N, rank, csize);

	const int gsize = warp_planes ? 32 : kThreads;
	const int gi = warp_planes ? (threadIdx.x & 31) : threadIdx.x;
	const int group = warp_planes ? (threadIdx.x >> 5) : 0;
	const int ngroups = warp_planes ? kThreads / 32 : 1;

	__shared__ float warp_part[2][kThreads / 32];
	__shared__ float cta_part[2];
	__shared__ float coef[3];

	const float mean = save_mean[c], invstd = save_invvar[c];

	// pass 1: sum(dy) and sum(dy * xhat)
	float sdy = 0.0f, sdyx = 0.0f;
	for (int64_t pl = group; pl < ps.count; pl += ngroups) {
		const int64_t off = ((ps.first + pl) * C + c) * S;
		const T* px = x + off;
		const T* pg = dy + off;
		plane_visit<T>(px, S, gi, gsize, vec_ok,
			[&](int64_t i) {
				Pack<T> a = *reinterpret_cast<const Pack<T>*>(px + i);
				Pack<T> g = *reinterpret_cast<const Pack<T>*>(pg + i);
				#pragma unroll
				for (int e = 0; e < VecOf<T>::N; e++) {
					float gv = to_f<T>(g.v[e]);
					sdy += gv;
					sdyx += gv * (to_f<T>(a.v[e]) - mean);
				}
			},
			[&](int64_t i) { float gv = to_f<T>(pg[i]); sdy += gv; sdyx += gv * (to_f<T>(px[i]) - mean); });
	}
	sdy = warp_sum(sdy);
	sdyx = warp_sum(sdyx);
	if ((threadIdx.x & 31) == 0) { warp_part[0][threadIdx.x >> 5] = sdy; warp_part[1][threadIdx.x >> 5] = sdyx; }
	__syncthreads();
	if (threadIdx.x == 0) {
		float a = 0.0f, b = 0.0f;
		for (int w = 0; w < kThreads / 32; w++) { a += warp_part[0][w]; b += warp_part[1][w]; }
		cta_part[0] = a;
		cta_part[1] = b;
	}
	cluster.sync();
	if (threadIdx.x == 0) {
		float tdy = 0.0f, tdyx = 0.0f;
		for (unsigned r = 0; r < csize; r++) {
			const float* remote = cluster.map_shared_rank(&cta_part[0], r);
			tdy += remote[0];
			tdyx += remote[1];
		}
		const float dsc = tdyx * invstd;        // sum(dy * xhat)
		const float m = (float)(N * S);
		const float c1 = scale[c] * invstd;
		coef[0] = c1;
		coef[1] = c1 * tdy / m;
		coef[2] = c1 * dsc / m * invstd;        // multiplies (x - mean)
		if (rank == 0) { dscale[c] = dsc; dbias[c] = tdy; }
	}
	cluster.sync();
	const float c1 = coef[0], c2 = coef[1], c3 = coef[2];

	// pass 2: dx = c1*dy - c2 - (x - mean)*c3
	for (int64_t pl = group; pl < ps.count; pl += ngroups) {
		const int64_t off = ((ps.first + pl) * C + c) * S;
		const T* px = x + off;
		const T* pg = dy + off;
		T* pd = dx + off;
		plane_visit<T>(px, S, gi, gsize, vec_ok,
			[&](int64_t i) {
				Pack<T> a = *reinterpret_cast<const Pack<T>*>(px + i);
				Pack<T> g = *reinterpret_cast<const Pack<T>*>(pg + i);
				#pragma unroll
				for (int e = 0; e < VecOf<T>::N; e++)
					g.v[e] = from_f<T>(c1 * to_f<T>(g.v[e]) - c2 - (to_f<T>(a.v[e]) - mean) * c3);
				*reinterpret_cast<Pack<T>*>(pd + i) = g;
			},
			[&](int64_t i) { pd[i] = from_f<T>(c1 * to_f<T>(pg[i]) - c2 - (to_f<T>(px[i]) - mean) * c3); });
	}
}

int pick_cluster(int64_t N, int64_t C)
{
	int cs = 1;
	while (cs < kMaxCluster && C * cs < 2ll * pz_num_sms() && cs * 2 <= N) cs *= 2;
	return cs;
}

template <typename K, typename... Args>
int launch_cluster(K kern, dim3 grid, int threads, int cs, cudaStream_t stream, Args... args)
{
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = grid;
	cfg.blockDim = dim3((unsigned)threads);
	cfg.dynamicSmemBytes = 0;
	cfg.stream = stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = (unsigned)cs;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	PZ_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
	pz_count_launch(1);
	return PZ_OK;
}

bool same_misalignment(std::initializer_list<const void*> ptrs)
{
	uintptr_t m = (uintptr_t)(*ptrs.begin()) % 16;
	for (const void* p : ptrs)
		if ((uintptr_t)p % 16 != m) return false;
	return true;
}

template <typename T>
int fwd_train(const void* x, void* y, int64_t N, int64_t C, int64_t S, const float* scale, const float* bias, float* rm,
			  float* rv, float* sm, float* siv, double eps, double factor, void* stream)
{
	const int cs = pick_cluster(N, C);
	const int vec_ok = same_misalignment({x, y}) && ((uintptr_t)x % sizeof(T) == 0);
	const int warp_planes = S < 2048;
	PzProfScope prof(PZ_PROF_BN_FWD, pz_stream(stream), 0.0, 2.0 * (double)N * C * S * sizeof(T));
	return launch_cluster(bn_fwd_train_kernel<T>, dim3((unsigned)cs, (unsigned)C), kThreads, cs, pz_stream(stream), (const T*)x,
						  (T*)y, N, C, S, scale, bias, rm, rv, sm, siv, (float)eps, (float)factor, vec_ok, warp_planes);
}

template <typename T>
int fwd_infer(const void* x, void* y, int64_t N, int64_t C, int64_t S, const float* scale, const float* bias, const float* mean,
			  const float* var, double eps, void* stream)
{
	const int64_t planes = N * C;
	const int vec_ok = same_misalignment({x, y}) && ((uintptr_t)x % sizeof(T) == 0);
	int64_t blocks = planes < (int64_t)pz_num_sms() * 8 ? planes : (int64_t)pz_num_sms() * 8;
	bn_fwd_infer_kernel<T><<<(unsigned)blocks, 256, 0, pz_stream(stream)>>>((const T*)x, (T*)y, planes, C, S, scale, bias, mean, var,
																			(float)eps, vec_ok);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

template <typename T>
int bwd(const void* x, const void* dy, void* dx, int64_t N, int64_t C, int64_t S, const float* scale, const float* sm,
		const float* siv, float* dscale, float* dbias, void* stream)
{
	const int cs = pick_cluster(N, C);
	const int vec_ok = same_misalignment({x, dy, dx}) && ((uintptr_t)x % sizeof(T) == 0);
	const int warp_planes = S < 2048;
	PzProfScope prof(PZ_PROF_BN_BWD, pz_stream(stream), 0.0, 3.0 * (double)N * C * S * sizeof(T));
	return launch_cluster(bn_bwd_kernel<T>, dim3((unsigned)cs, (unsigned)C), kThreads, cs, pz_stream(stream), (const T*)x,
						  (const T*)dy, (T*)dx, N, C, S, scale, sm, siv, dscale, dbias, vec_ok, warp_planes);
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

int pz_bn_fwd_train(int dtype, const void* x, void* y, int64_t N, int64_t C, int64_t S, const float* scale, const float* bias,
					float* running_mean, float* running_var, float* save_mean, float* save_invvar, double eps, double factor,
					void* stream)
{
	PZ_REQUIRE(N > 0 && C > 0 && S > 0, "batchnorm: empty tensor");
	PZ_REQUIRE(C <= 65535, "batchnorm: too many channels for one launch (%lld)", (long long)C);
	PZ_DISPATCH_FLOAT(dtype, fwd_train<T>(x, y, N, C, S, scale, bias, running_mean, running_var, save_mean, save_invvar, eps, factor, stream));
}

int pz_bn_fwd_infer(int dtype, const void* x, void* y, int64_t N, int64_t C, int64_t S, const float* scale, const float* bias,
					const float* mean, const float* var, double eps, void* stream)
{
	PZ_REQUIRE(N > 0 && C > 0 && S > 0, "batchnorm: empty tensor");
	PZ_DISPATCH_FLOAT(dtype, fwd_infer<T>(x, y, N, C, S, scale, bias, mean, var, eps, stream));
}

int pz_bn_bwd(int dtype, const void* x, const void* dy, void* dx, int64_t N, int64_t C, int64_t S, const float* scale,
			  const float* save_mean, const float* save_invvar, float* dscale, float* dbias, void* stream)
{
	PZ_REQUIRE(N > 0 && C > 0 && S > 0, "batchnorm: empty tensor");
	PZ_REQUIRE(C <= 65535, "batchnorm: too many channels for one launch (%lld)", (long long)C);
	PZ_DISPATCH_FLOAT(dtype, bwd<T>(x, dy, dx, N, C, S, scale, save_mean, save_invvar, dscale, dbias, stream));
}

}  // extern "C"
