// Training closure either side of the hot path: the cross-entropy cost that produces the top gradient,
// the label-mismatch count behind the validation accuracy, and the Nesterov / Adam parameter updates.  All bandwidth-bound,
// one pass over their tensors.
#include "pz_common.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace {

constexpr int kThreads = 256;

template <typename T> __device__ __forceinline__ float to_f(T v) { return (float)v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v) { return (T)v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float block_sum(float v, float* red)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
	__syncthreads();
	float s = 0.0f;
	if (threadIdx.x < 32) {
		s = threadIdx.x < kThreads / 32 ? red[threadIdx.x] : 0.0f;
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
	}
	return s;      // valid in thread 0
}

// reference: Cuda/Kernels/Costs.py:77-106 (cost template + crossEntropyLogic), :133-157 (weighted variant).
// index = (b, c, m) over (numSamples, numCases, spatialDim); grad = w_c * ((c == label) - p) / numSamples;
// error += -w_c * log(p) / spatialDim where c == label.  One atomic per block instead of one per sample.
__global__ void __launch_bounds__(kThreads) cross_entropy_kernel(const float* __restrict__ probs, const int* __restrict__ labels,
																const float* __restrict__ weights, long long size, int mapStride,
																int spatialDim, int numCases, float invSamples, float invSpatial,
																float* __restrict__ error, float* __restrict__ grad)
{
	__shared__ float red[kThreads / 32];
	float err = 0.0f;
	for (long long index = (long long)blockIdx.x * kThreads + threadIdx.x; index < size; index += (long long)gridDim.x * kThreads) {
		const int b = (int)(index / mapStride);
		const int m = (int)(index % spatialDim);
		const int c = (int)((index / spatialDim) % numCases);
		const float score = probs[index];
		const int label = labels[(long long)b * spatialDim + m];
		const float w = weights ? weights[c] : 1.0f;
		grad[index] = w * ((c == label ? 1.0f : 0.0f) - score) * invSamples;
		if (c == label) err += -w * logf(score) * invSpatial;
	}
	const float s = block_sum(err, red);
	if (threadIdx.x == 0 && s != 0.0f) atomicAdd(error, s);
}

// reference: Cuda/Kernels/Costs.py:109-130 (svmL1Logic / svmL2Logic in the same cost template): cls = +1 for the labelled
// class, -1 otherwise; l1: hinge, l2: squared hinge.  error is summed over every (sample, class, position).
__global__ void __launch_bounds__(kThreads) svm_kernel(const float* __restrict__ scores, const int* __restrict__ labels, long long size,
													  int mapStride, int spatialDim, int numCases, int numSamples, int l2,
													  float* __restrict__ error, float* __restrict__ grad)
{
	__shared__ float red[kThreads / 32];
	float err = 0.0f;
	for (long long index = (long long)blockIdx.x * kThreads + threadIdx.x; index < size; index += (long long)gridDim.x * kThreads) {
		const int b = (int)(index / mapStride);
		const int m = (int)(index % spatialDim);
		const int c = (int)((index / spatialDim) % numCases);
		const float score = scores[index];
		const int label = labels[(long long)b * spatialDim + m];
		const float cls = label == c ? 1.0f : -1.0f;
		if (l2) {
			const float e = fmaxf(0.0f, 1.0f - score * cls);
			grad[index] = 2.0f * cls * e / numCases / numSamples;
			err += e * e / numCases / spatialDim;
		} else {
			grad[index] = score * cls < 1.0f ? cls / numCases / numSamples : 0.0f;
			err += fmaxf(0.0f, 1.0f - score * cls) / numCases / spatialDim;
		}
	}
	const float s = block_sum(err, red);
	if (threadIdx.x == 0 && s != 0.0f) atomicAdd(error, s);
}

// reference: Cuda/Kernels/Costs.py:178-182 (calcAccuracy reduction: sum of x[i] != y[i] as float)
__global__ void __launch_bounds__(kThreads) mismatch_kernel(const int* __restrict__ x, const int* __restrict__ y, long long n,
														   float* __restrict__ out)
{
	__shared__ float red[kThreads / 32];
	float cnt = 0.0f;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads)
		cnt += x[i] != y[i] ? 1.0f : 0.0f;
	const float s = block_sum(cnt, red);
	if (threadIdx.x == 0 && s != 0.0f) atomicAdd(out, s);
}

// reference: Cuda/Kernels/ElementWise.py:815-857 (nesterovMomSGDKer); the parameter update uses the OLD momentum
template <typename T>
__global__ void __launch_bounds__(kThreads) nesterov_kernel(T* __restrict__ param, const T* __restrict__ grad, T* __restrict__ mom,
														   float lr, float mr, long long n)
{
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
		const float g = to_f(grad[i]), m = to_f(mom[i]);
		param[i] = from_f<T>(to_f(param[i]) + mr * mr * m + (1.0f + mr) * lr * g);
		mom[i] = from_f<T>(mr * m + lr * g);
	}
}

// reference: Cuda/Kernels/ElementWise.py:709-755 (adamKer): fp32 first / second moments whatever the parameter type
template <typename T>
__global__ void __launch_bounds__(kThreads) adam_kernel(T* __restrict__ param, const T* __restrict__ grad, float* __restrict__ mg,
													   float* __restrict__ ms, float lr, float fix1, float fix2, float eps, long long n)
{
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
		const float g = to_f(grad[i]);
		float a = mg[i], s = ms[i];
		a += fix1 * (g - a);
		s += fix2 * (g * g - s);
		param[i] = from_f<T>(to_f(param[i]) + lr * a / (sqrtf(s) + eps));
		mg[i] = a;
		ms[i] = s;
	}
}

unsigned grid_for(long long n)
{
	long long blocks = pz_cdiv(n, (long long)kThreads);
	const long long cap = (long long)pz_num_sms() * 8;
	if (blocks > cap) blocks = cap;
	return (unsigned)(blocks < 1 ? 1 : blocks);
}

}  // namespace

extern "C" {

int pz_cross_entropy(const void* probs, const void* labels, const void* weights, int64_t samples, int64_t cases, int64_t spatial,
					 void* error, void* grad, void* stream)
{
	PZ_REQUIRE(samples >= 0 && cases > 0 && spatial > 0, "cross entropy: bad shape");
	const long long size = (long long)samples * cases * spatial;
	if (size == 0) return PZ_OK;
	PZ_REQUIRE(cases * spatial < (1ll << 31) && samples * spatial < (1ll << 31), "cross entropy: tensor too large");
	PzProfScope prof(PZ_PROF_ELTWISE, pz_stream(stream), 0.0, 8.0 * (double)size);
	cross_entropy_kernel<<<grid_for(size), kThreads, 0, pz_stream(stream)>>>(
		(const float*)probs, (const int*)labels, (const float*)weights, size, (int)(cases * spatial), (int)spatial, (int)cases,
		1.0f / (float)samples, 1.0f / (float)spatial, (float*)error, (float*)grad);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

int pz_svm(int l2, const void* scores, const void* labels, int64_t samples, int64_t cases, int64_t spatial, void* error, void* grad,
		   void* stream)
{
	PZ_REQUIRE(samples >= 0 && cases > 0 && spatial > 0, "svm: bad shape");
	const long long size = (long long)samples * cases * spatial;
	if (size == 0) return PZ_OK;
	PZ_REQUIRE(cases * spatial < (1ll << 31) && samples * spatial < (1ll << 31), "svm: tensor too large");
	svm_kernel<<<grid_for(size), kThreads, 0, pz_stream(stream)>>>((const float*)scores, (const int*)labels, size, (int)(cases * spatial),
																  (int)spatial, (int)cases, (int)samples, l2 ? 1 : 0, (float*)error, (float*)grad);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

int pz_count_mismatch(const void* x, const void* y, int64_t n, void* out, void* stream)
{
	if (n <= 0) return PZ_OK;
	mismatch_kernel<<<grid_for(n), kThreads, 0, pz_stream(stream)>>>((const int*)x, (const int*)y, (long long)n, (float*)out);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

}  // extern "C"

template <typename T>
static int nesterov_launch(void* param, const void* grad, void* mom, float lr, float mr, int64_t n, void* stream)
{
	PzProfScope prof(PZ_PROF_ELTWISE, pz_stream(stream), 0.0, 5.0 * (double)n * sizeof(T));
	nesterov_kernel<T><<<grid_for(n), kThreads, 0, pz_stream(stream)>>>((T*)param, (const T*)grad, (T*)mom, lr, mr, (long long)n);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

template <typename T>
static int adam_launch(void* param, const void* grad, void* mg, void* ms, float lr, float fix1, float fix2, float eps, int64_t n,
					   void* stream)
{
	PzProfScope prof(PZ_PROF_ELTWISE, pz_stream(stream), 0.0, (double)n * (3.0 * sizeof(T) + 16.0));
	adam_kernel<T><<<grid_for(n), kThreads, 0, pz_stream(stream)>>>((T*)param, (const T*)grad, (float*)mg, (float*)ms, lr, fix1, fix2,
																	eps, (long long)n);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

extern "C" {

#define PZ_TRAIN_DISPATCH(dtype, ...)                                                    \
	switch (dtype) {                                                                     \
		case PZ_F32: { using T = float; return __VA_ARGS__; }                            \
		case PZ_F16: { using T = __half; return __VA_ARGS__; }                           \
		case PZ_BF16: { using T = __nv_bfloat16; return __VA_ARGS__; }                   \
		default: pz_set_error(PZ_ERR_UNSUPPORTED, "unsupported dtype %d", (int)(dtype)); \
				 return PZ_ERR_UNSUPPORTED;                                              \
	}

int pz_sgd_nesterov(int dtype, void* param, const void* grad, void* mom, float lr, float mr, int64_t n, void* stream)
{
	if (n <= 0) return PZ_OK;
	PZ_TRAIN_DISPATCH(dtype, nesterov_launch<T>(param, grad, mom, lr, mr, n, stream));
}

int pz_adam(int dtype, void* param, const void* grad, void* mg, void* ms, float lr, float fix1, float fix2, float eps, int64_t n,
			void* stream)
{
	if (n <= 0) return PZ_OK;
	PZ_TRAIN_DISPATCH(dtype, adam_launch<T>(param, grad, mg, ms, lr, fix1, fix2, eps, n, stream));
}

}  // extern "C"

// ------------------------------------------------------------------------------------------ axis permutation
// reference: Cuda/Source/Libs/CuDnnMemory.c (cudnnTransformTensor-based transpose / moveaxis / swapaxes).  One thread per
// output element: coalesced writes, gathered reads through the read-only path.  out[i0..ik] = in[sum_d i_d * stride_in[d]].
namespace {

constexpr int kMaxDims = 8;
struct PermuteGeo {
	int ndim;
	unsigned shape[kMaxDims];          // output shape
	long long stride[kMaxDims];        // input element stride of every output axis
};

template <typename T>
__global__ void __launch_bounds__(kThreads) permute_kernel(T* __restrict__ out, const T* __restrict__ in, PermuteGeo g, long long total)
{
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
		long long rem = i, off = 0;
		#pragma unroll
		for (int d = kMaxDims - 1; d >= 0; d--) {
			if (d < g.ndim) {
				const long long q = rem / g.shape[d];
				off += (rem - q * g.shape[d]) * g.stride[d];
				rem = q;
			}
		}
		out[i] = in[off];
	}
}

}  // namespace

extern "C" int pz_permute(int itemsize, void* out, const void* in, int ndim, const int64_t* out_shape, const int64_t* in_stride,
						  void* stream)
{
	PZ_REQUIRE(ndim >= 1 && ndim <= kMaxDims, "permute: between 1 and %d axes are supported (got %d)", kMaxDims, ndim);
	PermuteGeo g{};
	g.ndim = ndim;
	long long total = 1;
	for (int d = 0; d < ndim; d++) {
		PZ_REQUIRE(out_shape[d] >= 0 && out_shape[d] < (1ll << 32), "permute: bad extent");
		g.shape[d] = (unsigned)out_shape[d];
		g.stride[d] = in_stride[d];
		total *= out_shape[d];
	}
	if (total == 0) return PZ_OK;
	PzProfScope prof(PZ_PROF_ELTWISE, pz_stream(stream), 0.0, 2.0 * (double)total * itemsize);
	const unsigned grid = grid_for(total);
	switch (itemsize) {
		case 1: permute_kernel<uint8_t><<<grid, kThreads, 0, pz_stream(stream)>>>((uint8_t*)out, (const uint8_t*)in, g, total); break;
		case 2: permute_kernel<uint16_t><<<grid, kThreads, 0, pz_stream(stream)>>>((uint16_t*)out, (const uint16_t*)in, g, total); break;
		case 4: permute_kernel<uint32_t><<<grid, kThreads, 0, pz_stream(stream)>>>((uint32_t*)out, (const uint32_t*)in, g, total); break;
		case 8: permute_kernel<uint64_t><<<grid, kThreads, 0, pz_stream(stream)>>>((uint64_t*)out, (const uint64_t*)in, g, total); break;
		default: pz_set_error(PZ_ERR_UNSUPPORTED, "permute: unsupported item size %d", itemsize); return PZ_ERR_UNSUPPORTED;
	}
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

// ------------------------------------------------------------------------------------------ random fills + dropout
// reference: Cuda/Source/Libs/CuRand.c:119-230 (fillInteger / fillUniform / fillNormal of a generator object),
// Cuda/Kernels/ElementWise.py:495-580 (dropoutKer, dropout2dKer).  The generator is counter-based (Philox4x32-10, the
// published Random123 algorithm): element i of a fill is a pure function of (seed, offset + i / 4), so fills are
// reproducible from (seed, offset) and need no state in device memory; it is NOT bit-compatible with cuRAND's XORWOW stream.
namespace {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1)
{
	const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
	const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
	const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
	c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__device__ __forceinline__ void philox4x32(unsigned long long counter, unsigned long long seed, uint32_t (&out)[4])
{
	uint32_t c[4] = {(uint32_t)counter, (uint32_t)(counter >> 32), 0u, 0u};
	uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
	#pragma unroll
	for (int r = 0; r < 10; r++) {
		philox_round(c, k0, k1);
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// kind 0: raw 32-bit integers; 1: uniform in (lo, hi]; 2: normal(mean = lo, stddev = hi) by Box-Muller
__global__ void __launch_bounds__(kThreads) rng_fill_kernel(void* __restrict__ out, long long n, unsigned long long seed,
														   unsigned long long offset, const unsigned long long* __restrict__ offset_dev,
														   int kind, float lo, float hi)
{
	if (offset_dev) offset = *offset_dev;          // generator state kept on the device (graph replays draw new numbers)
	const long long quads = (n + 3) / 4;
	for (long long q = (long long)blockIdx.x * kThreads + threadIdx.x; q < quads; q += (long long)gridDim.x * kThreads) {
		uint32_t r[4];
		philox4x32(offset + (unsigned long long)q, seed, r);
		float f[4];
		if (kind == 1) {
			#pragma unroll
			for (int e = 0; e < 4; e++) f[e] = lo + (hi - lo) * ((float)r[e] * 2.3283064365386963e-10f + 2.3283064365386963e-10f * 0.5f);
		} else if (kind == 2) {
			#pragma unroll
			for (int e = 0; e < 4; e += 2) {
				const float u1 = (float)r[e] * 2.3283064365386963e-10f + 2.3283064365386963e-10f * 0.5f;
				const float u2 = (float)r[e + 1] * 2.3283064365386963e-10f;
				const float rad = sqrtf(-2.0f * logf(u1));
				float sn, cs;
				sincospif(2.0f * u2, &sn, &cs);
				f[e] = lo + hi * rad * cs;
				f[e + 1] = lo + hi * rad * sn;
			}
		}
		#pragma unroll
		for (int e = 0; e < 4; e++) {
			const long long i = q * 4 + e;
			if (i < n) {
				if (kind == 0) ((uint32_t*)out)[i] = r[e];
				else ((float*)out)[i] = f[e];
			}
		}
	}
}

// out = x * (b < v) / p with one random word per element (mapsize = 1) or per map of `mapsize` elements (dropout2d)
template <typename T, typename B>
__global__ void __launch_bounds__(kThreads) dropout_kernel(T* __restrict__ out, const T* __restrict__ in, const B* __restrict__ b,
														  unsigned v, float p, long long count, int mapsize, long long start, long long step)
{
	for (long long k = (long long)blockIdx.x * kThreads + threadIdx.x; k < count; k += (long long)gridDim.x * kThreads) {
		const long long i = start + k * step;
		const B word = b[mapsize == 1 ? i : i / mapsize];
		out[i] = from_f<T>(to_f(in[i]) * ((unsigned)word < v ? 1.0f : 0.0f) / p);
	}
}

}  // namespace

extern "C" int pz_rng_fill(int kind, void* out, int64_t n, uint64_t seed, uint64_t offset, float a, float b, void* stream)
{
	PZ_REQUIRE(kind >= 0 && kind <= 2, "rng fill: unknown kind %d", kind);
	if (n <= 0) return PZ_OK;
	rng_fill_kernel<<<grid_for((n + 3) / 4), kThreads, 0, pz_stream(stream)>>>(out, (long long)n, seed, offset, nullptr, kind, a, b);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

namespace {
__global__ void rng_advance_kernel(unsigned long long* offset, unsigned long long by) { *offset += by; }
}

// the same fill with the generator offset read from (and then advanced in) device memory
extern "C" int pz_rng_fill_dev(int kind, void* out, int64_t n, uint64_t seed, void* offset_dev, float a, float b, void* stream)
{
	PZ_REQUIRE(kind >= 0 && kind <= 2, "rng fill: unknown kind %d", kind);
	PZ_REQUIRE(offset_dev != nullptr, "rng fill: null generator state");
	if (n <= 0) return PZ_OK;
	rng_fill_kernel<<<grid_for((n + 3) / 4), kThreads, 0, pz_stream(stream)>>>(out, (long long)n, seed, 0ull,
																				(const unsigned long long*)offset_dev, kind, a, b);
	rng_advance_kernel<<<1, 1, 0, pz_stream(stream)>>>((unsigned long long*)offset_dev, (unsigned long long)((n + 3) / 4));
	pz_count_launch(2);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

extern "C" int pz_dropout_slice(int dtype, void* out, const void* in, const void* rands, uint32_t partition, float p, int64_t n,
								int64_t mapsize, int64_t start, int64_t stop, int64_t step, void* stream)
{
	PZ_REQUIRE(p > 0.0f && mapsize >= 1 && mapsize < (1ll << 31), "dropout: bad arguments");
	PZ_REQUIRE(step >= 1 && start >= 0, "dropout: bad slice");
	if (stop > n) stop = n;
	if (stop <= start) return PZ_OK;
	const long long count = (stop - start + step - 1) / step;
	PzProfScope prof(PZ_PROF_ELTWISE, pz_stream(stream), 0.0, 3.0 * (double)count * pz_dtype_size(dtype));
	const unsigned grid = grid_for(count);
	switch (dtype) {
		case PZ_F32:
			dropout_kernel<float, uint32_t><<<grid, kThreads, 0, pz_stream(stream)>>>((float*)out, (const float*)in, (const uint32_t*)rands,
																					  partition, p, count, (int)mapsize, start, step);
			break;
		case PZ_F16:
			dropout_kernel<__half, uint16_t><<<grid, kThreads, 0, pz_stream(stream)>>>((__half*)out, (const __half*)in, (const uint16_t*)rands,
																					   partition, p, count, (int)mapsize, start, step);
			break;
		case PZ_BF16:
			dropout_kernel<__nv_bfloat16, uint16_t><<<grid, kThreads, 0, pz_stream(stream)>>>(
				(__nv_bfloat16*)out, (const __nv_bfloat16*)in, (const uint16_t*)rands, partition, p, count, (int)mapsize, start, step);
			break;
		default: pz_set_error(PZ_ERR_UNSUPPORTED, "unsupported dtype %d", dtype); return PZ_ERR_UNSUPPORTED;
	}
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

extern "C" int pz_dropout(int dtype, void* out, const void* in, const void* rands, uint32_t partition, float p, int64_t n,
						  int64_t mapsize, void* stream)
{
	return pz_dropout_slice(dtype, out, in, rands, partition, p, n, mapsize, 0, n, 1, stream);
}

// ------------------------------------------------------------------------------------------ local response normalisation
// reference: Cuda/Source/Libs/CuDnnNorm.c:329-690 (cudnnLRNCrossChannel*, cudnnDivisiveNormalization*), host formulas of
// Cuda/Wrappers/CuDnnNorm.py:185-268.  norm_i = K + alpha / |N|^d * sum_{j in win(i)} x_j^2, win(i) = [i - lb, i + la) clipped,
// lb = (N - 1) / 2, la = N - lb; d = 1 across maps (mode 0), d = 2 within a map (mode 1).
//   y_i  = x_i * norm_i^-beta
//   dx_i = g_i * norm_i^-beta - (2 alpha beta / |N|^d) * x_i * sum_{j in win(i)} g_j x_j norm_j^-(beta + 1)
namespace {

struct LrnGeo {
	int C, H, W, n, lb, la, mode;
	float scale, beta, K;              // scale = alpha / N^d
};

template <typename T>
__device__ __forceinline__ float lrn_norm(const T* __restrict__ x, const LrnGeo& g, long long img, int c, int h, int w)
{
	float s = 0.0f;
	if (g.mode == 0) {
		const long long plane = (long long)g.H * g.W, pos = (long long)h * g.W + w;
		for (int j = max(0, c - g.lb); j < min(g.C, c + g.la); j++) { const float v = to_f(x[img + j * plane + pos]); s = fmaf(v, v, s); }
	} else {
		const long long base = img + (long long)c * g.H * g.W;
		for (int yy = max(0, h - g.lb); yy < min(g.H, h + g.la); yy++)
			for (int xx = max(0, w - g.lb); xx < min(g.W, w + g.la); xx++) { const float v = to_f(x[base + (long long)yy * g.W + xx]); s = fmaf(v, v, s); }
	}
	return g.K + g.scale * s;
}

// pass 0: y = x * norm^-beta.  pass 1: tmp = g * x * norm^-(beta+1).  pass 2: dx from g, x, tmp.
template <typename T, int PASS>
__global__ void __launch_bounds__(kThreads) lrn_kernel(const T* __restrict__ x, const T* __restrict__ grad, T* __restrict__ out,
													  float* __restrict__ tmp, LrnGeo g, long long total)
{
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kThreads) {
		const int w = (int)(i % g.W);
		long long r = i / g.W;
		const int h = (int)(r % g.H);
		r /= g.H;
		const int c = (int)(r % g.C);
		const long long img = (r / g.C) * (long long)g.C * g.H * g.W;
		const float norm = lrn_norm(x, g, img, c, h, w);
		const float xv = to_f(x[i]);
		if (PASS == 0) {
			out[i] = from_f<T>(xv * powf(norm, -g.beta));
		} else if (PASS == 1) {
			tmp[i] = to_f(grad[i]) * xv * powf(norm, -(g.beta + 1.0f));
		} else {
			float s = 0.0f;
			if (g.mode == 0) {
				const long long plane = (long long)g.H * g.W, pos = (long long)h * g.W + w;
				for (int j = max(0, c - g.lb); j < min(g.C, c + g.la); j++) s += tmp[img + j * plane + pos];
			} else {
				const long long base = img + (long long)c * g.H * g.W;
				for (int yy = max(0, h - g.lb); yy < min(g.H, h + g.la); yy++)
					for (int xx = max(0, w - g.lb); xx < min(g.W, w + g.la); xx++) s += tmp[base + (long long)yy * g.W + xx];
			}
			out[i] = from_f<T>(to_f(grad[i]) * powf(norm, -g.beta) - 2.0f * g.beta * g.scale * xv * s);
		}
	}
}

template <typename T>
int lrn_launch(int pass, const void* x, const void* grad, void* out, float* tmp, const LrnGeo& g, long long total, void* stream)
{
	const unsigned grid = grid_for(total);
	cudaStream_t s = pz_stream(stream);
	if (pass == 0) lrn_kernel<T, 0><<<grid, kThreads, 0, s>>>((const T*)x, nullptr, (T*)out, nullptr, g, total);
	else {
		lrn_kernel<T, 1><<<grid, kThreads, 0, s>>>((const T*)x, (const T*)grad, nullptr, tmp, g, total);
		lrn_kernel<T, 2><<<grid, kThreads, 0, s>>>((const T*)x, (const T*)grad, (T*)out, tmp, g, total);
		pz_count_launch(1);
	}
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

int lrn_geo(LrnGeo& g, int mode, int64_t C, int64_t H, int64_t W, int n, float alpha, float beta, float K)
{
	PZ_REQUIRE(mode == 0 || mode == 1, "lrn: unknown mode %d", mode);
	PZ_REQUIRE(n >= 1 && C > 0 && H > 0 && W > 0 && C < (1ll << 31) && H < (1ll << 31) && W < (1ll << 31), "lrn: bad geometry");
	g.C = (int)C; g.H = (int)H; g.W = (int)W;
	g.n = n; g.lb = (n - 1) / 2; g.la = n - g.lb; g.mode = mode;
	g.scale = mode == 0 ? alpha / (float)n : alpha / ((float)n * (float)n);
	g.beta = beta; g.K = K;
	return PZ_OK;
}

}  // namespace

extern "C" int pz_lrn_fwd(int dtype, int mode, const void* x, void* y, int64_t N, int64_t C, int64_t H, int64_t W, int n, float alpha,
						  float beta, float K, void* stream)
{
	LrnGeo g{};
	int st = lrn_geo(g, mode, C, H, W, n, alpha, beta, K);
	if (st != PZ_OK) return st;
	const long long total = (long long)N * C * H * W;
	if (total <= 0) return PZ_OK;
	PzProfScope prof(PZ_PROF_ELTWISE, pz_stream(stream), 0.0, 2.0 * (double)total * pz_dtype_size(dtype));
	PZ_TRAIN_DISPATCH(dtype, lrn_launch<T>(0, x, nullptr, y, nullptr, g, total, stream));
}

extern "C" int pz_lrn_bwd(int dtype, int mode, const void* x, const void* grad, void* dx, void* tmp, int64_t N, int64_t C, int64_t H,
						  int64_t W, int n, float alpha, float beta, float K, void* stream)
{
	LrnGeo g{};
	int st = lrn_geo(g, mode, C, H, W, n, alpha, beta, K);
	if (st != PZ_OK) return st;
	const long long total = (long long)N * C * H * W;
	if (total <= 0) return PZ_OK;
	PzProfScope prof(PZ_PROF_ELTWISE, pz_stream(stream), 0.0, 3.0 * (double)total * pz_dtype_size(dtype));
	PZ_TRAIN_DISPATCH(dtype, lrn_launch<T>(1, x, grad, dx, (float*)tmp, g, total, stream));
}

// ------------------------------------------------------------------------------------------ vector reductions
// reference: Cuda/Source/Libs/CuBlas.c (cublasSdot / cublasSasum / cublasSnrm2 behind blas.dot / l1norm / l2norm).
// kind 0: sum x*y, 1: sum |x|, 2: sum x*x (the caller takes the root); fp32 accumulation, *out += result.
namespace {

template <typename T>
__global__ void __launch_bounds__(kThreads) vec_reduce_kernel(const T* __restrict__ x, const T* __restrict__ y, long long n, int kind,
															 float* __restrict__ out)
{
	__shared__ float red[kThreads / 32];
	float s = 0.0f;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
		const float a = to_f(x[i]);
		s += kind == 0 ? a * to_f(y[i]) : (kind == 1 ? fabsf(a) : a * a);
	}
	const float t = block_sum(s, red);
	if (threadIdx.x == 0 && t != 0.0f) atomicAdd(out, t);
}

template <typename T>
int vec_reduce_launch(const void* x, const void* y, int64_t n, int kind, void* out, void* stream)
{
	vec_reduce_kernel<T><<<grid_for(n), kThreads, 0, pz_stream(stream)>>>((const T*)x, (const T*)y, (long long)n, kind, (float*)out);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

}  // namespace

extern "C" int pz_vec_reduce(int dtype, int kind, const void* x, const void* y, int64_t n, void* out, void* stream)
{
	PZ_REQUIRE(kind >= 0 && kind <= 2, "vector reduction: unknown kind %d", kind);
	if (n <= 0) return PZ_OK;
	PZ_TRAIN_DISPATCH(dtype, vec_reduce_launch<T>(x, kind == 0 ? y : x, n, kind, out, stream));
}

// ------------------------------------------------------------------------------------------ grouped matrix-vector product
// reference: Cuda/Kernels/MatVec.py:93-124,311-343 (vecMulOnRow / vecMulOnCol behind matmod.matvec and mulTensorOnVecGroup):
// mat [z][h][w]; on_rows: out[z][h] = beta*out + alpha * sum_w mat*vec[z][w]; else out[z][w] = beta*out + alpha * sum_h mat*vec[z][h].
// Exact fp32 accumulation (not a tensor-core contraction).
namespace {

template <typename T>
__global__ void __launch_bounds__(kThreads) matvec_rows_kernel(T* __restrict__ out, const T* __restrict__ mat, const T* __restrict__ vec,
															  long long rows, int h, int w, float alpha, float beta)
{
	// one warp per output element (z, row)
	const int lane = threadIdx.x & 31;
	for (long long r = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * (kThreads / 32)) {
		const long long z = r / h;
		const T* m = mat + r * w;
		const T* v = vec + z * w;
		float acc = 0.0f;
		for (int i = lane; i < w; i += 32) acc = fmaf(to_f(m[i]), to_f(v[i]), acc);
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
		if (lane == 0) out[r] = from_f<T>((beta != 0.0f ? beta * to_f(out[r]) : 0.0f) + alpha * acc);
	}
}

template <typename T>
__global__ void __launch_bounds__(kThreads) matvec_cols_kernel(T* __restrict__ out, const T* __restrict__ mat, const T* __restrict__ vec,
															  long long cols, int h, int w, float alpha, float beta)
{
	// one thread per output element (z, column): consecutive threads read consecutive columns of a row
	for (long long c = (long long)blockIdx.x * kThreads + threadIdx.x; c < cols; c += (long long)gridDim.x * kThreads) {
		const long long z = c / w;
		const int col = (int)(c - z * w);
		const T* m = mat + z * (long long)h * w + col;
		const T* v = vec + z * h;
		float acc = 0.0f;
		for (int i = 0; i < h; i++) acc = fmaf(to_f(m[(long long)i * w]), to_f(v[i]), acc);
		out[c] = from_f<T>((beta != 0.0f ? beta * to_f(out[c]) : 0.0f) + alpha * acc);
	}
}

template <typename T>
int matvec_launch(void* out, const void* mat, const void* vec, int64_t z, int64_t h, int64_t w, int on_rows, float alpha, float beta,
				  void* stream)
{
	cudaStream_t s = pz_stream(stream);
	if (on_rows)
		matvec_rows_kernel<T><<<grid_for(z * h * 32), kThreads, 0, s>>>((T*)out, (const T*)mat, (const T*)vec, (long long)(z * h), (int)h, (int)w, alpha, beta);
	else
		matvec_cols_kernel<T><<<grid_for(z * w), kThreads, 0, s>>>((T*)out, (const T*)mat, (const T*)vec, (long long)(z * w), (int)h, (int)w, alpha, beta);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

}  // namespace

extern "C" int pz_matvec(int dtype, void* out, const void* mat, const void* vec, int64_t z, int64_t h, int64_t w, int on_rows,
						 float alpha, float beta, void* stream)
{
	PZ_REQUIRE(z >= 0 && h >= 0 && w >= 0 && h < (1ll << 31) && w < (1ll << 31), "matvec: bad shape");
	if (z == 0 || (on_rows ? h : w) == 0) return PZ_OK;
	PZ_TRAIN_DISPATCH(dtype, matvec_launch<T>(out, mat, vec, z, h, w, on_rows, alpha, beta, stream));
}
