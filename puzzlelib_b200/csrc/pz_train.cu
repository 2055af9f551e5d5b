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
