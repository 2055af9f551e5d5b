// pz_softmax.cu -- softmax forward / backward (HBM-bandwidth bound).
//
// Replaces cudnnSoftmaxForward / cudnnSoftmaxBackward with CUDNN_SOFTMAX_ACCURATE (max-subtracted) as called
// by CuDnn_Context_softmaxNd / _softmaxNdBackward (reference Cuda/Source/Libs/CuDnn.c:974-1131).
// Tensor viewed as [N][C][S]:
//   mode SPATIAL        (CUDNN_SOFTMAX_MODE_CHANNEL):  normalise over C for every (n, s)
//   mode PER_ACTIVATION (CUDNN_SOFTMAX_MODE_INSTANCE): normalise over C*S for every n
//
// B200 design: when the normalised axis is contiguous (S == 1 or per-activation mode) one WARP owns one row:
// the row is read once into registers when it fits (<= 32 * 32 elements), otherwise re-read from L1/L2; max and
// sum use warp shuffles.  When S > 1 in spatial mode, threads run along the contiguous s axis (coalesced) and
// each thread walks C with stride S using the online (running max / rescaled sum) formulation in one pass
// plus one write pass.
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

__device__ __forceinline__ float warp_max(float v)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
	return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

constexpr int kWarpsPerBlock = 8;
constexpr int kRegElems = 32;     // elements per lane kept in registers

// ---- contiguous rows, one warp per row
template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) softmax_row_fwd(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int64_t L)
{
	const int lane = threadIdx.x & 31;
	const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
	const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
	for (int64_t row = warp0; row < rows; row += nwarps) {
		const T* xr = x + row * L;
		T* yr = y + row * L;
		if (L <= 32 * kRegElems) {
			float v[kRegElems];
			float m = -FLT_MAX;
			#pragma unroll
			for (int i = 0; i < kRegElems; i++) {
				const int64_t j = lane + 32 * i;
				v[i] = j < L ? to_f<T>(xr[j]) : -FLT_MAX;
				m = fmaxf(m, v[i]);
			}
			m = warp_max(m);
			float s = 0.0f;
			#pragma unroll
			for (int i = 0; i < kRegElems; i++) {
				const int64_t j = lane + 32 * i;
				v[i] = j < L ? expf(v[i] - m) : 0.0f;
				s += v[i];
			}
			s = warp_sum(s);
			const float inv = 1.0f / s;
			#pragma unroll
			for (int i = 0; i < kRegElems; i++) {
				const int64_t j = lane + 32 * i;
				if (j < L) yr[j] = from_f<T>(v[i] * inv);
			}
		} else {
			float m = -FLT_MAX;
			for (int64_t j = lane; j < L; j += 32) m = fmaxf(m, to_f<T>(xr[j]));
			m = warp_max(m);
			float s = 0.0f;
			for (int64_t j = lane; j < L; j += 32) s += expf(to_f<T>(xr[j]) - m);
			s = warp_sum(s);
			const float inv = 1.0f / s;
			for (int64_t j = lane; j < L; j += 32) yr[j] = from_f<T>(expf(to_f<T>(xr[j]) - m) * inv);
		}
	}
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) softmax_row_bwd(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx,
																   int64_t rows, int64_t L)
{
	const int lane = threadIdx.x & 31;
	const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
	const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
	for (int64_t row = warp0; row < rows; row += nwarps) {
		const T* yr = y + row * L;
		const T* gr = dy + row * L;
		T* dr = dx + row * L;
		float dot = 0.0f;
		for (int64_t j = lane; j < L; j += 32) dot += to_f<T>(yr[j]) * to_f<T>(gr[j]);
		dot = warp_sum(dot);
		for (int64_t j = lane; j < L; j += 32) dr[j] = from_f<T>(to_f<T>(yr[j]) * (to_f<T>(gr[j]) - dot));
	}
}

// ---- strided (spatial mode with S > 1): thread per (n, s)
template <typename T>
__global__ void __launch_bounds__(256) softmax_strided_fwd(const T* __restrict__ x, T* __restrict__ y, int64_t N, int64_t C, int64_t S)
{
	const int64_t total = N * S;
	for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
		const int64_t n = i / S, s = i % S;
		const T* xp = x + n * C * S + s;
		T* yp = y + n * C * S + s;
		float m = -FLT_MAX, sum = 0.0f;
		for (int64_t c = 0; c < C; c++) {
			const float v = to_f<T>(xp[c * S]);
			if (v > m) { sum = sum * expf(m - v); m = v; }
			sum += expf(v - m);
		}
		const float inv = 1.0f / sum;
		for (int64_t c = 0; c < C; c++) yp[c * S] = from_f<T>(expf(to_f<T>(xp[c * S]) - m) * inv);
	}
}

template <typename T>
__global__ void __launch_bounds__(256) softmax_strided_bwd(const T* __restrict__ y, const T* __restrict__ dy, T* __restrict__ dx, int64_t N,
													   int64_t C, int64_t S)
{
	const int64_t total = N * S;
	for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
		const int64_t n = i / S, s = i % S;
		const int64_t off = n * C * S + s;
		float dot = 0.0f;
		for (int64_t c = 0; c < C; c++) dot += to_f<T>(y[off + c * S]) * to_f<T>(dy[off + c * S]);
		for (int64_t c = 0; c < C; c++) dx[off + c * S] = from_f<T>(to_f<T>(y[off + c * S]) * (to_f<T>(dy[off + c * S]) - dot));
	}
}

unsigned cap_blocks(int64_t blocks)
{
	const int64_t cap = (int64_t)pz_num_sms() * 16;
	if (blocks > cap) blocks = cap;
	return (unsigned)(blocks < 1 ? 1 : blocks);
}

template <typename T>
int fwd(int mode, const void* x, void* y, int64_t N, int64_t C, int64_t S, void* stream)
{
	if (mode == PZ_SOFTMAX_PER_ACTIVATION || S == 1) {
		const int64_t rows = N, L = C * S;
		softmax_row_fwd<T><<<cap_blocks(pz_cdiv(rows, kWarpsPerBlock)), kWarpsPerBlock * 32, 0, pz_stream(stream)>>>((const T*)x, (T*)y, rows, L);
	} else {
		softmax_strided_fwd<T><<<cap_blocks(pz_cdiv(N * S, 256)), 256, 0, pz_stream(stream)>>>((const T*)x, (T*)y, N, C, S);
	}
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

template <typename T>
int bwd(int mode, const void* y, const void* dy, void* dx, int64_t N, int64_t C, int64_t S, void* stream)
{
	if (mode == PZ_SOFTMAX_PER_ACTIVATION || S == 1) {
		const int64_t rows = N, L = C * S;
		softmax_row_bwd<T><<<cap_blocks(pz_cdiv(rows, kWarpsPerBlock)), kWarpsPerBlock * 32, 0, pz_stream(stream)>>>((const T*)y, (const T*)dy,
																												  (T*)dx, rows, L);
	} else {
		softmax_strided_bwd<T><<<cap_blocks(pz_cdiv(N * S, 256)), 256, 0, pz_stream(stream)>>>((const T*)y, (const T*)dy, (T*)dx, N, C, S);
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

int pz_softmax_fwd(int dtype, int mode, const void* x, void* y, int64_t N, int64_t C, int64_t S, void* stream)
{
	PZ_REQUIRE(N > 0 && C > 0 && S > 0, "softmax: empty tensor");
	PZ_REQUIRE(mode == PZ_SOFTMAX_PER_ACTIVATION || mode == PZ_SOFTMAX_SPATIAL, "softmax: unknown mode %d", mode);
	PZ_DISPATCH_FLOAT(dtype, fwd<T>(mode, x, y, N, C, S, stream));
}

int pz_softmax_bwd(int dtype, int mode, const void* y, const void* dy, void* dx, int64_t N, int64_t C, int64_t S, void* stream)
{
	PZ_REQUIRE(N > 0 && C > 0 && S > 0, "softmax: empty tensor");
	PZ_REQUIRE(mode == PZ_SOFTMAX_PER_ACTIVATION || mode == PZ_SOFTMAX_SPATIAL, "softmax: unknown mode %d", mode);
	PZ_DISPATCH_FLOAT(dtype, bwd<T>(mode, y, dy, dx, N, C, S, stream));
}

}  // extern "C"
