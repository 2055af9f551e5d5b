// pz_extra.cu -- the remaining elementwise kernels the reference's Backend/Kernels/ElementWise.py and Costs.py bind from
// the backend object (optimizer updates beyond momentum SGD / Adam, weight decay, L1 helpers, RBM sampling, the pointwise
// costs), plus the strided `slice=` launch form of the elementwise kernels (Cuda/SourceModule.py:162-200: the element
// index runs over start, start + step, ... < stop of EVERY pointer argument).
//
// None of this is on the ResNet / VGG hot path: one generic grid-stride kernel with a runtime switch, scalar accesses
// (a strided slice cannot be vectorised anyway).  Math in fp32 for every storage type, like the reference kernels.
#include "pz_common.h"

#include <climits>

namespace {

constexpr int kThreads = 256;

template <typename T> __device__ __forceinline__ float ld(const void* p, long long i) { return (float)((const T*)p)[i]; }
template <> __device__ __forceinline__ float ld<__half>(const void* p, long long i) { return __half2float(((const __half*)p)[i]); }
template <> __device__ __forceinline__ float ld<__nv_bfloat16>(const void* p, long long i) { return __bfloat162float(((const __nv_bfloat16*)p)[i]); }
template <typename T> __device__ __forceinline__ void st(void* p, long long i, float v) { ((T*)p)[i] = (T)v; }
template <> __device__ __forceinline__ void st<__half>(void* p, long long i, float v) { ((__half*)p)[i] = __float2half_rn(v); }
template <> __device__ __forceinline__ void st<__nv_bfloat16>(void* p, long long i, float v) { ((__nv_bfloat16*)p)[i] = __float2bfloat16_rn(v); }

struct EwArgs {
	void* p[6];
	float s[4];
	long long start, step, count;      // element i = start + k * step, k < count
	int aux[2];
};

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
	return s;
}

// reference formulas: Cuda/Kernels/ElementWise.py (file:line next to every case)
template <typename T>
__global__ void __launch_bounds__(kThreads) ew_generic_kernel(int op, EwArgs a)
{
	__shared__ float red[kThreads / 32];
	float err = 0.0f;

	for (long long k = (long long)blockIdx.x * kThreads + threadIdx.x; k < a.count; k += (long long)gridDim.x * kThreads) {
		const long long i = a.start + k * a.step;
		switch (op) {
			case PZ_EW_ABS:              // absKer :1117-1122
				st<T>(a.p[0], i, fabsf(ld<T>(a.p[1], i)));
				break;
			case PZ_EW_WEIGHT_DECAY:     // weightDecayKer :1109-1114  grad -= rate * param
				st<T>(a.p[0], i, ld<T>(a.p[0], i) - a.s[0] * ld<T>(a.p[1], i));
				break;
			case PZ_EW_L1_PENALTY: {     // l1penaltyKer :1125-1130
				const float d = ld<T>(a.p[2], i);
				st<T>(a.p[0], i, ld<T>(a.p[1], i) - a.s[0] * ((0.0f <= d ? 1.0f : 0.0f) - (d < 0.0f ? 1.0f : 0.0f)));
				break;
			}
			case PZ_EW_L1_GRAD:          // l1gradKer :1133-1138
				st<T>(a.p[0], i, ld<T>(a.p[1], i) - ld<T>(a.p[2], i) > 0.0f ? -a.s[0] : a.s[0]);
				break;
			case PZ_EW_RBM: {            // rbmKer :1100-1106  out = uni < sigmoid(in)
				const float p = 1.0f / (1.0f + expf(-ld<T>(a.p[1], i)));
				st<T>(a.p[0], i, ld<T>(a.p[2], i) < p ? 1.0f : 0.0f);
				break;
			}
			case PZ_EW_RMSPROP: {        // rmspropKer :860-903  (param, grad, ms; lr, factor, eps)
				const float g = ld<T>(a.p[1], i);
				const float ms = a.s[1] * ld<T>(a.p[2], i) + (1.0f - a.s[1]) * g * g;
				st<T>(a.p[0], i, ld<T>(a.p[0], i) + a.s[0] * g / (sqrtf(ms) + a.s[2]));
				st<T>(a.p[2], i, ms);
				break;
			}
			case PZ_EW_RMSPROP_GRAVES: { // rmspropGravesKer :906-954  (param, grad, mg, ms, delta; lr, alpha, momRate, eps)
				const float g = ld<T>(a.p[1], i);
				const float mg = a.s[1] * ld<T>(a.p[2], i) + (1.0f - a.s[1]) * g;
				const float ms = a.s[1] * ld<T>(a.p[3], i) + (1.0f - a.s[1]) * g * g;
				const float delta = a.s[2] * ld<T>(a.p[4], i) + a.s[0] * g / sqrtf(ms - mg * mg + a.s[3]);
				st<T>(a.p[0], i, ld<T>(a.p[0], i) + delta);
				st<T>(a.p[2], i, mg);
				st<T>(a.p[3], i, ms);
				st<T>(a.p[4], i, delta);
				break;
			}
			case PZ_EW_ADAGRAD: {        // adagradKer :664-706  (param, grad, h; lr, eps)
				const float g = ld<T>(a.p[1], i);
				const float h = ld<T>(a.p[2], i) + g * g;
				st<T>(a.p[0], i, ld<T>(a.p[0], i) + a.s[0] * g / (sqrtf(h) + a.s[1]));
				st<T>(a.p[2], i, h);
				break;
			}
			case PZ_EW_ADADELTA: {       // adadeltaKer :614-661  (param, grad, msg, msdx; rho, eps)
				const float g = ld<T>(a.p[1], i);
				float msg = ld<T>(a.p[2], i), msdx = ld<T>(a.p[3], i);
				msg += (1.0f - a.s[0]) * (g * g - msg);
				const float dx = sqrtf((msdx + a.s[1]) / (msg + a.s[1])) * g;
				msdx += (1.0f - a.s[0]) * (dx * dx - msdx);
				st<T>(a.p[0], i, ld<T>(a.p[0], i) + dx);
				st<T>(a.p[2], i, msg);
				st<T>(a.p[3], i, msdx);
				break;
			}
			case PZ_EW_SMORMS3: {        // smorms3Ker :957-1003  (param, grad: T; mem, mg, ms: fp32; lr, eps)
				const float g = ld<T>(a.p[1], i);
				float mem = ((float*)a.p[2])[i], mg = ((float*)a.p[3])[i], ms = ((float*)a.p[4])[i];
				const float r = 1.0f / (mem + 1.0f);
				mg = (1.0f - r) * mg + r * g;
				ms = (1.0f - r) * ms + r * g * g;
				const float x = mg * mg / (ms + a.s[1]);
				mem = 1.0f + mem * (1.0f - x);
				st<T>(a.p[0], i, ld<T>(a.p[0], i) + g * fminf(a.s[0], x) / (sqrtf(ms) + a.s[1]));
				((float*)a.p[2])[i] = mem;
				((float*)a.p[3])[i] = mg;
				((float*)a.p[4])[i] = ms;
				break;
			}
			// ---- pointwise costs (fp32 only), reference: Cuda/Kernels/Costs.py:8-74; p = {a, b, totalError, grad(s)}
			case PZ_EW_BCE: {            // (scores, labels:int, totalError, grad; numsamples, spatialDim)
				const float prob = 1.0f / (1.0f + expf(-((const float*)a.p[0])[i]));
				const int label = ((const int*)a.p[1])[i];
				err += (label == 1 ? -logf(prob) : -logf(1.0f - prob)) / (float)a.aux[1];
				((float*)a.p[3])[i] = ((label == 1 ? 1.0f : 0.0f) - prob) / (float)a.aux[0] / (float)a.aux[1];
				break;
			}
			case PZ_EW_HINGE: {          // (scores, labels:int, totalError, grad; numsamples, numcases)
				const float score = ((const float*)a.p[0])[i];
				const int label = ((const int*)a.p[1])[i];
				err += fmaxf(0.0f, 1.0f - score * label) / (float)a.aux[1];
				((float*)a.p[3])[i] = score * label < 1.0f ? (float)label / (float)a.aux[0] / (float)a.aux[1] : 0.0f;
				break;
			}
			case PZ_EW_SMOOTH_L1: {      // (pred, target, totalError, grad; norm, fullnorm)
				const float diff = ((const float*)a.p[0])[i] - ((const float*)a.p[1])[i];
				const float sign = diff > 0.0f ? 1.0f : -1.0f;
				err += diff * sign < 1.0f ? diff * diff / 2.0f * a.s[0] : (sign * diff - 0.5f) * a.s[0];
				((float*)a.p[3])[i] = diff * sign < 1.0f ? diff * a.s[1] : sign * a.s[1];
				break;
			}
			case PZ_EW_L1_HINGE: {       // (x1, x2, labels:int, totalError, g1, g2; numsamples, numcases)
				const float diff = ((const float*)a.p[0])[i] - ((const float*)a.p[1])[i];
				const float sign = diff > 0.0f ? 1.0f : -1.0f;
				const int label = ((const int*)a.p[2])[i / a.aux[1]];
				const float ad = fabsf(diff), norm = (float)a.aux[0] * (float)a.aux[1];
				err += (label == 0 ? fmaxf(0.0f, 1.0f - ad) : ad) / (float)a.aux[1];
				((float*)a.p[4])[i] = (label == 0 ? (ad < 1.0f ? -sign : 0.0f) : sign) / norm;
				((float*)a.p[5])[i] = (label == 0 ? (ad < 1.0f ? sign : 0.0f) : -sign) / norm;
				break;
			}
			default: break;
		}
	}

	if (op >= PZ_EW_BCE) {
		const float s = block_sum(err, red);
		const int slot = op == PZ_EW_L1_HINGE ? 3 : 2;
		if (threadIdx.x == 0 && s != 0.0f) atomicAdd((float*)a.p[slot], s);
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

extern "C" int pz_eltwise(int op, int dtype, void* const* ptrs, int nptrs, const float* scalars, int nscalars, const int* aux,
						  int64_t n, int64_t start, int64_t stop, int64_t step, void* stream)
{
	PZ_REQUIRE(op >= PZ_EW_ABS && op <= PZ_EW_L1_HINGE, "eltwise: unknown op %d", op);
	PZ_REQUIRE(nptrs >= 1 && nptrs <= 6 && nscalars >= 0 && nscalars <= 4, "eltwise: bad argument counts");
	PZ_REQUIRE(step >= 1 && start >= 0, "eltwise: bad slice");
	if (stop > n) stop = n;
	if (stop <= start) return PZ_OK;
	if (op >= PZ_EW_BCE) PZ_REQUIRE(dtype == PZ_F32, "eltwise: the pointwise costs are float32 only");

	EwArgs a{};
	for (int i = 0; i < nptrs; i++) a.p[i] = ptrs[i];
	for (int i = 0; i < nscalars; i++) a.s[i] = scalars[i];
	if (aux) { a.aux[0] = aux[0]; a.aux[1] = aux[1]; }
	a.start = start;
	a.step = step;
	a.count = (stop - start + step - 1) / step;

	const unsigned grid = grid_for(a.count);
	cudaStream_t s = pz_stream(stream);
	switch (dtype) {
		case PZ_F32: ew_generic_kernel<float><<<grid, kThreads, 0, s>>>(op, a); break;
		case PZ_F16: ew_generic_kernel<__half><<<grid, kThreads, 0, s>>>(op, a); break;
		case PZ_BF16: ew_generic_kernel<__nv_bfloat16><<<grid, kThreads, 0, s>>>(op, a); break;
		default: pz_set_error(PZ_ERR_UNSUPPORTED, "unsupported dtype %d", dtype); return PZ_ERR_UNSUPPORTED;
	}
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

// ---- the remaining accuracy / divergence reductions of getAccuracyKernel (reference: Cuda/Kernels/Costs.py:172-205):
//   kind 0 calcBCEAccuracy(x f32, y i32):             sum of (y == 1 ? x <= 0 : x > 0)
//   kind 1 klDivergence(x f32, y f32, grad, gradnorm): grad = (y - x) * gradnorm; sum of y > 0 ? y * (log y - log x) : 0
//   kind 2 l1HingeAccuracy(d f32, labels i32):        sum of ((d <= 1) != labels)
namespace {

__global__ void __launch_bounds__(kThreads) cost_reduce_kernel(int kind, const float* __restrict__ x, const void* __restrict__ y,
															  float* __restrict__ grad, float gradnorm, long long n, float* __restrict__ out)
{
	__shared__ float red[kThreads / 32];
	float acc = 0.0f;
	for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
		const float xv = x[i];
		if (kind == 0) {
			acc += (((const int*)y)[i] == 1 ? xv <= 0.0f : xv > 0.0f) ? 1.0f : 0.0f;
		} else if (kind == 1) {
			const float yv = ((const float*)y)[i];
			grad[i] = (yv - xv) * gradnorm;
			acc += yv > 0.0f ? yv * (logf(yv) - logf(xv)) : 0.0f;
		} else {
			acc += ((xv <= 1.0f ? 1 : 0) != ((const int*)y)[i]) ? 1.0f : 0.0f;
		}
	}
	const float s = block_sum(acc, red);
	if (threadIdx.x == 0 && s != 0.0f) atomicAdd(out, s);
}

template <bool MAX>
__global__ void __launch_bounds__(1024) int_minmax_kernel(const int* __restrict__ in, long long n, int* __restrict__ out)
{
	__shared__ int part[32];
	int acc = MAX ? INT_MIN : INT_MAX;
	for (long long i = threadIdx.x; i < n; i += blockDim.x) acc = MAX ? max(acc, in[i]) : min(acc, in[i]);
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		const int other = __shfl_xor_sync(0xffffffffu, acc, o);
		acc = MAX ? max(acc, other) : min(acc, other);
	}
	if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x < 32) {
		acc = part[threadIdx.x];
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			const int other = __shfl_xor_sync(0xffffffffu, acc, o);
			acc = MAX ? max(acc, other) : min(acc, other);
		}
		if (threadIdx.x == 0) out[0] = acc;
	}
}

}  // namespace

extern "C" int pz_cost_reduce(int kind, const void* x, const void* y, void* grad, float gradnorm, int64_t n, void* out, void* stream)
{
	PZ_REQUIRE(kind >= 0 && kind <= 2, "cost reduction: unknown kind %d", kind);
	if (n <= 0) return PZ_OK;
	cost_reduce_kernel<<<grid_for(n), kThreads, 0, pz_stream(stream)>>>(kind, (const float*)x, y, (float*)grad, gradnorm, (long long)n,
																	   (float*)out);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

// GPUArray.min() / max() of an int32 array (Cost/CrossEntropy.py:84-92 verifies labels with them)
extern "C" int pz_reduce_minmax_i32(const void* in, int64_t n, int want_max, void* out, void* stream)
{
	PZ_REQUIRE(n > 0, "empty reduction");
	if (want_max) int_minmax_kernel<true><<<1, 1024, 0, pz_stream(stream)>>>((const int*)in, (long long)n, (int*)out);
	else int_minmax_kernel<false><<<1, 1024, 0, pz_stream(stream)>>>((const int*)in, (long long)n, (int*)out);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}
