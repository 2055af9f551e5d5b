// pz_elementwise.cu -- bandwidth-bound elementwise kernels (activations, BLAS-1 style updates, casts).
//
// Replaces the NVRTC-JIT one-thread-per-element kernels of the reference
// (Cuda/Kernels/ElementWise.py, Cuda/SourceModule.py:176-200): here every kernel moves 128-bit words
// (ld.global.v4 / st.global.v4), runs a grid-stride loop sized to the SM count with several independent
// loads in flight per thread, and peels unaligned heads/tails so views at odd offsets still work.
// Math is done in fp32 for every storage type, like the reference; this file is compiled with
// --use_fast_math because the reference compiles its kernels that way (SourceModule.py:105-112).
#include "pz_common.h"

namespace {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T> struct alignas(16) Pack { T v[16 / sizeof(T)]; };

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// NIN inputs (read-only) + NIO in/out arrays.  Op::apply(float* io, const float* in) works on one element:
// io[] holds the current values of the in/out arrays (undefined for pure outputs) and receives the results.
template <typename T, int NIO, int NIN, bool READ_IO, typename Op>
__global__ void __launch_bounds__(kThreads) ew_kernel(T* io0, T* io1, const T* in0, const T* in1, const T* in2,
													  int64_t n, int64_t head, int64_t nvec, Op op)
{
	constexpr int VEC = 16 / sizeof(T);
	T* ios[2] = {io0, io1};
	const T* ins[3] = {in0, in1, in2};

	const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;

	// vector body
	for (int64_t base = tid; base < nvec; base += nthreads * kUnroll) {
		Pack<T> pin[kUnroll][NIN > 0 ? NIN : 1];
		Pack<T> pio[kUnroll][NIO];

		#pragma unroll
		for (int u = 0; u < kUnroll; u++) {
			int64_t v = base + (int64_t)u * nthreads;
			if (v < nvec) {
				#pragma unroll
				for (int k = 0; k < NIN; k++)
					pin[u][k] = *reinterpret_cast<const Pack<T>*>(ins[k] + head + v * VEC);
				if (READ_IO) {
					#pragma unroll
					for (int k = 0; k < NIO; k++)
						pio[u][k] = *reinterpret_cast<const Pack<T>*>(ios[k] + head + v * VEC);
				}
			}
		}

		#pragma unroll
		for (int u = 0; u < kUnroll; u++) {
			int64_t v = base + (int64_t)u * nthreads;
			if (v < nvec) {
				#pragma unroll
				for (int e = 0; e < VEC; e++) {
					float fin[NIN > 0 ? NIN : 1], fio[NIO];
					#pragma unroll
					for (int k = 0; k < NIN; k++) fin[k] = to_f<T>(pin[u][k].v[e]);
					#pragma unroll
					for (int k = 0; k < NIO; k++) fio[k] = READ_IO ? to_f<T>(pio[u][k].v[e]) : 0.0f;
					op.apply(fio, fin);
					#pragma unroll
					for (int k = 0; k < NIO; k++) pio[u][k].v[e] = from_f<T>(fio[k]);
				}
				#pragma unroll
				for (int k = 0; k < NIO; k++)
					*reinterpret_cast<Pack<T>*>(ios[k] + head + v * VEC) = pio[u][k];
			}
		}
	}

	// scalar head + tail (at most 2 * (VEC - 1) elements, or everything when pointers are mutually misaligned)
	const int64_t tailstart = head + nvec * VEC;
	const int64_t nscalar = head + (n - tailstart);
	for (int64_t s = tid; s < nscalar; s += nthreads) {
		int64_t i = s < head ? s : tailstart + (s - head);
		float fin[NIN > 0 ? NIN : 1], fio[NIO];
		#pragma unroll
		for (int k = 0; k < NIN; k++) fin[k] = to_f<T>(ins[k][i]);
		#pragma unroll
		for (int k = 0; k < NIO; k++) fio[k] = READ_IO ? to_f<T>(ios[k][i]) : 0.0f;
		op.apply(fio, fin);
		#pragma unroll
		for (int k = 0; k < NIO; k++) ios[k][i] = from_f<T>(fio[k]);
	}
}

template <typename T, int NIO, int NIN, bool READ_IO, typename Op>
int ew_launch(void* io0, void* io1, const void* in0, const void* in1, const void* in2, int64_t n, Op op, void* stream)
{
	if (n <= 0) return PZ_OK;
	constexpr int VEC = 16 / sizeof(T);

	// all arrays must share the same misalignment for the vector path
	uintptr_t ptrs[5] = {(uintptr_t)io0, (uintptr_t)io1, (uintptr_t)in0, (uintptr_t)in1, (uintptr_t)in2};
	bool used[5] = {NIO > 0, NIO > 1, NIN > 0, NIN > 1, NIN > 2};
	uintptr_t mis = ptrs[0] % 16;
	bool same = true;
	for (int i = 0; i < 5; i++)
		if (used[i] && ptrs[i] % 16 != mis) same = false;

	int64_t head = 0, nvec = 0;
	if (same && mis % sizeof(T) == 0) {
		head = mis ? (int64_t)((16 - mis) / sizeof(T)) : 0;
		if (head > n) head = n;
		nvec = (n - head) / VEC;
	}

	int64_t work = nvec > 0 ? pz_cdiv(nvec, kUnroll) : n;
	int64_t blocks = pz_cdiv(work, kThreads);
	int64_t maxblocks = (int64_t)pz_num_sms() * 8;
	if (blocks > maxblocks) blocks = maxblocks;
	if (blocks < 1) blocks = 1;

	PzProfScope prof(PZ_PROF_ELTWISE, pz_stream(stream), 0.0, (double)n * sizeof(T) * (NIN + NIO * (READ_IO ? 2 : 1)));
	ew_kernel<T, NIO, NIN, READ_IO, Op><<<(unsigned)blocks, kThreads, 0, pz_stream(stream)>>>(
		(T*)io0, (T*)io1, (const T*)in0, (const T*)in1, (const T*)in2, n, head, nvec, op);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

// `slice=` form of an elementwise launch (reference: Cuda/SourceModule.py:162-200, the `<name>_strided` twin every
// ElementwiseKernel is compiled with): element i = start + k * step < stop of every pointer argument
template <typename T, int NIO, int NIN, bool READ_IO, typename Op>
__global__ void __launch_bounds__(kThreads) ew_strided_kernel(T* io0, T* io1, const T* in0, const T* in1, const T* in2,
															  int64_t start, int64_t step, int64_t count, Op op)
{
	T* ios[2] = {io0, io1};
	const T* ins[3] = {in0, in1, in2};
	for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t)gridDim.x * blockDim.x) {
		const int64_t i = start + k * step;
		float fin[NIN > 0 ? NIN : 1], fio[NIO];
		#pragma unroll
		for (int j = 0; j < NIN; j++) fin[j] = to_f<T>(ins[j][i]);
		#pragma unroll
		for (int j = 0; j < NIO; j++) fio[j] = READ_IO ? to_f<T>(ios[j][i]) : 0.0f;
		op.apply(fio, fin);
		#pragma unroll
		for (int j = 0; j < NIO; j++) ios[j][i] = from_f<T>(fio[j]);
	}
}

template <typename T, int NIO, int NIN, bool READ_IO, typename Op>
int ew_launch_slice(void* io0, void* io1, const void* in0, const void* in1, const void* in2, int64_t n, int64_t start, int64_t stop,
					int64_t step, Op op, void* stream)
{
	PZ_REQUIRE(step >= 1 && start >= 0, "bad slice (start %lld, step %lld)", (long long)start, (long long)step);
	if (stop > n) stop = n;
	if (stop <= start) return PZ_OK;
	const int64_t count = (stop - start + step - 1) / step;
	int64_t blocks = pz_cdiv(count, kThreads);
	const int64_t maxblocks = (int64_t)pz_num_sms() * 8;
	if (blocks > maxblocks) blocks = maxblocks;
	ew_strided_kernel<T, NIO, NIN, READ_IO, Op><<<(unsigned)blocks, kThreads, 0, pz_stream(stream)>>>(
		(T*)io0, (T*)io1, (const T*)in0, (const T*)in1, (const T*)in2, start, step, count, op);
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

// ------------------------------------------------------------------ activation functors
// Formulas are kept in the reference's algebraic form (ElementWise.py:18,45,73,100,128,156,184-188,
// 215-220,249-254,281-286,313,341,369-373,403-407,436-441,468-477) so that NaN / -0 / boundary
// behaviour is the same.
struct ActFwd {
	int kind;
	float a, b;
	__device__ __forceinline__ void apply(float* io, const float* in) const
	{
		float x = in[0], y;
		switch (kind) {
			case PZ_ACT_SIGMOID: y = 1.0f / (1.0f + expf(-x)); break;
			case PZ_ACT_TANH: y = tanhf(x); break;
			case PZ_ACT_RELU: y = x * (x > 0.0f); break;
			case PZ_ACT_LEAKYRELU: y = x * ((x > 0.0f) + a * (x <= 0.0f)); break;
			case PZ_ACT_ELU: y = x * (x > 0.0f) + a * (expf(x) - 1.0f) * (x <= 0.0f); break;
			case PZ_ACT_SOFTPLUS: y = logf(1.0f + expf(x)); break;
			case PZ_ACT_CLIP: y = fminf(b, fmaxf(a, x)); break;
			default: y = 0.5f * x * (1.0f + erff(x / 1.4142135623730951f)); break;  // gelu
		}
		io[0] = y;
	}
};

struct ActBwd {
	int kind;
	float a, b;
	__device__ __forceinline__ void apply(float* io, const float* in) const
	{
		float g = in[0], d = in[1], r;
		switch (kind) {
			case PZ_ACT_SIGMOID: r = g * d * (1.0f - d); break;
			case PZ_ACT_TANH: r = g * (1.0f - d * d); break;
			case PZ_ACT_RELU: r = g * (d > 0.0f); break;
			case PZ_ACT_LEAKYRELU: r = g * ((d > 0.0f) + a * (d <= 0.0f)); break;
			case PZ_ACT_ELU: r = g * ((d > 0.0f) + (d + a) * (d <= 0.0f)); break;
			case PZ_ACT_SOFTPLUS: r = g * (1.0f - expf(-d)); break;
			case PZ_ACT_CLIP: r = g * (d > a && d < b); break;
			default:  // gelu: d is the INPUT; the reference's Gaussian term uses 1/sqrt(pi) (sic)
				r = g * (0.5f * (1.0f + erff(d / 1.4142135623730951f)) + d / 1.7724538509055159f * expf(-0.5f * d * d));
				break;
		}
		io[0] = r;
	}
};

// Specialised relu functors: the hot ones for ResNet/VGG get their own instantiation (no switch in the loop).
struct ReluFwd {
	__device__ __forceinline__ void apply(float* io, const float* in) const { io[0] = in[0] * (in[0] > 0.0f); }
};
struct ReluBwd {
	__device__ __forceinline__ void apply(float* io, const float* in) const { io[0] = in[0] * (in[1] > 0.0f); }
};

// y = (0 + x1*a1) + x2*a2 with the intermediate rounded to the storage type: bit for bit what `y.fill(0); y += a1*x1; y += a2*x2`
// leaves in y (three launches of toVectorAddVectorKer, ElementWise.py:583-609), in one pass.  in[1] unused when a2 == 0 / x2 NULL.
template <typename T, bool TWO>
struct Axpy2 {
	float a1, a2;
	__device__ __forceinline__ void apply(float* io, const float* in) const
	{
		const float first = to_f<T>(from_f<T>(fmaf(in[0], a1, 0.0f)));
		io[0] = TWO ? fmaf(in[1], a2, first) : first;
	}
};
// the same sum followed by the ReLU the next module applies to it: io[0] = the sum (stored, as the reference's Add module keeps it),
// io[1] = max(sum, 0) computed from the ROUNDED sum -- bit for bit what the separate relu kernel would read and write
template <typename T>
struct Axpy2Relu {
	float a1, a2;
	__device__ __forceinline__ void apply(float* io, const float* in) const
	{
		const float first = to_f<T>(from_f<T>(fmaf(in[0], a1, 0.0f)));
		const float sum = to_f<T>(from_f<T>(fmaf(in[1], a2, first)));
		io[0] = sum;
		io[1] = sum * (sum > 0.0f);
	}
};
// ... and followed by the ReLU DERIVATIVE (Replicate.updateGrad's sum of two gradients, then the previous block's Activation.updateGrad):
// io[0] = the sum, io[1] = sum * (ref > 0) with in[2] = ref, the ReLU's output
template <typename T>
struct Axpy2ReluBwd {
	float a1, a2;
	__device__ __forceinline__ void apply(float* io, const float* in) const
	{
		const float first = to_f<T>(from_f<T>(fmaf(in[0], a1, 0.0f)));
		const float sum = to_f<T>(from_f<T>(fmaf(in[1], a2, first)));
		io[0] = sum;
		io[1] = sum * (in[2] > 0.0f);
	}
};
struct Axpy {   // y = y + x * alpha   (ElementWise.py:591)
	float alpha;
	__device__ __forceinline__ void apply(float* io, const float* in) const { io[0] = io[0] + in[0] * alpha; }
};
struct Axpby {  // out = x * alpha + y * beta   (ElementWise.py:1030)
	float alpha, beta;
	__device__ __forceinline__ void apply(float* io, const float* in) const { io[0] = in[0] * alpha + in[1] * beta; }
};
struct ScaleShift {  // out = a * in + b   (ElementWise.py:1084)
	float a, b;
	__device__ __forceinline__ void apply(float* io, const float* in) const { io[0] = a * in[0] + b; }
};
struct Mul {
	__device__ __forceinline__ void apply(float* io, const float* in) const { io[0] = in[0] * in[1]; }
};
struct Add2 {   // (0 + a) + b, the value Add.py:15-23 produces with fill(0) + two axpy(alpha=1) passes
	__device__ __forceinline__ void apply(float* io, const float* in) const { io[0] = (0.0f + in[0] * 1.0f) + in[1] * 1.0f; }
};
struct SgdMom {  // io[0] = param, io[1] = mom, in[0] = grad   (ElementWise.py:773-786)
	float lr, mr;
	__device__ __forceinline__ void apply(float* io, const float* in) const
	{
		float m = mr * io[1] + lr * in[0];
		io[1] = m;
		io[0] = io[0] + m;
	}
};

// ------------------------------------------------------------------ casts
template <typename TD, typename TS>
__global__ void __launch_bounds__(kThreads) cast_kernel(TD* __restrict__ dst, const TS* __restrict__ src, int64_t n)
{
	// 4 elements per thread per step; vector access when both sides are aligned for it
	const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
	const bool aligned = ((uintptr_t)dst % (4 * sizeof(TD)) == 0) && ((uintptr_t)src % (4 * sizeof(TS)) == 0);
	const int64_t n4 = aligned ? n / 4 : 0;

	struct alignas(4 * sizeof(TS)) S4 { TS v[4]; };
	struct alignas(4 * sizeof(TD)) D4 { TD v[4]; };

	for (int64_t i = tid; i < n4; i += nthreads) {
		S4 s = reinterpret_cast<const S4*>(src)[i];
		D4 d;
		#pragma unroll
		for (int e = 0; e < 4; e++) d.v[e] = from_f<TD>(to_f<TS>(s.v[e]));
		reinterpret_cast<D4*>(dst)[i] = d;
	}
	for (int64_t i = n4 * 4 + tid; i < n; i += nthreads) dst[i] = from_f<TD>(to_f<TS>(src[i]));
}

template <typename TD, typename TS>
int cast_launch(void* dst, const void* src, int64_t n, void* stream)
{
	if (n <= 0) return PZ_OK;
	int64_t blocks = pz_cdiv(pz_cdiv(n, 4), kThreads);
	int64_t maxblocks = (int64_t)pz_num_sms() * 8;
	if (blocks > maxblocks) blocks = maxblocks;
	cast_kernel<TD, TS><<<(unsigned)blocks, kThreads, 0, pz_stream(stream)>>>((TD*)dst, (const TS*)src, n);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

// ------------------------------------------------------------------ min / max reduction
template <typename T, bool MAX>
__global__ void __launch_bounds__(1024) minmax_kernel(const T* __restrict__ in, int64_t n, T* __restrict__ out)
{
	// single CTA, two-stage (thread -> warp shuffle -> smem -> warp shuffle); used for GPUArray.min()/max()
	__shared__ float part[32];
	float acc = MAX ? -INFINITY : INFINITY;
	for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
		float v = to_f<T>(in[i]);
		acc = MAX ? fmaxf(acc, v) : fminf(acc, v);
	}
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		float other = __shfl_xor_sync(0xffffffffu, acc, o);
		acc = MAX ? fmaxf(acc, other) : fminf(acc, other);
	}
	if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x < 32) {
		acc = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : (MAX ? -INFINITY : INFINITY);
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) {
			float other = __shfl_xor_sync(0xffffffffu, acc, o);
			acc = MAX ? fmaxf(acc, other) : fminf(acc, other);
		}
		if (threadIdx.x == 0) out[0] = from_f<T>(acc);
	}
}

}  // namespace

template <typename T>
__global__ void add2d_kernel(T* __restrict__ dst, int64_t dpitch, const T* __restrict__ src, int64_t spitch, int64_t width, int64_t rows)
{
	const int64_t total = width * rows;
	for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
		const int64_t r = i / width, c = i - r * width;
		T* d = dst + r * dpitch + c;
		*d = (T)((float)*d + (float)src[r * spitch + c]);
	}
}

extern "C" {

int pz_act_fwd(int kind, int dtype, void* out, const void* in, int64_t n, float a, float b, void* stream)
{
	PZ_REQUIRE(kind >= PZ_ACT_SIGMOID && kind <= PZ_ACT_GELU, "unknown activation kind %d", kind);
	if (kind == PZ_ACT_RELU) {
		PZ_DISPATCH_FLOAT(dtype, ew_launch<T, 1, 1, false>(out, nullptr, in, nullptr, nullptr, n, ReluFwd{}, stream));
	}
	PZ_DISPATCH_FLOAT(dtype, ew_launch<T, 1, 1, false>(out, nullptr, in, nullptr, nullptr, n, ActFwd{kind, a, b}, stream));
}

int pz_act_bwd(int kind, int dtype, void* ingrad, const void* outgrad, const void* ref, int64_t n, float a, float b,
			   void* stream)
{
	PZ_REQUIRE(kind >= PZ_ACT_SIGMOID && kind <= PZ_ACT_GELU, "unknown activation kind %d", kind);
	if (kind == PZ_ACT_RELU) {
		PZ_DISPATCH_FLOAT(dtype, ew_launch<T, 1, 2, false>(ingrad, nullptr, outgrad, ref, nullptr, n, ReluBwd{}, stream));
	}
	PZ_DISPATCH_FLOAT(dtype, ew_launch<T, 1, 2, false>(ingrad, nullptr, outgrad, ref, nullptr, n, ActBwd{kind, a, b}, stream));
}

int pz_act_fwd_slice(int kind, int dtype, void* out, const void* in, int64_t n, float a, float b, int64_t start, int64_t stop,
					 int64_t step, void* stream)
{
	PZ_REQUIRE(kind >= PZ_ACT_SIGMOID && kind <= PZ_ACT_GELU, "unknown activation kind %d", kind);
	PZ_DISPATCH_FLOAT(dtype, ew_launch_slice<T, 1, 1, false>(out, nullptr, in, nullptr, nullptr, n, start, stop, step, ActFwd{kind, a, b}, stream));
}

int pz_act_bwd_slice(int kind, int dtype, void* ingrad, const void* outgrad, const void* ref, int64_t n, float a, float b,
					 int64_t start, int64_t stop, int64_t step, void* stream)
{
	PZ_REQUIRE(kind >= PZ_ACT_SIGMOID && kind <= PZ_ACT_GELU, "unknown activation kind %d", kind);
	PZ_DISPATCH_FLOAT(dtype, ew_launch_slice<T, 1, 2, false>(ingrad, nullptr, outgrad, ref, nullptr, n, start, stop, step, ActBwd{kind, a, b}, stream));
}

int pz_axpby_slice(int dtype, void* out, const void* x, float alpha, const void* y, float beta, int64_t n, int64_t start, int64_t stop,
				   int64_t step, void* stream)
{
	PZ_DISPATCH_FLOAT(dtype, ew_launch_slice<T, 1, 2, false>(out, nullptr, x, y, nullptr, n, start, stop, step, Axpby{alpha, beta}, stream));
}

int pz_mul_slice(int dtype, void* out, const void* a, const void* b, int64_t n, int64_t start, int64_t stop, int64_t step, void* stream)
{
	PZ_DISPATCH_FLOAT(dtype, ew_launch_slice<T, 1, 2, false>(out, nullptr, a, b, nullptr, n, start, stop, step, Mul{}, stream));
}

int pz_axpy(int dtype, void* y, const void* x, float alpha, int64_t n, void* stream)
{
	PZ_DISPATCH_FLOAT(dtype, ew_launch<T, 1, 1, true>(y, nullptr, x, nullptr, nullptr, n, Axpy{alpha}, stream));
}

int pz_axpy2_relu(int dtype, void* y, void* out, const void* x1, float a1, const void* x2, float a2, int64_t n, void* stream)
{
	PZ_DISPATCH_FLOAT(dtype, ew_launch<T, 2, 2, false>(y, out, x1, x2, nullptr, n, Axpy2Relu<T>{a1, a2}, stream));
}

int pz_axpy2_relu_bwd(int dtype, void* y, void* ingrad, const void* x1, float a1, const void* x2, float a2, const void* ref, int64_t n,
					  void* stream)
{
	PZ_DISPATCH_FLOAT(dtype, ew_launch<T, 2, 3, false>(y, ingrad, x1, x2, ref, n, Axpy2ReluBwd<T>{a1, a2}, stream));
}

int pz_axpy2(int dtype, void* y, const void* x1, float a1, const void* x2, float a2, int64_t n, void* stream)
{
	if (x2 == nullptr) {
		PZ_DISPATCH_FLOAT(dtype, ew_launch<T, 1, 1, false>(y, nullptr, x1, nullptr, nullptr, n, Axpy2<T, false>{a1, 0.0f}, stream));
	}
	PZ_DISPATCH_FLOAT(dtype, ew_launch<T, 1, 2, false>(y, nullptr, x1, x2, nullptr, n, Axpy2<T, true>{a1, a2}, stream));
}

int pz_axpby(int dtype, void* out, const void* x, float alpha, const void* y, float beta, int64_t n, void* stream)
{
	PZ_DISPATCH_FLOAT(dtype, ew_launch<T, 1, 2, false>(out, nullptr, x, y, nullptr, n, Axpby{alpha, beta}, stream));
}

int pz_scale_shift(int dtype, void* out, const void* in, float a, float b, int64_t n, void* stream)
{
	PZ_DISPATCH_FLOAT(dtype, ew_launch<T, 1, 1, false>(out, nullptr, in, nullptr, nullptr, n, ScaleShift{a, b}, stream));
}

int pz_mul(int dtype, void* out, const void* a, const void* b, int64_t n, void* stream)
{
	PZ_DISPATCH_FLOAT(dtype, ew_launch<T, 1, 2, false>(out, nullptr, a, b, nullptr, n, Mul{}, stream));
}

int pz_add2(int dtype, void* out, const void* a, const void* b, int64_t n, void* stream)
{
	PZ_DISPATCH_FLOAT(dtype, ew_launch<T, 1, 2, false>(out, nullptr, a, b, nullptr, n, Add2{}, stream));
}

int pz_sgd_momentum(int dtype, void* param, const void* grad, void* mom, float lr, float mr, int64_t n, void* stream)
{
	PZ_DISPATCH_FLOAT(dtype, ew_launch<T, 2, 1, true>(param, mom, grad, nullptr, nullptr, n, SgdMom{lr, mr}, stream));
}

int pz_cast(int dd, void* dst, int sd, const void* src, int64_t n, void* stream)
{
	if (dd == sd) return pz_memcpy_d2d(dst, src, (size_t)n * pz_dtype_size(dd), stream);
#define PZ_CAST_CASE(D, TD, S, TS) if (dd == D && sd == S) return cast_launch<TD, TS>(dst, src, n, stream);
	PZ_CAST_CASE(PZ_F16, __half, PZ_F32, float)
	PZ_CAST_CASE(PZ_F32, float, PZ_F16, __half)
	PZ_CAST_CASE(PZ_BF16, __nv_bfloat16, PZ_F32, float)
	PZ_CAST_CASE(PZ_F32, float, PZ_BF16, __nv_bfloat16)
	PZ_CAST_CASE(PZ_F16, __half, PZ_BF16, __nv_bfloat16)
	PZ_CAST_CASE(PZ_BF16, __nv_bfloat16, PZ_F16, __half)
#undef PZ_CAST_CASE
	pz_set_error(PZ_ERR_UNSUPPORTED, "unsupported cast %d -> %d", sd, dd);
	return PZ_ERR_UNSUPPORTED;
}

// dst[r][i] += src[r][i] over `rows` rows of `width` elements with independent row pitches (in elements): the scatter-add of the
// 3-d transposed convolution's per-slice input gradients into dx[n, c, d0:d0+T] (dnn3d.py)
int pz_add2d(int dtype, void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width, int64_t rows, void* stream)
{
	if (width <= 0 || rows <= 0) return PZ_OK;
	PZ_REQUIRE(width * rows < (1ll << 40), "add2d: too large");
	int64_t blocks = pz_cdiv(width * rows, 256);
	if (blocks > (int64_t)pz_num_sms() * 16) blocks = (int64_t)pz_num_sms() * 16;
	switch (dtype) {
		case PZ_F32: add2d_kernel<float><<<(unsigned)blocks, 256, 0, pz_stream(stream)>>>((float*)dst, dpitch, (const float*)src, spitch, width, rows); break;
		case PZ_F16: add2d_kernel<__half><<<(unsigned)blocks, 256, 0, pz_stream(stream)>>>((__half*)dst, dpitch, (const __half*)src, spitch, width, rows); break;
		case PZ_BF16: add2d_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, pz_stream(stream)>>>((__nv_bfloat16*)dst, dpitch, (const __nv_bfloat16*)src, spitch, width, rows); break;
		default: pz_set_error(PZ_ERR_UNSUPPORTED, "unsupported dtype %d", dtype); return PZ_ERR_UNSUPPORTED;
	}
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

int pz_reduce_minmax(int dtype, const void* in, int64_t n, int want_max, void* out, void* stream)
{
	PZ_REQUIRE(n > 0, "empty reduction");
	switch (dtype) {
		case PZ_F32:
			if (want_max) minmax_kernel<float, true><<<1, 1024, 0, pz_stream(stream)>>>((const float*)in, n, (float*)out);
			else minmax_kernel<float, false><<<1, 1024, 0, pz_stream(stream)>>>((const float*)in, n, (float*)out);
			break;
		case PZ_F16:
			if (want_max) minmax_kernel<__half, true><<<1, 1024, 0, pz_stream(stream)>>>((const __half*)in, n, (__half*)out);
			else minmax_kernel<__half, false><<<1, 1024, 0, pz_stream(stream)>>>((const __half*)in, n, (__half*)out);
			break;
		case PZ_BF16:
			if (want_max) minmax_kernel<__nv_bfloat16, true><<<1, 1024, 0, pz_stream(stream)>>>((const __nv_bfloat16*)in, n, (__nv_bfloat16*)out);
			else minmax_kernel<__nv_bfloat16, false><<<1, 1024, 0, pz_stream(stream)>>>((const __nv_bfloat16*)in, n, (__nv_bfloat16*)out);
			break;
		default:
			pz_set_error(PZ_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
			return PZ_ERR_UNSUPPORTED;
	}
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

}  // extern "C"
