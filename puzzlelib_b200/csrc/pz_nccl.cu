// pz_nccl.cu -- data-parallel gradient synchronisation over NCCL (NVLink 5 / NVSwitch inside one 8-GPU box).
//
// Replaces the reference's Grid.py data plane: a parent/child STAR over CUDA-IPC mapped buffers where the parent
// adds each child's flat gradient buffer with an axpy kernel and every child copies the result back
// (reference Grid.py:66-157; Cuda/Source/Core/Buffer.c:411-424).  Here every rank calls one ncclAllReduce(sum)
// on the same flat per-dtype gradient buffer and the 1/P scale is folded into the kernel that consumes the
// gradient (the momentum-SGD update that follows it in Optimizer.updateGlobalState, Optimizers/Optimizer.py:159-170).
//
// libnccl is opened lazily with dlopen(RTLD_LOCAL) so that single-GPU users never need it and so that a test
// process that also imports torch (which bundles its own libnccl) cannot get its symbols mixed with ours.
#include "pz_common.h"

#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
	void* handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(ncclResult_t) = nullptr;
	ncclResult_t (*GetVersion)(int*) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
};

NcclApi g_api;

int load_api()
{
	if (g_api.handle) return PZ_OK;
	const char* names[] = {getenv("PZB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
	void* h = nullptr;
	for (const char* n : names) {
		if (!n || !*n) continue;
		h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
		if (h) break;
	}
	if (!h) {
		pz_set_error(PZ_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
		return PZ_ERR_NCCL;
	}
#define PZ_SYM(field, name)                                                    \
	*(void**)(&g_api.field) = dlsym(h, name);                                  \
	if (!g_api.field) {                                                        \
		pz_set_error(PZ_ERR_NCCL, "libnccl is missing symbol %s", name);       \
		dlclose(h);                                                            \
		return PZ_ERR_NCCL;                                                    \
	}
	PZ_SYM(GetUniqueId, "ncclGetUniqueId")
	PZ_SYM(CommInitRank, "ncclCommInitRank")
	PZ_SYM(CommDestroy, "ncclCommDestroy")
	PZ_SYM(AllReduce, "ncclAllReduce")
	PZ_SYM(Broadcast, "ncclBroadcast")
	PZ_SYM(GetErrorString, "ncclGetErrorString")
	PZ_SYM(GetVersion, "ncclGetVersion")
	PZ_SYM(GroupStart, "ncclGroupStart")
	PZ_SYM(GroupEnd, "ncclGroupEnd")
#undef PZ_SYM
	g_api.handle = h;
	return PZ_OK;
}

#define PZ_CHECK_NCCL(expr)                                                                              \
	do {                                                                                                 \
		ncclResult_t _r = (expr);                                                                        \
		if (_r != ncclSuccess) {                                                                         \
			pz_set_error(PZ_ERR_NCCL, "%s (%s:%d)", g_api.GetErrorString(_r), __FILE__, __LINE__);       \
			return PZ_ERR_NCCL;                                                                          \
		}                                                                                                \
	} while (0)

int nccl_dtype(int dtype, ncclDataType_t* out)
{
	switch (dtype) {
		case PZ_F32: *out = ncclFloat32; return PZ_OK;
		case PZ_F16: *out = ncclFloat16; return PZ_OK;
		case PZ_BF16: *out = ncclBfloat16; return PZ_OK;
		case PZ_F64: *out = ncclFloat64; return PZ_OK;
		case PZ_I32: *out = ncclInt32; return PZ_OK;
		case PZ_I64: *out = ncclInt64; return PZ_OK;
		case PZ_U8: *out = ncclUint8; return PZ_OK;
		default:
			pz_set_error(PZ_ERR_UNSUPPORTED, "nccl: unsupported dtype %d", dtype);
			return PZ_ERR_UNSUPPORTED;
	}
}

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T> struct alignas(16) Pack { T v[16 / sizeof(T)]; };

// grad <- grad * scale (the mean of Grid.sumTensor); mom <- mr * mom + lr * grad; param <- param + mom.
// One pass: 3 reads + 3 writes of 128-bit words per element group.
template <typename T>
__global__ void __launch_bounds__(256) mean_sgd_kernel(T* __restrict__ param, T* __restrict__ grad, T* __restrict__ mom, int64_t n,
													   float scale, float lr, float mr)
{
	constexpr int V = 16 / sizeof(T);
	const int64_t nvec = n / V;
	const int64_t tid = (int64_t)blockIdx.x * 256 + threadIdx.x, nth = (int64_t)gridDim.x * 256;
	for (int64_t i = tid; i < nvec; i += nth) {
		Pack<T> p = reinterpret_cast<Pack<T>*>(param)[i], g = reinterpret_cast<Pack<T>*>(grad)[i], m = reinterpret_cast<Pack<T>*>(mom)[i];
		#pragma unroll
		for (int e = 0; e < V; e++) {
			const T gs = from_f<T>(to_f<T>(g.v[e]) * scale);      // stored mean gradient, rounded like the reference's addKer output
			const float mv = mr * to_f<T>(m.v[e]) + lr * to_f<T>(gs);
			g.v[e] = gs;
			m.v[e] = from_f<T>(mv);
			p.v[e] = from_f<T>(to_f<T>(p.v[e]) + to_f<T>(m.v[e]));
		}
		reinterpret_cast<Pack<T>*>(param)[i] = p;
		reinterpret_cast<Pack<T>*>(grad)[i] = g;
		reinterpret_cast<Pack<T>*>(mom)[i] = m;
	}
	for (int64_t i = nvec * V + tid; i < n; i += nth) {
		const T gs = from_f<T>(to_f<T>(grad[i]) * scale);
		const T mv = from_f<T>(mr * to_f<T>(mom[i]) + lr * to_f<T>(gs));
		grad[i] = gs;
		mom[i] = mv;
		param[i] = from_f<T>(to_f<T>(param[i]) + to_f<T>(mv));
	}
}

template <typename T>
int mean_sgd_launch(void* param, void* grad, void* mom, int64_t n, float scale, float lr, float mr, void* stream)
{
	if (n <= 0) return PZ_OK;
	constexpr int V = 16 / sizeof(T);
	int64_t blocks = pz_cdiv(pz_cdiv(n, V), 256);
	const int64_t cap = (int64_t)pz_num_sms() * 8;
	if (blocks > cap) blocks = cap;
	mean_sgd_kernel<T><<<(unsigned)blocks, 256, 0, pz_stream(stream)>>>((T*)param, (T*)grad, (T*)mom, n, scale, lr, mr);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

}  // namespace

extern "C" {

int pz_nccl_version(int* version)
{
	int st = load_api();
	if (st != PZ_OK) return st;
	PZ_CHECK_NCCL(g_api.GetVersion(version));
	return PZ_OK;
}

int pz_nccl_unique_id(void* id128)
{
	int st = load_api();
	if (st != PZ_OK) return st;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
	ncclUniqueId id;
	PZ_CHECK_NCCL(g_api.GetUniqueId(&id));
	memcpy(id128, &id, sizeof(id));
	return PZ_OK;
}

int pz_nccl_comm_init(void** comm, int nranks, int rank, const void* id128)
{
	int st = load_api();
	if (st != PZ_OK) return st;
	PZ_REQUIRE(nranks > 0 && rank >= 0 && rank < nranks, "nccl: invalid rank %d of %d", rank, nranks);
	ncclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	ncclComm_t c;
	PZ_CHECK_NCCL(g_api.CommInitRank(&c, nranks, id, rank));
	*comm = (void*)c;
	return PZ_OK;
}

int pz_nccl_comm_destroy(void* comm)
{
	if (!comm) return PZ_OK;
	int st = load_api();
	if (st != PZ_OK) return st;
	PZ_CHECK_NCCL(g_api.CommDestroy((ncclComm_t)comm));
	return PZ_OK;
}

int pz_nccl_allreduce_mean(void* comm, int dtype, void* buf, int64_t count, float scale, void* stream)
{
	int st = load_api();
	if (st != PZ_OK) return st;
	ncclDataType_t dt;
	st = nccl_dtype(dtype, &dt);
	if (st != PZ_OK) return st;
	if (count <= 0) return PZ_OK;
	PZ_CHECK_NCCL(g_api.AllReduce(buf, buf, (size_t)count, dt, ncclSum, (ncclComm_t)comm, pz_stream(stream)));
	if (scale != 1.0f) return pz_scale_shift(dtype, buf, buf, scale, 0.0f, count, stream);
	return PZ_OK;
}

// One bucket of the overlapped gradient synchronisation: the mean over the ranks of `nseg` disjoint segments of the flat
// gradient buffer, as ONE grouped NCCL launch (ncclAvg) on `stream` -- the communication stream of grid.GradientSync, which
// runs it while the compute stream is still producing the gradients of the earlier layers.
int pz_nccl_allreduce_avg_segments(void* comm, int dtype, void* const* ptrs, const int64_t* counts, int nseg, void* stream)
{
	int st = load_api();
	if (st != PZ_OK) return st;
	ncclDataType_t dt;
	st = nccl_dtype(dtype, &dt);
	if (st != PZ_OK) return st;
	PZ_REQUIRE(nseg >= 0 && (nseg == 0 || (ptrs != nullptr && counts != nullptr)), "nccl: bad segment list");
	PZ_REQUIRE(stream != nullptr, "nccl: the segment all-reduce needs an explicit stream");
	if (nseg == 0) return PZ_OK;
	PZ_CHECK_NCCL(g_api.GroupStart());
	for (int i = 0; i < nseg; i++) {
		if (counts[i] <= 0) continue;
		ncclResult_t r = g_api.AllReduce(ptrs[i], ptrs[i], (size_t)counts[i], dt, ncclAvg, (ncclComm_t)comm, (cudaStream_t)stream);
		if (r != ncclSuccess) {
			g_api.GroupEnd();
			pz_set_error(PZ_ERR_NCCL, "%s (%s:%d)", g_api.GetErrorString(r), __FILE__, __LINE__);
			return PZ_ERR_NCCL;
		}
	}
	PZ_CHECK_NCCL(g_api.GroupEnd());
	pz_count_launch(1);
	return PZ_OK;
}

int pz_nccl_broadcast(void* comm, int dtype, void* buf, int64_t count, int root, void* stream)
{
	int st = load_api();
	if (st != PZ_OK) return st;
	ncclDataType_t dt;
	st = nccl_dtype(dtype, &dt);
	if (st != PZ_OK) return st;
	if (count <= 0) return PZ_OK;
	PZ_CHECK_NCCL(g_api.Broadcast(buf, buf, (size_t)count, dt, root, (ncclComm_t)comm, pz_stream(stream)));
	return PZ_OK;
}

int pz_mean_sgd_momentum(int dtype, void* param, void* grad, void* mom, int64_t count, float scale, float learn_rate,
						 float mom_rate, void* stream)
{
	switch (dtype) {
		case PZ_F32: return mean_sgd_launch<float>(param, grad, mom, count, scale, learn_rate, mom_rate, stream);
		case PZ_F16: return mean_sgd_launch<__half>(param, grad, mom, count, scale, learn_rate, mom_rate, stream);
		case PZ_BF16: return mean_sgd_launch<__nv_bfloat16>(param, grad, mom, count, scale, learn_rate, mom_rate, stream);
		default:
			pz_set_error(PZ_ERR_UNSUPPORTED, "unsupported dtype %d", dtype);
			return PZ_ERR_UNSUPPORTED;
	}
}

int pz_nccl_allreduce_sgd_momentum(void* comm, int dtype, void* param, void* grad, void* mom, int64_t count, float scale,
								   float learn_rate, float mom_rate, void* stream)
{
	int st = load_api();
	if (st != PZ_OK) return st;
	ncclDataType_t dt;
	st = nccl_dtype(dtype, &dt);
	if (st != PZ_OK) return st;
	if (count <= 0) return PZ_OK;
	PZ_CHECK_NCCL(g_api.AllReduce(grad, grad, (size_t)count, dt, ncclSum, (ncclComm_t)comm, pz_stream(stream)));
	return pz_mean_sgd_momentum(dtype, param, grad, mom, count, scale, learn_rate, mom_rate, stream);
}

}  // extern "C"
