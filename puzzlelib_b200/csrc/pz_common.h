// pz_common.h -- shared helpers for the libpzb200 C-ABI library (sm_100a only).
//
// Error model: every exported entry point returns an int status (0 = OK) and
// leaves a thread-local message retrievable through pz_last_error(); the
// ctypes shim raises the reference's exception classes from it
// (reference: Cuda/Source/Core/Common.h:119-139, Libs/Libs.h:30-44).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/pzb200.h"

void pz_set_error(int code, const char* fmt, ...);

#define PZ_CHECK_CUDA(expr)                                                                      \
	do {                                                                                         \
		cudaError_t _e = (expr);                                                                 \
		if (_e != cudaSuccess) {                                                                 \
			pz_set_error(PZ_ERR_CUDA, "%s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
			return PZ_ERR_CUDA;                                                                  \
		}                                                                                        \
	} while (0)

#define PZ_REQUIRE(cond, ...)                     \
	do {                                          \
		if (!(cond)) {                            \
			pz_set_error(PZ_ERR_VALUE, __VA_ARGS__); \
			return PZ_ERR_VALUE;                  \
		}                                         \
	} while (0)

#define PZ_LAUNCH_CHECK()                                                                        \
	do {                                                                                         \
		cudaError_t _e = cudaGetLastError();                                                     \
		if (_e != cudaSuccess) {                                                                 \
			pz_set_error(PZ_ERR_CUDA, "%s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
			return PZ_ERR_CUDA;                                                                  \
		}                                                                                        \
	} while (0)

// a NULL stream argument means "the library's current stream": the legacy default stream (what the reference uses for
// everything) unless pz_set_default_stream() redirected it, e.g. to a capturing stream (pz_graph_begin)
cudaStream_t pz_stream(void* s);
void pz_norm_graph_begin(cudaStream_t stream);      // batch-norm accumulator reset hook (pz_norm.cu)

int pz_num_sms();
// library-owned device scratch, grown on demand and used in stream order (prepared conv filters, pooling winner maps);
// nullptr when the allocation fails
void* pz_scratch(size_t bytes);
void pz_count_launch(int n);

// ---- optional per-launch profiling (bench.py roofline): CUDA events around the launches of one kernel family
// PZ_PROF_GEMM: tcgen05 launches whose arithmetic intensity (algorithmic flops / bytes) is above the machine ridge -- bounded by
// the tensor pipe; PZ_PROF_GEMM_HBM: the ones below it (ResNet's 1x1 convolutions) -- bounded by HBM
enum { PZ_PROF_GEMM = 0, PZ_PROF_BN_FWD = 1, PZ_PROF_BN_BWD = 2, PZ_PROF_ELTWISE = 3, PZ_PROF_POOL = 4, PZ_PROF_OTHER = 5,
	   PZ_PROF_GEMM_HBM = 6, PZ_PROF_FAMILIES = 7 };
constexpr double kPzRidgeFlopPerByte = 108.0;     // ~709 TFLOP/s tf32 / 6.55 TB/s measured on this pool's B200s
bool pz_prof_on();
void pz_prof_begin(int family, cudaStream_t stream, double flops, double bytes);
void pz_prof_end(cudaStream_t stream);

struct PzProfScope {
	cudaStream_t stream;
	bool on;
	PzProfScope(int family, cudaStream_t s, double flops, double bytes) : stream(s), on(pz_prof_on())
	{
		if (on) pz_prof_begin(family, s, flops, bytes);
	}
	~PzProfScope()
	{
		if (on) pz_prof_end(stream);
	}
};

static inline int64_t pz_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

static inline size_t pz_dtype_size(int dtype)
{
	switch (dtype) {
		case PZ_F32: case PZ_I32: case PZ_U32: return 4;
		case PZ_F16: case PZ_BF16: case PZ_I16: case PZ_U16: return 2;
		case PZ_I8: case PZ_U8: return 1;
		case PZ_F64: case PZ_I64: case PZ_U64: return 8;
		default: return 0;
	}
}
