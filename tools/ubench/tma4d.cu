// One 4-d tiled TMA load with parameters from the command line: which box shapes / coordinates / swizzles the copy engine accepts.
//   tma4d W H C N  bx by bc  c0 c1 c2 c3  swizzle(0 none, 3 = 128B)  [es = 4]
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma4d tma4d.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__global__ void k(const __grid_constant__ CUtensorMap tm, int c0, int c1, int c2, int c3, unsigned bytes, float* out, int nout)
{
	extern __shared__ __align__(1024) unsigned char smem[];
	__shared__ __align__(8) unsigned long long bar;
	unsigned sb = (unsigned)__cvta_generic_to_shared(&bar);
	unsigned dst = ((unsigned)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(bytes) : "memory");
		asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
					 ::"r"(dst), "l"((unsigned long long)&tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(sb) : "memory");
	}
	unsigned ok = 0;
	for (int it = 0; it < 2000000 && !ok; it++)
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(sb) : "memory");
	__syncthreads();
	const float* s = (const float*)(smem + (dst - (unsigned)__cvta_generic_to_shared(smem)));
	for (int i = threadIdx.x; i < nout; i += blockDim.x) out[i] = ok ? s[i] : -12345.0f;
}

int main(int argc, char** argv)
{
	if (argc < 13) { printf("usage\n"); return 2; }
	long W = atol(argv[1]), H = atol(argv[2]), C = atol(argv[3]), N = atol(argv[4]);
	unsigned bx = atoi(argv[5]), by = atoi(argv[6]), bc = atoi(argv[7]);
	int c0 = atoi(argv[8]), c1 = atoi(argv[9]), c2 = atoi(argv[10]), c3 = atoi(argv[11]);
	int sw = atoi(argv[12]);
	size_t total = (size_t)W * H * C * N;
	std::vector<float> h(total);
	for (size_t i = 0; i < total; i++) h[i] = (float)(i % 1000003);
	float* d; cudaMalloc(&d, total * 4); cudaMemcpy(d, h.data(), total * 4, cudaMemcpyHostToDevice);
	typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
								 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
	void* fn = nullptr; cudaDriverEntryPointQueryResult q;
	cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
	CUtensorMap tm; memset(&tm, 0, sizeof(tm));
	cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
	cuuint64_t str[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
	cuuint32_t box[4] = {bx, by, bc, 1}, es[4] = {1, 1, 1, 1};
	CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)sw,
								CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	printf("dims %ldx%ldx%ldx%ld box %ux%ux%u coords %d,%d,%d,%d swizzle %d: encode %d", W, H, C, N, bx, by, bc, c0, c1, c2, c3, sw, (int)r);
	if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
	unsigned bytes = bx * by * bc * 4;
	int nout = bytes / 4;
	float* out; cudaMalloc(&out, bytes);
	cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
	k<<<1, 128, bytes + 2048>>>(tm, c0, c1, c2, c3, bytes, out, nout);
	cudaError_t e = cudaDeviceSynchronize();
	printf(" run: %s", cudaGetErrorString(e));
	if (e == cudaSuccess) {
		std::vector<float> o(nout); cudaMemcpy(o.data(), out, bytes, cudaMemcpyDeviceToHost);
		// expected first element: the tensor value at the box origin (0 if outside)
		auto at = [&](long w, long hh, long c, long n) -> float {
			if (w < 0 || w >= W || hh < 0 || hh >= H || c < 0 || c >= C || n < 0 || n >= N) return 0.0f;
			return h[((n * C + c) * H + hh) * W + w]; };
		long bad = 0;
		if (sw == 0) {
			for (unsigned c = 0; c < bc; c++) for (unsigned y = 0; y < by; y++) for (unsigned x = 0; x < bx; x++)
				if (o[(c * by + y) * bx + x] != at(c0 + x, c1 + y, c2 + c, c3)) bad++;
		} else {
			// 128-byte swizzle of a dense box: 16-byte chunk index ^= (128-byte line index & 7)
			for (unsigned c = 0; c < bc; c++) for (unsigned y = 0; y < by; y++) for (unsigned x = 0; x < bx; x++) {
				unsigned lin = ((c * by + y) * bx + x) * 4, line = lin >> 7, chunk = (lin >> 4) & 7, within = lin & 15;
				unsigned addr = (line << 7) | (((chunk ^ (line & 7)) & 7) << 4) | within;
				if (o[addr / 4] != at(c0 + x, c1 + y, c2 + c, c3)) bad++;
			}
		}
		printf(" first %.0f %.0f %.0f mismatches %ld of %d", o[0], o[1], o[2], bad, nout);
	}
	printf("\n");
	return 0;
}
