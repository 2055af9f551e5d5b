// Micro-benchmark: how much global-load bandwidth does one SM sustain from a persistent 512-thread CTA as a function of
// bytes per lane (4 / 8 / 16) and loads in flight per thread?  Answers whether the conv producers (4-byte lanes, 128-byte
// requests) are bound by the number of outstanding requests per SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <typename V, int UNROLL>
__global__ void __launch_bounds__(1024, 1) read_kernel(const V* __restrict__ src, size_t nvec, float* out, int misalign)
{
	const char* base = (const char*)src + misalign;
	const size_t stride = (size_t)gridDim.x * blockDim.x;
	float acc = 0.f;
	size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	for (; i + (UNROLL - 1) * stride < nvec; i += UNROLL * stride) {
		V v[UNROLL];
		#pragma unroll
		for (int u = 0; u < UNROLL; u++) v[u] = __ldg((const V*)(base + (i + u * stride) * sizeof(V)));
		#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
			const float* f = (const float*)&v[u];
			#pragma unroll
			for (int e = 0; e < (int)(sizeof(V) / 4); e++) acc += f[e];
		}
	}
	if (acc == 123.456f) out[0] = acc;
}

template <typename V, int UNROLL>
void run(const char* name, const void* buf, size_t bytes, float* out, int threads, int misalign)
{
	const size_t nvec = (bytes - 64) / sizeof(V);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	read_kernel<V, UNROLL><<<148, threads>>>((const V*)buf, nvec, out, misalign);
	cudaEventRecord(e0);
	for (int r = 0; r < 5; r++) read_kernel<V, UNROLL><<<148, threads>>>((const V*)buf, nvec, out, misalign);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	printf("%-28s threads/SM %4d  misalign %d : %7.1f GB/s  (%.1f GB/s per SM)\n", name, threads, misalign, bytes * 5 / ms / 1e6,
		   bytes * 5 / ms / 1e6 / 148);
}

int main()
{
	const size_t bytes = 1ull << 30;
	void* buf;
	float* out;
	cudaMalloc(&buf, bytes);
	cudaMalloc(&out, 4);
	cudaMemset(buf, 0, bytes);
	for (int threads : {512, 1024}) {
		for (int mis : {0, 4}) {
			run<float, 8>("4 B/lane, 8 in flight", buf, bytes, out, threads, mis);
			run<float, 16>("4 B/lane, 16 in flight", buf, bytes, out, threads, mis);
			run<float, 32>("4 B/lane, 32 in flight", buf, bytes, out, threads, mis);
		}
		run<float2, 8>("8 B/lane, 8 in flight", buf, bytes, out, threads, 0);
		run<float2, 16>("8 B/lane, 16 in flight", buf, bytes, out, threads, 0);
		run<float4, 2>("16 B/lane, 2 in flight", buf, bytes, out, threads, 0);
		run<float4, 4>("16 B/lane, 4 in flight", buf, bytes, out, threads, 0);
		run<float4, 8>("16 B/lane, 8 in flight", buf, bytes, out, threads, 0);
		run<float4, 16>("16 B/lane, 16 in flight", buf, bytes, out, threads, 0);
	}
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
