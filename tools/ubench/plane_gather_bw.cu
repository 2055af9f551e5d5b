// Micro-benchmark: the access pattern of the conv producers without the GEMM around it.  A persistent 512-thread CTA per SM walks
// over "tiles" of ROWS consecutive positions; per k-block it reads 32 channel planes x ROWS positions (4-byte lanes, 128 bytes
// per warp request, planes `plane` floats apart) with all 32 loads of a thread in flight, then stores them to shared memory.
// Question: is 3.3 TB/s the pattern's own ceiling (DRAM page locality of 512-byte runs) or the kernel's?
// build + run on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pg plane_gather_bw.cu && /tmp/pg
#include <cstdio>
#include <cuda_runtime.h>

template <int ROWS>
__global__ void __launch_bounds__(512, 1) gather_kernel(const float* __restrict__ x, int plane, int chans, int images, float* out)
{
	__shared__ float tile[4][2048];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int grp = warp >> 2, gw = warp & 3;                    // 4 groups of 4 warps, a k-block per group
	const long long rows_total = (long long)images * plane;
	const int tiles = (int)(rows_total / ROWS);
	const int cblocks = chans / 32;
	float acc = 0.f;
	// group g takes k-blocks g, g+4, ... of the CTA's (tile, channel block) sequence
	for (long long j = grp;; j += 4) {
		const int t = blockIdx.x + (int)(j / cblocks) * gridDim.x;
		if (t >= tiles) break;
		const int cb = (int)(j % cblocks);
		const long long row0 = (long long)t * ROWS;
		const int n = (int)(row0 / plane);
		const int pos = (int)(row0 - (long long)n * plane);
		// thread: rows gw*32 + lane (+128 per extra row block), channels cb*32 .. +31
		constexpr int RPT = ROWS / 128;
		float v[32 * RPT];
		#pragma unroll
		for (int r = 0; r < RPT; r++) {
			const float* p = x + ((long long)n * chans + cb * 32) * plane + pos + r * 128 + gw * 32 + lane;
			#pragma unroll
			for (int c = 0; c < 32; c++)
				asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v[r * 32 + c]) : "l"(p + (long long)c * plane));
		}
		#pragma unroll
		for (int i = 0; i < 32 * RPT; i++)
			asm volatile("st.shared.f32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&tile[grp][(i * 128 + gw * 32 + lane) & 2047])), "f"(v[i]) : "memory");
		acc += tile[grp][threadIdx.x & 127];
	}
	if (acc == 123.456f) out[0] = acc;
}

template <int ROWS>
void run(const float* x, int plane, int chans, int images, float* out)
{
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	gather_kernel<ROWS><<<148, 512>>>(x, plane, chans, images, out);
	cudaEventRecord(e0);
	for (int r = 0; r < 5; r++) gather_kernel<ROWS><<<148, 512>>>(x, plane, chans, images, out);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double bytes = (double)images * chans * plane * 4.0;
	printf("plane %5d chans %4d rows/tile %3d (%4d B runs): %7.1f GB/s (%.1f GB/s per SM)  %s\n", plane, chans, ROWS, ROWS * 4,
		   bytes * 5 / ms / 1e6, bytes * 5 / ms / 1e6 / 148, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
	const size_t bytes = 1ull << 30;
	float* buf;
	float* out;
	cudaMalloc(&buf, bytes);
	cudaMalloc(&out, 4);
	cudaMemset(buf, 0, bytes);
	// ResNet-50 shapes at N = 64: (plane, channels)
	const int shapes[][2] = {{3025, 256}, {3025, 64}, {784, 512}, {196, 1024}, {3072, 256}};
	for (auto& s : shapes) {
		run<128>(buf, s[0], s[1], 64, out);
		run<256>(buf, s[0], s[1], 64, out);
	}
	return 0;
}
