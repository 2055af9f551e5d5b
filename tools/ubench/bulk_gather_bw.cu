// Micro-benchmark: the gather of the 1x1 convolution producers done by the copy engine instead of the LSU.  A persistent CTA per SM
// walks over tiles of 128 consecutive positions; per k-block ONE warp issues 32 cp.async.bulk copies (one 512..544-byte run per
// channel plane, the 16-byte aligned cover of the run) into a ring of NSLOT staging buffers; the other warps wait for a slot,
// transpose it into a [128 positions][32 channels] tile in shared memory (what the MMA wants) and release it.
// Question: does the copy engine sustain more than the ~3.3 TB/s the LSU gather gives the engine (5.0 TB/s in plane_gather_bw.cu)?
// build + run on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bg bulk_gather_bw.cu && /tmp/bg
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int ROWS = 128, STRIDE = 544, SLOT_BYTES = 32 * STRIDE, NCONS = 15;

template <int NSLOT>
__global__ void __launch_bounds__(512, 1) bulk_gather_kernel(const float* __restrict__ x, int plane, int chans, int images, float* out)
{
	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t* slots = smem;                                     // [NSLOT][32][STRIDE]
	float* tile = (float*)(smem + NSLOT * SLOT_BYTES);         // [2][128][32] transposed output (stands for the MMA stage ring)
	__shared__ uint64_t full[NSLOT], empty[NSLOT];
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	if (threadIdx.x == 0) {
		for (int i = 0; i < NSLOT; i++) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), NCONS); }
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	const long long rows_total = (long long)images * plane;
	const int tiles = (int)(rows_total / ROWS), cblocks = chans / 32;
	float acc = 0.f;
	if (warp == 0) {
		// loader: k-block j of this CTA -> slot j % NSLOT
		for (long long j = 0;; j++) {
			const int t = blockIdx.x + (int)(j / cblocks) * gridDim.x;
			if (t >= tiles) break;
			const int cb = (int)(j % cblocks), slot = (int)(j % NSLOT);
			if (j >= NSLOT) mbar_wait(smem_u32(&empty[slot]), (uint32_t)((j / NSLOT - 1) & 1));
			const long long row0 = (long long)t * ROWS;
			const int n = (int)(row0 / plane);
			const int pos = (int)(row0 - (long long)n * plane);
			// lane = channel: the run of 128 positions (clipped to the image; the rest of a tile that crosses an image is ignored here)
			const long long e0 = ((long long)n * chans + cb * 32 + lane) * plane + pos;
			const int len = min(ROWS, plane - pos);
			const long long a0 = e0 & ~3ll, a1 = (e0 + len + 3) & ~3ll;
			const uint32_t bytes = (uint32_t)(a1 - a0) * 4u;
			mbar_expect(smem_u32(&full[slot]), bytes);
			bulk_g2s(smem_u32(slots + slot * SLOT_BYTES + lane * STRIDE), x + a0, bytes, smem_u32(&full[slot]));
			__syncwarp();
			if (lane == 0) mbar_arrive(smem_u32(&full[slot]));
		}
	} else {
		const int cw = warp - 1;                               // consumer warp 0..14
		for (long long j = 0;; j++) {
			const int t = blockIdx.x + (int)(j / cblocks) * gridDim.x;
			if (t >= tiles) break;
			const int slot = (int)(j % NSLOT);
			mbar_wait(smem_u32(&full[slot]), (uint32_t)((j / NSLOT) & 1));
			// transpose: element (channel c, position r) -> tile[r][c]; 4096 elements over 480 threads
			const float* s = (const float*)(slots + slot * SLOT_BYTES);
			float* d = tile + (j & 1) * 4096;
			for (int i = cw * 32 + lane; i < 1024; i += NCONS * 32) {
				const int r = i & 127, c4 = i >> 7;            // 128 positions x 8 chunks of 4 channels
				const float4 v = make_float4(s[(c4 * 4 + 0) * (STRIDE / 4) + r], s[(c4 * 4 + 1) * (STRIDE / 4) + r], s[(c4 * 4 + 2) * (STRIDE / 4) + r],
											 s[(c4 * 4 + 3) * (STRIDE / 4) + r]);
				*reinterpret_cast<float4*>(d + r * 32 + ((c4 ^ (r & 7)) << 2)) = v;
			}
			__syncwarp();
			if (lane == 0) mbar_arrive(smem_u32(&empty[slot]));
			acc += d[threadIdx.x & 1023];
		}
	}
	if (acc == 123.456f) out[0] = acc;
}

template <int NSLOT>
void run(const float* x, int plane, int chans, int images, float* out)
{
	const int smem = NSLOT * SLOT_BYTES + 2 * 4096 * 4;
	cudaFuncSetAttribute(bulk_gather_kernel<NSLOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	bulk_gather_kernel<NSLOT><<<148, 512, smem>>>(x, plane, chans, images, out);
	cudaEventRecord(e0);
	for (int r = 0; r < 5; r++) bulk_gather_kernel<NSLOT><<<148, 512, smem>>>(x, plane, chans, images, out);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double bytes = (double)images * chans * plane * 4.0;
	printf("plane %5d chans %4d  %d slots (%3d KB in flight / SM): %7.1f GB/s (%.1f GB/s per SM)  %s\n", plane, chans, NSLOT, NSLOT * SLOT_BYTES / 1024,
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
	const int shapes[][2] = {{3025, 256}, {3025, 64}, {784, 512}, {196, 1024}};
	for (auto& s : shapes) {
		run<2>(buf, s[0], s[1], 64, out);
		run<4>(buf, s[0], s[1], 64, out);
		run<6>(buf, s[0], s[1], 64, out);
		run<8>(buf, s[0], s[1], 64, out);
		run<10>(buf, s[0], s[1], 64, out);
	}
	return 0;
}
