// Micro-benchmark: which property of the conv epilogue's store stream limits the write-heavy 1x1 layers?
// A persistent CTA per SM with W storing warps writes "tiles" of 128 positions x 256 channel planes of an NCHW tensor
// (plane = 55*55 floats, like ResNet-50's 64 -> 256 layer at batch 64): per tile and channel either
//   mode 0: four warps each store 32 consecutive floats   (128-byte requests, 4-byte lanes)   <- the engine's epilogue today
//   mode 1: one warp stores 128 consecutive floats as 16-byte lanes (512-byte requests; needs 16-byte aligned rows: plane 56*56)
//   mode 2: one warp stores 128 consecutive floats as four 4-byte-lane instructions in a row (same warp, adjacent 128-byte requests)
//   mode 4: mode 0 with the four warps of a row kept in lockstep by a named barrier every 16 channels
// and, for reference, mode 3: the same bytes written as one contiguous stream (what an elementwise kernel does).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_pattern store_pattern.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) store_kernel(float* __restrict__ out, int mode, int plane, int channels, long long tiles, int images)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
	const int tiles_per_image = (plane + 127) / 128;
	for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
		const int n = (int)(t / tiles_per_image), p0 = (int)(t % tiles_per_image) * 128;
		float* img = out + (long long)n * channels * plane;
		if (mode == 3) {
			// contiguous: this tile's share of the tensor as one run
			float4* dst = (float4*)(out + t * 128ll * channels);
			for (int i = threadIdx.x; i < 32 * channels; i += blockDim.x) dst[i] = make_float4(1.f, 2.f, 3.f, 4.f);
			continue;
		}
		if (mode == 4) {
			// mode 0 with the four quadrant warps of a row kept in lockstep: a named barrier every 16 channels
			const int q = warp & 3, team = warp >> 2;
			int k = 0;
			for (int c = team; c < channels; c += nwarps >> 2) {
				const int p = p0 + q * 32 + lane;
				if (p < plane) img[(long long)c * plane + p] = 1.f;
				if ((++k & 15) == 0) asm volatile("bar.sync %0, 128;" ::"r"(1 + team));
			}
			continue;
		}
		if (mode == 0) {
			// warp w: lane quadrant w % 4 (32 positions), channels w / 4, w / 4 + nwarps / 4, ...
			const int q = warp & 3;
			for (int c = warp >> 2; c < channels; c += nwarps >> 2) {
				const int p = p0 + q * 32 + lane;
				if (p < plane) img[(long long)c * plane + p] = 1.f;
			}
		} else if (mode == 1) {
			for (int c = warp; c < channels; c += nwarps) {
				const int p = p0 + lane * 4;
				if (p + 3 < plane) *(float4*)(img + (long long)c * plane + p) = make_float4(1.f, 2.f, 3.f, 4.f);
			}
		} else {
			for (int c = warp; c < channels; c += nwarps) {
				#pragma unroll
				for (int j = 0; j < 4; j++) {
					const int p = p0 + j * 32 + lane;
					if (p < plane) img[(long long)c * plane + p] = 1.f;
				}
			}
		}
	}
}

int main()
{
	const int images = 64, channels = 256;
	float* buf;
	cudaMalloc(&buf, (size_t)images * channels * 56 * 56 * 4 + 4096);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	const char* names[] = {"128 B requests, 4 warps/row", "512 B requests, 16 B lanes", "4 x 128 B, same warp", "contiguous float4", "4 warps/row, lockstep"};
	for (int plane : {55 * 55, 56 * 56}) {
		for (int warps : {8, 16, 32}) {
			for (int mode = 0; mode < 5; mode++) {
				if (mode == 1 && plane % 4) continue;
				if (mode == 3 && plane != 55 * 55) continue;
				const long long tiles = (long long)images * ((plane + 127) / 128);
				store_kernel<<<148, warps * 32>>>(buf, mode, plane, channels, tiles, images);
				cudaEventRecord(e0);
				for (int r = 0; r < 5; r++) store_kernel<<<148, warps * 32>>>(buf, mode, plane, channels, tiles, images);
				cudaEventRecord(e1);
				cudaEventSynchronize(e1);
				float ms;
				cudaEventElapsedTime(&ms, e0, e1);
				const double bytes = (double)images * channels * plane * 4;
				printf("plane %4d  warps/SM %2d  %-30s : %7.1f GB/s\n", plane, warps, names[mode], bytes * 5 / ms / 1e6);
			}
		}
	}
	printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
	return 0;
}
