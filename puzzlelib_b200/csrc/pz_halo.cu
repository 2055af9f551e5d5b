// pz_halo.cu -- stride-1 R x S convolutions (fprop and dgrad of 3x3 / 5x5 / 7x7 filters) with the activation operand staged ONCE
// per channel block in shared memory as a halo tile, instead of being gathered once per filter tap.
//
// B200 design.  The images are addressed in "padded position" space: position m' = n*L + hp*Wp + wp of the zero-padded
// (Hp x Wp) image, L = Hp*Wp.  For the output position whose window starts at m', filter tap (r, s) reads m' + r*Wp + s -- the
// same shift for EVERY output position, image borders included (they read the zero padding).  A tile of 128 consecutive
// output positions therefore needs the 128 + (R-1)*Wp + (S-1) consecutive padded positions starting at its first one: the
// halo.  Producer warps gather the halo of one block of 32 (fp32) / 64 (16-bit) channels into shared memory in the K-major
// SWIZZLE_128B layout (one 128-byte row per position) -- one global load per element, no tap masks -- and the MMA thread
// issues the R*S taps as tcgen05.mma instructions whose A descriptors simply start (r*Wp + s) rows further down the same
// halo: the swizzle is a function of the shared-memory address, so a row-shifted window of a swizzled tile is a valid tile.
// The filter tap tiles (prepared, tf32-rounded, channel-ordered copy) stream through their own TMA ring.  Outputs at padded
// positions that are not real outputs (q >= Q or p >= P) are computed and dropped by the epilogue (7% of the rows of a
// 55x55 map, 13% of 28x28, 23% of 14x14).  Accumulators are double-buffered in TMEM, one persistent CTA per SM.
//
// Warp roles: 16 gather warps | 1 TMA-issuing warp | 1 MMA-issuing warp | 4 epilogue warps (704 threads; the register file is
// allocated in units of 4 warps, so 22 warps cost what 21 do).  The gather warps software-pipeline their loads: the halo of
// stage s+1 is in flight while the halo of stage s is rounded and stored.
#include "pz_umma.cuh"

namespace pzumma {

constexpr int NGATHER_WARPS = 16;
constexpr int NGATHER = NGATHER_WARPS * 32;
constexpr int NTHREADS_HALO = NGATHER + 64 + NEPI_WARPS * 32;
constexpr int PIPE_TASKS = 4;        // chunk tasks per thread whose loads are double-buffered in registers (halo <= 256 rows)
constexpr int HALO_STAGES = 2;
constexpr int MAX_BSTAGES = 8;
constexpr int MAX_TASKS = 8;         // halo chunk tasks per gather thread and stage (8 * 512 / 8 = 512 halo rows at most)

struct HaloParams {
	const void* x;               // source tensor (N, C_total, H, W): x for fprop, dy for dgrad
	Epilogue E;
	FastDiv ldiv, wpdiv;         // L = Hp * Wp, Wp
	int H, W, ph, pw, Wp;        // source map size and its offset inside the padded image
	int chans;                   // reduction channels per group
	long long img_stride, chan_stride, group_stride;   // elements
	int R, S, cblocks;           // filter taps, channel blocks per group
	int halo_rows, halo_bytes;   // rows of one halo stage, bytes per stage (multiple of 1024)
	int bstages;
	long long rows_total;        // N * L
	int out_h, out_w;            // valid output positions inside the padded image
	int tiles_m, tiles_n, groups;
	int tma_rows_per_group;
	int ab_bf16;
	double alg_flops, alg_bytes;
};

// A window that starts `rows` 128-byte rows into a swizzled tile: the hardware applies the 128-byte swizzle to the address it
// computes (start + row * 128 + k bytes), and the gather warps swizzle by the row index relative to the 1024-byte aligned halo
// base, i.e. by the same address bits -- so only the start address moves (verified on B200: setting the descriptor's
// base-offset field to the row phase instead gives wrong results).
__device__ __forceinline__ uint64_t make_smem_desc_rows(uint32_t halo, uint32_t rows) { return make_smem_desc(halo + rows * 128u); }

template <int BN, bool H16>
__global__ void __launch_bounds__(NTHREADS_HALO, 1) umma_halo_kernel(const __grid_constant__ HaloParams p, const __grid_constant__ CUtensorMap tmapB)
{
	using EL = typename std::conditional<H16, uint16_t, float>::type;
	constexpr int BKE = H16 ? BK16 : BK;
	constexpr int CH = H16 ? 8 : 4;                  // channels per 16-byte chunk
	constexpr int BSTAGE_BYTES = BN * 128;
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t halo0 = smem0;
	const uint32_t bst0 = halo0 + HALO_STAGES * p.halo_bytes;
	const uint32_t bars = bst0 + p.bstages * BSTAGE_BYTES;
	const uint32_t bar_hfull = bars, bar_hempty = bars + 16, bar_bfull = bars + 32, bar_bempty = bar_bfull + 8 * MAX_BSTAGES;
	const uint32_t bar_accfull = bar_bempty + 8 * MAX_BSTAGES, bar_accempty = bar_accfull + 16;
	const uint32_t tmem_slot = bar_accempty + 16;

	const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
	const int lane = threadIdx.x & 31;
	const int total_work = p.tiles_m * p.tiles_n * p.groups;
	const int RS = p.R * p.S;
	constexpr int MMA_WARP = NGATHER_WARPS + 1, TMA_WARP = NGATHER_WARPS;

	if (warp == MMA_WARP) {
		if (lane == 0) {
			for (int s = 0; s < HALO_STAGES; s++) {
				mbar_init(bar_hfull + 8 * s, NGATHER_WARPS);
				mbar_init(bar_hempty + 8 * s, 1);
			}
			for (int s = 0; s < p.bstages; s++) {
				mbar_init(bar_bfull + 8 * s, 1);
				mbar_init(bar_bempty + 8 * s, 1);
			}
			for (int a = 0; a < 2; a++) {
				mbar_init(bar_accfull + 8 * a, 1);
				mbar_init(bar_accempty + 8 * a, NEPI_WARPS);
			}
			fence_barrier_init();
		}
		__syncwarp();
		tmem_alloc(tmem_slot, 2 * BN);
	}

	tc_fence_before();
	__syncthreads();
	tc_fence_after();

	uint32_t tmem_base;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

	auto decode = [&](int t, int& m_tile, int& n_tile, int& group) {
		n_tile = t % p.tiles_n;         // n fastest: CTAs running together share the halo through L2
		t /= p.tiles_n;
		m_tile = t % p.tiles_m;
		group = t / p.tiles_m;
	};

	if (warp < NGATHER_WARPS) {
		// ===================== halo gather =====================
		const int t = threadIdx.x;
		const int rows_pad = (p.halo_rows + 31) & ~31;           // lanes of a warp stay on consecutive rows of one chunk
		const int ntasks = rows_pad * 8;
		int hs = 0;
		uint32_t hphase = 0;

		if (ntasks <= PIPE_TASKS * NGATHER) {
			// ---- pipelined gather (halos of at most 256 rows): two register buffers, loads of the next stage in flight
			uint32_t saddr[PIPE_TASKS];
			#pragma unroll
			for (int i = 0; i < PIPE_TASKS; i++) {
				const int task = t + i * NGATHER;
				saddr[i] = 0xffffffffu;
				if (task < ntasks) {
					const int chunk = task / rows_pad, row = task - chunk * rows_pad;
					if (row < p.halo_rows) saddr[i] = (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
					else saddr[i] = 0xfffffff0u | (uint32_t)chunk;      // nothing to store; keeps the chunk for the decode below
				}
			}
			// load cursor
			int lwork = blockIdx.x, lcb = 0;
			int rowoff[PIPE_TASKS];
			const EL* __restrict__ lbase = (const EL*)p.x;
			auto issue = [&](uint32_t (&v)[PIPE_TASKS][4]) {
				if (lcb == 0) {
					int m_tile, n_tile, group;
					decode(lwork, m_tile, n_tile, group);
					lbase = (const EL*)p.x + (long long)group * p.group_stride;
					#pragma unroll
					for (int i = 0; i < PIPE_TASKS; i++) {
						const int task = t + i * NGATHER;
						rowoff[i] = -1;
						if (task < ntasks) {
							const int row = task % rows_pad;
							const long long pm = (long long)m_tile * BM + row;
							if (row < p.halo_rows && pm < p.rows_total) {
								const uint32_t n = fdiv((uint32_t)pm, p.ldiv);
								const uint32_t rem = (uint32_t)pm - n * p.ldiv.d;
								const uint32_t hp = fdiv(rem, p.wpdiv);
								const int h = (int)hp - p.ph, w = (int)(rem - hp * p.wpdiv.d) - p.pw;
								if ((unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W)
									rowoff[i] = (int)((long long)n * p.img_stride + (long long)h * p.W + w);
							}
						}
					}
				}
				const int cbase = lcb * BKE;
				#pragma unroll
				for (int i = 0; i < PIPE_TASKS; i++) {
					const uint32_t sa = saddr[i];
					const int chunk = sa >= 0xfffffff0u ? (int)(sa & 7u) : (int)(((sa >> 4) ^ (sa >> 7)) & 7u);
					const int c0 = cbase + chunk * CH;
					const bool ok = rowoff[i] >= 0;
					const EL* __restrict__ src = lbase + (rowoff[i] + (long long)c0 * p.chan_stride);
					#pragma unroll
					for (int e = 0; e < 4; e++) {
						if (H16) {
							const uint32_t lo = ldg16_pred(src + (long long)(2 * e) * p.chan_stride, ok && c0 + 2 * e < p.chans);
							const uint32_t hi = ldg16_pred(src + (long long)(2 * e + 1) * p.chan_stride, ok && c0 + 2 * e + 1 < p.chans);
							v[i][e] = lo | (hi << 16);
						} else {
							// raw bits; rounded to tf32 when stored
							v[i][e] = __float_as_uint(ldg_pred(reinterpret_cast<const float*>(src) + (long long)e * p.chan_stride, ok && c0 + e < p.chans));
						}
					}
				}
				if (++lcb == p.cblocks) { lcb = 0; lwork += gridDim.x; }
			};
			auto commit = [&](const uint32_t (&v)[PIPE_TASKS][4]) {
				mbar_wait(bar_hempty + 8 * hs, hphase ^ 1);
				const uint32_t halo = halo0 + hs * p.halo_bytes;
				#pragma unroll
				for (int i = 0; i < PIPE_TASKS; i++)
					if (saddr[i] < 0xfffffff0u) {
						if (H16) sts128(halo + saddr[i], v[i][0], v[i][1], v[i][2], v[i][3]);
						else sts128(halo + saddr[i], v[i][0] + 0x1000u, v[i][1] + 0x1000u, v[i][2] + 0x1000u, v[i][3] + 0x1000u);   // to_tf32
					}
				fence_async_smem();
				__syncwarp();
				if (lane == 0) mbar_arrive(bar_hfull + 8 * hs);
				if (++hs == HALO_STAGES) { hs = 0; hphase ^= 1; }
			};
			uint32_t va[PIPE_TASKS][4], vb[PIPE_TASKS][4];
			int pending = 0;                              // stages loaded but not yet stored
			if (lwork < total_work) { issue(va); pending++; }
			while (pending > 0) {
				if (lwork < total_work) { issue(vb); pending++; }
				commit(va);
				pending--;
				if (pending == 0) break;
				if (lwork < total_work) { issue(va); pending++; }
				commit(vb);
				pending--;
			}
		} else
		for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
			int m_tile, n_tile, group;
			decode(work, m_tile, n_tile, group);
			const EL* __restrict__ base = (const EL*)p.x + (long long)group * p.group_stride;

			// per-tile decode of this thread's halo rows: element offset of (row, channel 0) or -1 for padding / out of range
			int rowoff[MAX_TASKS];             // tensors hold < 2^31 elements
			uint32_t saddr[MAX_TASKS];
			#pragma unroll
			for (int i = 0; i < MAX_TASKS; i++) {
				const int task = t + i * NGATHER;
				rowoff[i] = -1;
				saddr[i] = 0;
				if (task < ntasks) {
					const int chunk = task / rows_pad, row = task - chunk * rows_pad;
					saddr[i] = (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
					const long long pm = (long long)m_tile * BM + row;
					if (row < p.halo_rows && pm < p.rows_total) {
						const uint32_t n = fdiv((uint32_t)pm, p.ldiv);
						const uint32_t rem = (uint32_t)pm - n * p.ldiv.d;
						const uint32_t hp = fdiv(rem, p.wpdiv);
						const int h = (int)hp - p.ph, w = (int)(rem - hp * p.wpdiv.d) - p.pw;
						if ((unsigned)h < (unsigned)p.H && (unsigned)w < (unsigned)p.W)
							rowoff[i] = (int)((long long)n * p.img_stride + (long long)h * p.W + w);
					}
					if (row >= p.halo_rows) saddr[i] = 0xffffffffu;      // padding rows of the task grid: nothing to store
				} else
					saddr[i] = 0xffffffffu;
			}

			for (int cb = 0; cb < p.cblocks; cb++) {
				mbar_wait(bar_hempty + 8 * hs, hphase ^ 1);
				const uint32_t halo = halo0 + hs * p.halo_bytes;
				const int cbase = cb * BKE;
				#pragma unroll
				for (int i0 = 0; i0 < MAX_TASKS; i0 += 4) {
					if (t + i0 * NGATHER >= ntasks) break;
					uint32_t v[4][4];
					#pragma unroll
					for (int i = 0; i < 4; i++) {
						const uint32_t sa = saddr[i0 + i];
						const int c0 = cbase + (int)(((sa >> 4) ^ (sa >> 7)) & 7u) * CH;      // chunk index back from the swizzled address
						const bool ok = rowoff[i0 + i] >= 0;
						const EL* __restrict__ src = base + (rowoff[i0 + i] + (long long)c0 * p.chan_stride);
						#pragma unroll
						for (int e = 0; e < 4; e++) {
							if (H16) {
								const uint32_t lo = ldg16_pred(src + (long long)(2 * e) * p.chan_stride, ok && c0 + 2 * e < p.chans);
								const uint32_t hi = ldg16_pred(src + (long long)(2 * e + 1) * p.chan_stride, ok && c0 + 2 * e + 1 < p.chans);
								v[i][e] = lo | (hi << 16);
							} else {
								v[i][e] = to_tf32(ldg_pred(reinterpret_cast<const float*>(src) + (long long)e * p.chan_stride, ok && c0 + e < p.chans));
							}
						}
					}
					#pragma unroll
					for (int i = 0; i < 4; i++)
						if (saddr[i0 + i] != 0xffffffffu) sts128(halo + saddr[i0 + i], v[i][0], v[i][1], v[i][2], v[i][3]);
				}
				fence_async_smem();
				__syncwarp();
				if (lane == 0) mbar_arrive(bar_hfull + 8 * hs);
				if (++hs == HALO_STAGES) { hs = 0; hphase ^= 1; }
			}
		}
	} else if (warp == TMA_WARP) {
		// ===================== filter tap tiles by TMA =====================
		if (lane == 0) {
			int bs = 0;
			uint32_t bphase = 0;
			for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
				int m_tile, n_tile, group;
				decode(work, m_tile, n_tile, group);
				const int brow = group * p.tma_rows_per_group + n_tile * BN;
				for (int cb = 0; cb < p.cblocks; cb++)
					for (int tap = 0; tap < RS; tap++) {
						mbar_wait(bar_bempty + 8 * bs, bphase ^ 1);
						mbar_arrive_expect_tx(bar_bfull + 8 * bs, BSTAGE_BYTES);
						tma_load_2d(bst0 + bs * BSTAGE_BYTES, &tmapB, (tap * p.cblocks + cb) * BKE, brow, bar_bfull + 8 * bs);
						if (++bs == p.bstages) { bs = 0; bphase ^= 1; }
					}
			}
		}
		__syncwarp();
	} else if (warp == MMA_WARP) {
		// ===================== MMA issuer (one thread) =====================
		const uint32_t idesc = H16 ? make_idesc_f16(BM, BN, p.ab_bf16) : make_idesc_tf32(BM, BN);
		int hs = 0, bs = 0, as = 0;
		uint32_t hphase = 0, bphase = 0, aphase = 0;
		for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
			mbar_wait(bar_accempty + 8 * as, aphase ^ 1);
			tc_fence_after();
			const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
			for (int cb = 0; cb < p.cblocks; cb++) {
				mbar_wait(bar_hfull + 8 * hs, hphase);
				const uint32_t halo = halo0 + hs * p.halo_bytes;
				int tap = 0;
				for (int r = 0; r < p.R; r++)
					for (int s = 0; s < p.S; s++, tap++) {
						mbar_wait(bar_bfull + 8 * bs, bphase);
						tc_fence_after();
						if (lane == 0) {
							const uint64_t da = make_smem_desc_rows(halo, (uint32_t)(r * p.Wp + s));
							const uint64_t db = make_smem_desc(bst0 + bs * BSTAGE_BYTES);
							#pragma unroll
							for (int kk = 0; kk < 4; kk++) {
								const uint32_t acc = (cb > 0 || tap > 0 || kk > 0) ? 1u : 0u;
								if (H16) umma_f16(tmem_d, da + 2 * kk, db + 2 * kk, idesc, acc);
								else umma_tf32(tmem_d, da + 2 * kk, db + 2 * kk, idesc, acc);
							}
							umma_commit(bar_bempty + 8 * bs);
						}
						__syncwarp();
						if (++bs == p.bstages) { bs = 0; bphase ^= 1; }
					}
				if (lane == 0) umma_commit(bar_hempty + 8 * hs);
				__syncwarp();
				if (++hs == HALO_STAGES) { hs = 0; hphase ^= 1; }
			}
			if (lane == 0) umma_commit(bar_accfull + 8 * as);
			__syncwarp();
			if (++as == 2) { as = 0; aphase ^= 1; }
		}
		tc_fence_before();
	} else {
		// ===================== epilogue (4 warps; warp w may touch TMEM lanes 32*(w%4) .. +31) =====================
		const Epilogue& E = p.E;
		const int lg = warp & 3;
		int as = 0;
		uint32_t aphase = 0;
		for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
			int m_tile, n_tile, group;
			decode(work, m_tile, n_tile, group);
			const long long m = (long long)m_tile * BM + lg * 32 + lane;
			int m0 = 0, m1 = 0, m2 = 0;
			bool mvalid = m < p.rows_total;
			if (mvalid) split3((uint32_t)m, E.md12, E.md2, m0, m1, m2);
			mvalid = mvalid && m1 < p.out_h && m2 < p.out_w;       // padded positions that are not outputs are dropped
			const int oes = E.out_kind == OUT_F32 ? 4 : 2;
			char* outp = (char*)E.out + ((long long)group * E.group_stride + ((long long)m0 * E.ms0 + (long long)m1 * E.ms1 + (long long)m2 * E.ms2)) * oes;
			const char* biasp = E.bias ? (const char*)E.bias + (long long)group * E.bias_group_stride * oes : nullptr;

			mbar_wait(bar_accfull + 8 * as, aphase);
			tc_fence_after();
			const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN) + ((uint32_t)(lg * 32) << 16);
			const int ncols = min(BN, E.N - n_tile * BN);
			const int fast = E.out_kind != OUT_F32 ? 0 : (E.bias_mode == 1 ? 2 : 1);

			uint32_t v0[EPI_COLS], v1[EPI_COLS];
			tmem_ld16(tmem_d, v0);
			#pragma unroll 1
			for (int c0 = 0; c0 < ncols; c0 += 2 * EPI_COLS) {
				tmem_wait_ld(v0);
				if (c0 + EPI_COLS < ncols) tmem_ld16(tmem_d + (uint32_t)(c0 + EPI_COLS), v1);
				epilogue_chunk(E, v0, outp, biasp, 0.0f, mvalid, true, n_tile * BN + c0, fast, 0, 0);
				if (c0 + EPI_COLS < ncols) {
					tmem_wait_ld(v1);
					if (c0 + 2 * EPI_COLS < ncols) tmem_ld16(tmem_d + (uint32_t)(c0 + 2 * EPI_COLS), v0);
					epilogue_chunk(E, v1, outp, biasp, 0.0f, mvalid, true, n_tile * BN + c0 + EPI_COLS, fast, 0, 0);
				}
			}
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_accempty + 8 * as);
			if (++as == 2) { as = 0; aphase ^= 1; }
		}
	}

	tc_fence_before();
	__syncthreads();
	if (warp == MMA_WARP) {
		tc_fence_after();
		tmem_dealloc(tmem_base, 2 * BN);
	}
}

// ------------------------------------------------------------------------------------------ host side
template <int BN, bool H16>
static int launch_halo_inst(const HaloParams& p, const CUtensorMap& tmap, int grid, size_t smem, cudaStream_t stream)
{
	auto kern = umma_halo_kernel<BN, H16>;
	static bool configured = false;
	if (!configured) {
		PZ_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
		configured = true;
	}
	{
		const bool hbm = p.alg_bytes > 0.0 && p.alg_flops / p.alg_bytes < kPzRidgeFlopPerByte;
		PzProfScope prof(hbm ? PZ_PROF_GEMM_HBM : PZ_PROF_GEMM, stream, p.alg_flops, p.alg_bytes);
		kern<<<grid, NTHREADS_HALO, smem, stream>>>(p, tmap);
	}
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

// Plans and launches the halo kernel; returns PZ_ERR_UNSUPPORTED (without setting an error) when the geometry does not fit, so
// that the caller can take the gather path instead.
int launch_halo(const HaloGeometry& g, int dtype, const TmaSource& tsrc, const Epilogue& Ein, double alg_flops, double alg_bytes,
				cudaStream_t stream)
{
	static const int disabled = [] { const char* e = getenv("PZ_NO_HALO"); return e ? atoi(e) : 0; }();
	if (disabled) return PZ_ERR_UNSUPPORTED;
	const bool h16 = dtype != PZ_F32;
	const int bke = elems_per_kblock(dtype);
	const int Hp = g.Hs + 2 * g.ph, Wp = g.Ws + 2 * g.pw;
	if (g.ph < 0 || g.pw < 0 || Hp - g.R + 1 != g.out_h || Wp - g.S + 1 != g.out_w) return PZ_ERR_UNSUPPORTED;
	const long long L = (long long)Hp * Wp, rows_total = (long long)g.N * L;
	if (rows_total >= (1ll << 31) || g.R * g.S < 2) return PZ_ERR_UNSUPPORTED;

	HaloParams p{};
	p.halo_rows = BM + (g.R - 1) * Wp + (g.S - 1);
	p.halo_bytes = (p.halo_rows * 128 + 1023) & ~1023;
	if (((p.halo_rows + 31) & ~31) * 8 > MAX_TASKS * NGATHER) return PZ_ERR_UNSUPPORTED;
	// a 128-position tile of a wide map drags (R-1) whole padded rows along: beyond ~3x the tile the per-tap gather is cheaper
	// (measured on VGG's 224-wide layers); those need a 2-d tiled halo
	if (p.halo_rows > 3 * BM + 32) return PZ_ERR_UNSUPPORTED;
	const int budget = 196 * 1024 - HALO_STAGES * p.halo_bytes;
	int bn = g.out_chans <= 64 ? 64 : (g.out_chans <= 128 ? 128 : 256);
	while (bn > 64 && budget / (bn * 128) < 4) bn /= 2;
	p.bstages = budget / (bn * 128);
	if (p.bstages < 3) return PZ_ERR_UNSUPPORTED;
	if (p.bstages > MAX_BSTAGES) p.bstages = MAX_BSTAGES;
	// junk rows cost tensor work only; refuse absurd cases (tiny maps with huge padding)
	if ((double)g.out_h * g.out_w < 0.45 * (double)L) return PZ_ERR_UNSUPPORTED;

	p.x = g.src;
	p.E = Ein;
	p.E.md12 = make_fastdiv((uint32_t)L);
	p.E.md2 = make_fastdiv((uint32_t)Wp);
	p.ldiv = make_fastdiv((uint32_t)L);
	p.wpdiv = make_fastdiv((uint32_t)Wp);
	p.H = g.Hs; p.W = g.Ws; p.ph = g.ph; p.pw = g.pw; p.Wp = Wp;
	p.chans = g.chans;
	p.img_stride = g.img_stride; p.chan_stride = g.chan_stride; p.group_stride = g.group_stride;
	p.R = g.R; p.S = g.S;
	p.cblocks = (g.chans + bke - 1) / bke;
	p.rows_total = rows_total;
	p.out_h = g.out_h; p.out_w = g.out_w;
	p.tiles_m = (int)pz_cdiv(rows_total, BM);
	p.tiles_n = (int)pz_cdiv(g.out_chans, bn);
	p.groups = g.groups;
	p.tma_rows_per_group = g.out_chans;
	p.ab_bf16 = dtype == PZ_BF16 ? 1 : 0;
	p.alg_flops = alg_flops;
	p.alg_bytes = alg_bytes;
	const long long units = (long long)p.tiles_m * p.tiles_n * p.groups;
	if (units >= (1ll << 31)) return PZ_ERR_UNSUPPORTED;
	const int grid = (int)(units < pz_num_sms() ? units : pz_num_sms());

	alignas(64) CUtensorMap tmap;
	int st = make_filter_tmap(&tmap, dtype, tsrc, bn);
	if (st != PZ_OK) return st;
	const size_t smem = (size_t)HALO_STAGES * p.halo_bytes + (size_t)p.bstages * bn * 128 + 512 + 1024;

#define PZ_HALO(BNV, H) if (bn == BNV && h16 == H) return launch_halo_inst<BNV, H>(p, tmap, grid, smem, stream);
	PZ_HALO(64, false) PZ_HALO(128, false) PZ_HALO(256, false)
	PZ_HALO(64, true) PZ_HALO(128, true) PZ_HALO(256, true)
#undef PZ_HALO
	return PZ_ERR_UNSUPPORTED;
}

}  // namespace pzumma
