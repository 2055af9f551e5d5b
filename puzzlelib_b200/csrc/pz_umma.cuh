// pz_umma.cuh -- one tcgen05 / TMEM GEMM engine for every dense contraction of the backend
// (Linear fwd/dgrad/wgrad, conv fprop/dgrad/wgrad, deconv).
//
// D[M x N] (+)= A[M x K] * B[N x K]^T with fp32 storage, TF32 tensor-core math, fp32 accumulation in TMEM.
//
// B200 design
//   * PERSISTENT kernel, one CTA per SM, static round-robin schedule over 128 x BN output tiles (BN = 64 / 128 / 256);
//     accumulators live in TMEM, DOUBLE-BUFFERED (2 x BN fp32 columns x 128 lanes) so that the epilogue of tile i
//     overlaps the mainloop of tile i+1; tcgen05.mma.cta_group::1.kind::tf32, M=128, N=BN, K=8 per instruction,
//     issued by ONE thread;
//   * operands are staged in shared memory as K-major tiles of 32 floats (128-byte rows) in the canonical
//     SWIZZLE_128B layout (8-row x 128-byte atoms, 16-byte chunk index XOR row%8) - the same layout a TMA
//     SWIZZLE_128B box produces, so a TMA producer can replace a gather producer tile by tile;
//   * warp roles: 16 producer warps in 4 independent groups | 1 MMA-issuer warp | 8 epilogue warps (896 threads;
//     setmaxnreg hands the producers 88 registers, the epilogue 56, the MMA warpgroup 32);
//   * the producer warps gather the activation operand straight from the reference's NCHW / row-major tensors
//     (no im2col buffer, no NHWC shadow copy), round fp32 -> tf32 (tcgen05 itself would truncate, which biases
//     every product by ~ -2^-11) and write the swizzled tile.  A group owns whole k-blocks (every 4th one): all of
//     a k-block's loads are issued -- as volatile asm, so that they stay ahead of the stores -- before the first
//     value is consumed, and the four groups keep four k-blocks in flight per SM;
//   * the filter operand of fprop / dgrad is pre-rounded + zero-padded by a tiny prep kernel and then fetched by
//     TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B box of 32 floats x BN rows) - no thread instructions at all;
//   * a ring of mbarriers (full: 4 warp arrivals [+ TMA transaction bytes], empty: tcgen05.commit) pipelines
//     4 - 8 smem stages (192 KB); accum-full / accum-empty barriers hand TMEM buffers to the epilogue warps;
//   * epilogue warps: tcgen05.ld 32 lanes x 32 columns, alpha/beta/bias, coalesced stores (TMEM lanes are always
//     mapped to the contiguous dimension of the output) or red.add for split-K.
//
// An operand is described by `Operand`: element (row, k) lives at
//     ptr[ r0*rs0 + k0*ks0 + hh*Wd + ww ],  (r0,r1,r2) = split(row), (k0,k1,k2) = split(k),
//     hh = (r1*ah + k1*bh + ch) / cdh,  ww = (r2*aw + k2*bw + cw) / cdw     (exact division required)
// and is zero unless 0<=hh<H, 0<=ww<W, row<rows, k<kdim.  That one formula covers dense matrices in either
// orientation, im2col views of NCHW tensors for fprop / wgrad, and the transposed-conv gather of dgrad with
// any stride / padding / dilation.
#pragma once

#include "pz_common.h"

#include <cuda.h>
#include <type_traits>

namespace pzumma {

// -DPZ_TIMELINE: per-role phase clocks of umma_gemm_kernel, summed over the launch (lane 0 of every warp adds its counters at the
// end; read and reset with pz_debug_timeline).  Instrumented builds are for diagnosis only (tools/gpu_timeline.sh).
#ifdef PZ_TIMELINE
static __device__ unsigned long long pz_timeline[32];
#define PZ_TL_DECL(n) long long tl_[n] = {}; long long tl_t = clock64();
#define PZ_TL(i) { const long long now_ = clock64(); tl_[i] += now_ - tl_t; tl_t = now_; }
#define PZ_TL_FLUSH(base, n) if (lane == 0) { for (int i_ = 0; i_ < (n); i_++) atomicAdd(&pz_timeline[(base) + i_], (unsigned long long)tl_[i_]); }
#define PZ_TL_TOUCH(x) asm volatile("" ::"f"(x));
#else
#define PZ_TL_DECL(n)
#define PZ_TL(i)
#define PZ_TL_FLUSH(base, n)
#define PZ_TL_TOUCH(x)
#endif

constexpr int BM = 128;          // tile rows  = TMEM lanes
constexpr int BK = 32;           // floats per k-block = one 128-byte swizzle row
// Producers work in NGROUPS independent groups of NPROD_WARPS warps: group g fills k-blocks g, g + NGROUPS, ... of the CTA's
// k-block sequence -- a whole k-block per group, every thread with its share of it (32 four-byte loads per operand) in
// flight at once.  What bounds a gather of 4-byte lanes is bytes in flight per SM (tools/ubench/gather_bw.cu: 512 threads
// with 8 loads each sustain 23.6 GB/s per SM, with 32 loads each 44 GB/s = the HBM rate), and a register ring of several
// k-blocks per thread does NOT add to it: all global loads of a thread land on one hardware scoreboard, a wait on a
// scoreboard waits for every load outstanding on it, so the ring degenerates to one k-block per DRAM round trip.
constexpr int NGROUPS = 4;
constexpr int NPROD_WARPS = 4;               // warps of one producer group
constexpr int NPROD = NPROD_WARPS * 32;      // threads of one producer group
constexpr int NPROD_WARPS_ALL = NGROUPS * NPROD_WARPS;
constexpr int NEPI_WARPS = 4;               // (halo kernel)
constexpr int NEPI_WARPS_GEMM = 8;          // two warps per TMEM lane quadrant, alternating 16-column chunks
// 28 warps = 7 warpgroups: 0-3 producer groups, 4 = MMA issuer (warp 16; 17-19 only give their registers away), 5-6 = epilogue.
// The kernel starts at 72 registers per thread; setmaxnreg moves registers from the MMA / epilogue warpgroups to the
// producers (a k-block of both operands in registers).  The pool is the launch allocation of 896*72 registers.
constexpr int MMA_WARP = NPROD_WARPS_ALL;
constexpr int EPI_WARP0 = NPROD_WARPS_ALL + 4;
constexpr int NTHREADS = (EPI_WARP0 + NEPI_WARPS_GEMM) * 32;
constexpr int REGS_PROD = 88, REGS_MMA = 32, REGS_EPI = 56;    // 512*88 + 128*32 + 256*56 <= 896*72 (setmaxnreg.inc blocks until the CTA has released enough)
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
constexpr int INVALID = -(1 << 28);

struct FastDiv {
	uint32_t d, m, sh;           // d == 0 encodes an "infinite" divisor (quotient 0); d == 1 is the identity
};

inline FastDiv make_fastdiv(uint32_t d)
{
	FastDiv f{d, 0, 0};
	if (d <= 1) return f;
	uint32_t s = 0;
	while ((1ull << s) < d) s++;
	uint32_t p = 31 + s;
	f.m = (uint32_t)(((1ull << p) + d - 1) / d);
	f.sh = p - 32;
	return f;
}

__device__ __forceinline__ uint32_t fdiv(uint32_t n, const FastDiv& f)
{
	return f.d == 1 ? n : (__umulhi(n, f.m) >> f.sh);
}

__device__ __forceinline__ void split3(uint32_t x, const FastDiv& d12, const FastDiv& d2, int& x0, int& x1, int& x2)
{
	uint32_t q0 = fdiv(x, d12);
	uint32_t rem = x - q0 * d12.d;
	uint32_t q1 = fdiv(rem, d2);
	x0 = (int)q0;
	x1 = (int)q1;
	x2 = (int)(rem - q1 * d2.d);
}

// producer kinds.  The *_GENERAL / *_SIMPLE ones evaluate the full index formula per element (any geometry);
// MN_TAP / K_TAP / K_DENSE are the fast paths: per-row (or per-k) base offsets and tap-validity bit masks are
// computed once, the per-element work is one predicated load + one tf32 rounding.
enum { MODE_K_GENERAL = 0, MODE_K_SIMPLE = 1, MODE_MN_GENERAL = 2, MODE_MN_SIMPLE = 3, MODE_MN_TAP = 4, MODE_K_TAP = 5,
	   MODE_K_DENSE = 6, MODE_TMA = 7, MODE_MN_CHAN = 8, MODE_K_POS_TAP = 9, MODE_K_POS_DENSE = 10, MODE_MN_VEC = 11,
	   MODE_K_POS_VEC = 12, MODE_K_POS_TMA = 13, MODE_MN_TMA = 14 };

struct Operand {
	const void* ptr;             // float or 16-bit (half / bfloat16) elements; all strides below are in ELEMENTS
	FastDiv rd12, rd2, kd12, kd2;
	int rs0, ks0;
	int ah, bh, ch, aw, bw, cw;
	int H, W, Wd;
	int cdh, cdw;
	int rows, kdim;
	int R, S;                    // taps of the (r, s) sub-index (1, 1 for dense operands); used by the *_TAP producers
	long long group_stride;
	// MODE_MN_CHAN: k-block kb = tap * cblocks + channel block (32 channels per block, `chans` channels in all)
	// MODE_K_POS*: k-block kb = image * kbpi + position block (32 positions per block, `plane` positions per image)
	FastDiv kbdiv;               // divisor cblocks / kbpi
	int chans, plane;
};

enum { OUT_F32 = 0, OUT_F16 = 1, OUT_BF16 = 2 };

struct Epilogue {
	void* out;
	const void* bias;
	int out_kind;                // element type of out / bias (OUT_*); red.add (split-K, col2im) needs OUT_F32
	FastDiv md12, md2;
	int ms0, ms1, ms2;           // row m -> m0*ms0 + m1*ms1 + m2*ms2
	int ncs;                     // column n -> n*ncs
	int M, N;
	float alpha, beta;
	int bias_mode;               // 0 none, 1 bias[n], 2 bias[m]
	int atomic;                  // 1: out += alpha*acc with red.global.add (split-K)
	// staged store (fp32 plain stores with bias_mode 0 / 1, see the epilogue of umma_gemm_kernel): the accumulator goes through a
	// [32 columns][128 rows] shared-memory tile so that a warp writes the 512 contiguous bytes of one output column back to back.
	// 0 off, 1 four 128-byte requests per column, 2 one 16-byte-lane request (row offsets contiguous and 16-byte aligned)
	int staged;
	long long group_stride;
	int bias_group_stride;
	// col2im epilogue (dgrad of a filter with very few input channels, see pz_conv.cu): row m = (image, p, q) of dy, column
	// n = (c, r, s); the product is scattered with red.add to dx[image][c][p*sh - ph + r*dh][q*sw - pw + s*dw]
	int c2i;
	int c2i_sh, c2i_sw, c2i_ph, c2i_pw, c2i_dh, c2i_dw, c2i_H, c2i_W;
	FastDiv c2i_rs, c2i_s;
};

struct GemmParams {
	Operand A, B;
	Epilogue E;
	int kblocks;                 // ceil(K / 32)
	int splits;                  // split-K factor
	int kb_per_split;
	int tiles_m, tiles_n, groups; // tile grid; total work units = tiles_m * tiles_n * groups * splits
	FastDiv fd_tiles_n, fd_tiles_m, fd_splits;   // the same counts as magic-number divisors (the decode runs per tile per warp)
	int tma_rows_per_group;      // MODE_TMA: row offset of group g in the prepared filter = g * tma_rows_per_group
	int ab_bf16;                 // 16-bit operands: 0 = half, 1 = bfloat16 (selects the tcgen05 input format)
	// MODE_MN_TMA: rows come in chunks of 32 positions of one image (one TMA box, one epilogue warp each); chunk = image *
	// img_chunks + block of 32 positions (the last chunk of an image may be partly past the plane); a row tile is 4 consecutive chunks
	int img_chunks, total_chunks;
	FastDiv fd_img_chunks;
	int debug_skip;              // PZ_DEBUG_SKIP (timing experiments only, results are wrong): 1 no epilogue stores, 2 no filter TMA, 4 no MMAs
								 // 8 / 16: MODE_MN_TMA descriptor with LBO and SBO swapped / SBO 1024 (the variants that were tried)
	double alg_flops, alg_bytes; // host-side bookkeeping for the profiler (algorithmic work of this launch)
};

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	uint32_t ok;
	do {
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(ok)
			: "r"(bar), "r"(parity)
			: "memory");
	} while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols)
{
	asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
	asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols)
{
	asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
		::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
		: "memory");
}
constexpr int EPI_COLS = 16;     // accumulator columns per tcgen05.ld (two loads in flight: 32 registers)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
	asm volatile(
		"tcgen05.ld.sync.aligned.32x32b.x16.b32 "
		"{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
		: "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
		  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
		: "r"(taddr)
		: "memory");
}
// tcgen05.wait::ld with the destination registers as in/out operands, so that no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&v)[16])
{
	asm volatile(
		"tcgen05.wait::ld.sync.aligned;"
		: "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
		  "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
		:
		: "memory");
}
// fp32 -> tf32, round to nearest (ties away): the tensor core reads only the upper 19 bits of each 32-bit operand word, so
// adding half a tf32 ulp is all that is needed (1 instruction; cvt.rna.tf32 is emulated with 3 on sm_100).  Finite values
// round correctly (including overflow to Inf); an Inf input becomes a NaN (its result would be Inf or NaN anyway).
__device__ __forceinline__ uint32_t to_tf32(float x)
{
	return __float_as_uint(x) + 0x1000u;
}
// Predicated global loads of the producers as volatile asm: they stay in program order ahead of the (volatile) shared
// stores, so a k-block's loads are ALL in flight before the first value is consumed.  (With __ldg the compiler is free to sink
// loads between the stores to save registers -- measured: ten load / wait / store rounds per k-block instead of one.)
__device__ __forceinline__ float ldg_pred(const void* p, bool ok)
{
	float v;
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\tmov.b32 %0, 0;\n\t@p ld.global.nc.b32 %0, [%1];\n\t}"
				 : "=f"(v) : "l"(p), "r"((int)ok));
	return v;
}
// 16-byte lane (the address must be 16-byte aligned when the predicate holds)
__device__ __forceinline__ void ldg128_pred(const void* p, bool ok, float& a, float& b, float& c, float& d)
{
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
				 "@p ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];\n\t}"
				 : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "l"(p), "r"((int)ok));
}
__device__ __forceinline__ uint32_t ldg16_pred(const void* p, bool ok)
{
	uint16_t v;
	asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %2, 0;\n\tmov.b16 %0, 0;\n\t@p ld.global.nc.b16 %0, [%1];\n\t}"
				 : "=h"(v) : "l"(p), "r"((int)ok));
	return (uint32_t)v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
	asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout=SWIZZLE_128B(2) [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr)
{
	uint64_t d = 0;
	d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
	d |= (uint64_t)1 << 16;                 // LBO: unused for swizzled K-major, canonical value 1
	d |= (uint64_t)(1024 >> 4) << 32;       // SBO: 8 rows x 128 B between 8-row groups
	d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
	d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
	return d;
}

// MN-major descriptor of the A operand of MODE_MN_TMA.  The only shared-memory layout tcgen05 accepts for MN-major tf32 operands
// is SWIZZLE_128B_BASE32B (layout type 1; cutlass/gemm/collective/builders/sm100_common.inl): an atom is 32 floats along M
// (one 128-byte row) x 4 along K (rows 128 bytes apart, 512 bytes in all), the 32-byte chunks of a row XORed with the row
// index mod 4 (Swizzle<2,5,2> on the byte address) -- what a tensor map with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes.
// LBO = bytes between atoms along M, SBO = bytes between 4-k atoms.  The tile is four TMA boxes of [32 k][32 m] floats:
// LBO = 4096 (next box), SBO = 512 (next 4 k of the same box); one MMA (8 k) reads two atoms along K.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t smem_addr, int variant)
{
	const uint32_t sbo = (variant & 16) ? 1024 : 512, lbo = 4096;
	uint64_t d = 0;
	d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
	d |= (uint64_t)(((variant & 8) ? sbo : lbo) >> 4) << 16;
	d |= (uint64_t)(((variant & 8) ? lbo : sbo) >> 4) << 32;
	d |= (uint64_t)1 << 46;
	d |= (uint64_t)1 << 61;                 // SWIZZLE_128B_BASE32B
	return d;
}

// tcgen05 instruction descriptor (cute::UMMA::InstrDescriptor), kind::tf32, fp32 accumulate, both K-major
// (bit 15 = A is MN-major, bit 16 = B is MN-major)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n)
{
	return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// kind::f16 (half or bfloat16 inputs, fp32 accumulate, both K-major): a_format / b_format 0 = f16, 1 = bf16
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n, int bf16)
{
	return (1u << 4) | ((uint32_t)bf16 << 7) | ((uint32_t)bf16 << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"setp.ne.b32 p, %4, 0;\n\t"
		"tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
		::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
		: "memory");
}

// ------------------------------------------------------------------------------------------ operand producers
struct RowInfo { int rbase, hr, wr, valid; };

__device__ __forceinline__ RowInfo make_rowinfo(const Operand& op, int row)
{
	RowInfo ri;
	int r0, r1, r2;
	bool valid = row < op.rows;
	split3((uint32_t)(valid ? row : 0), op.rd12, op.rd2, r0, r1, r2);
	ri.rbase = r0 * op.rs0;
	ri.hr = valid ? r1 * op.ah + op.ch : INVALID;
	ri.wr = r2 * op.aw + op.cw;
	ri.valid = valid;
	return ri;
}

struct KInfo { int kbase, hk, wk, valid; };

__device__ __forceinline__ KInfo make_kinfo(const Operand& op, int k)
{
	KInfo ki;
	int k0, k1, k2;
	bool valid = k < op.kdim;
	split3((uint32_t)(valid ? k : 0), op.kd12, op.kd2, k0, k1, k2);
	ki.kbase = k0 * op.ks0;
	ki.hk = valid ? k1 * op.bh : INVALID;
	ki.wk = k2 * op.bw;
	ki.valid = valid;
	return ki;
}

// fetch one element given row and k info; SIMPLE: no spatial bound checks (the host proved them unnecessary)
template <bool SIMPLE, bool CDIV>
__device__ __forceinline__ float fetch(const Operand& op, const float* __restrict__ base, const RowInfo& ri, const KInfo& ki)
{
	int hh = ri.hr + ki.hk;
	int ww = ri.wr + ki.wk;
	if (SIMPLE) {
		bool ok = ri.valid && ki.valid;
		int off = ri.rbase + ki.kbase + hh * op.Wd + ww;
		return ldg_pred(base + off, ok);
	}
	bool ok = true;
	if (CDIV) {
		ok = (hh % op.cdh == 0) && (ww % op.cdw == 0);
		hh /= op.cdh;
		ww /= op.cdw;
	}
	ok = ok && ((unsigned)hh < (unsigned)op.H) && ((unsigned)ww < (unsigned)op.W);
	int off = ri.rbase + ki.kbase + hh * op.Wd + ww;
	return ldg_pred(base + off, ok);
}

// MN-contiguous producer: thread owns one tile row (kept in registers) and ROWS/32 16-byte chunks per stage.
template <int ROWS, bool SIMPLE, bool CDIV>
struct MnProducer {
	static constexpr int NCH = 8 * ROWS / NPROD;     // 16-byte chunks per thread per stage (8 chunks per 128-byte row)
	static constexpr int CSTEP = NPROD / ROWS;       // chunk stride between a thread's chunks
	static constexpr int NV = NCH * 4;               // values per thread per stage
	RowInfo ri;
	int row_local, chunk0;

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp, int lane, uint32_t)
	{
		int t = warp * 32 + lane;
		row_local = t % ROWS;
		chunk0 = (warp * 32) / ROWS;                 // warp-uniform
		ri = make_rowinfo(op, tile_row0 + row_local);
	}
	__device__ __forceinline__ void load(const Operand& op, const float* __restrict__ base, int kb, float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NCH; i++) {
			int kc = kb * BK + (chunk0 + i * CSTEP) * 4;   // warp-uniform
			#pragma unroll
			for (int e = 0; e < 4; e++) {
				KInfo ki = make_kinfo(op, kc + e);
				v[i * 4 + e] = fetch<SIMPLE, CDIV>(op, base, ri, ki);
			}
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NCH; i++) {
			int chunk = chunk0 + i * CSTEP;
			uint32_t addr = tile + row_local * 128 + ((chunk ^ (row_local & 7)) << 4);
			sts128(addr, to_tf32(v[i * 4]), to_tf32(v[i * 4 + 1]), to_tf32(v[i * 4 + 2]), to_tf32(v[i * 4 + 3]));
		}
	}
};

// K-contiguous producer: a warp reads one tile row x 32 consecutive k per request (lane = k), rows w, w+8, ...
template <int ROWS, bool SIMPLE>
struct KProducer {
	static constexpr int NR = ROWS / NPROD_WARPS;    // rows per warp per stage
	static constexpr int NV = NR;
	int warp, lane;
	uint32_t rowinfo_smem;                           // smem table of RowInfo[ROWS]

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp_, int lane_, uint32_t table)
	{
		warp = warp_;
		lane = lane_;
		rowinfo_smem = table;
		int t = warp * 32 + lane;
		if (t < ROWS) {
			RowInfo ri = make_rowinfo(op, tile_row0 + t);
			sts128(table + t * 16, (uint32_t)ri.rbase, (uint32_t)ri.hr, (uint32_t)ri.wr, (uint32_t)ri.valid);
		}
	}
	__device__ __forceinline__ void load(const Operand& op, const float* __restrict__ base, int kb, float (&v)[NV])
	{
		KInfo ki = make_kinfo(op, kb * BK + lane);
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			RowInfo ri;
			asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
						 : "=r"(ri.rbase), "=r"(ri.hr), "=r"(ri.wr), "=r"(ri.valid)
						 : "r"(rowinfo_smem + (warp + i * NPROD_WARPS) * 16));
			v[i] = fetch<SIMPLE, false>(op, base, ri, ki);
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			const uint32_t row = (uint32_t)(warp + i * NPROD_WARPS);
			sts32(tile + row * 128 + (((((uint32_t)lane >> 2) ^ (row & 7)) << 4) | (((uint32_t)lane & 3) << 2)), to_tf32(v[i]));
		}
	}
};


// ------------------------------------------------------------------------------------------ fast producers
// An element is addressed as  base[ rowpart + kpart ]  with
//   "spatial" index (the side that holds (n, y, x)):  off = i0*s0 + (i1*a_h + c_h)*Wd + (i2*a_w + c_w)
//   "tap" index     (the side that holds (c, r, s)):  off = c*cs + r*b_h*Wd + s*b_w,   tap t = r*S + s
// and is valid iff tap t is in bounds for that spatial position -- one bit of a per-position mask.
// Tap entries are packed as  (offset - tapmin) << SH | (TOP - t) : the validity test is then a left shift of the mask by
// the low bits followed by a sign test (2 instructions).  t == TOP marks an out-of-range index: that mask bit is never
// set because the host picks the 64-bit mask whenever a filter has more than 31 taps (63 at most).
template <bool WIDE> struct TapMask;
template <> struct TapMask<false> {
	static constexpr uint32_t TOP = 31u, SH = 5u;
	uint32_t m;
	__device__ __forceinline__ void clear() { m = 0; }
	__device__ __forceinline__ void set(int t) { m |= 1u << t; }
	__device__ __forceinline__ bool test(uint32_t ent) const { return (int)(m << (ent & 31u)) < 0; }
};
template <> struct TapMask<true> {
	static constexpr uint32_t TOP = 63u, SH = 6u;
	unsigned long long m;
	__device__ __forceinline__ void clear() { m = 0; }
	__device__ __forceinline__ void set(int t) { m |= 1ull << t; }
	__device__ __forceinline__ bool test(uint32_t ent) const { return (long long)(m << (ent & 63u)) < 0; }
};

__device__ __forceinline__ int tap_min(int nR, int nS, int bhW, int bw)
{
	return min(0, (nR - 1) * bhW) + min(0, (nS - 1) * bw);
}

__device__ __forceinline__ float ldg_off(const char* __restrict__ sb, uint32_t elem)
{
	return __ldg(reinterpret_cast<const float*>(sb + ((unsigned long long)elem << 2)));
}

// rows = spatial positions (MN-contiguous in memory), k = (c, r, s).  fprop, stride-1 dgrad, MN-contiguous dense.
// Thread owns one tile row and ROWS/32 16-byte chunks (4 consecutive k) per stage; loads coalesce across lanes.
// Each stage, lane l decodes k = kb*32 + l once into a 128-byte per-warp smem table that all lanes read back as uint4.
template <int ROWS, bool WIDE>
struct MnTapProducer {
	using TM = TapMask<WIDE>;
	static constexpr int NCH = 8 * ROWS / NPROD;
	static constexpr int CSTEP = NPROD / ROWS;
	static constexpr int NV = NCH * 4;
	TM mask;
	int poff, row_local, chunk0, lane;
	uint32_t tab;

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp, int lane_, uint32_t table)
	{
		const int t = warp * 32 + lane_;
		lane = lane_;
		tab = table + warp * 128;
		row_local = t % ROWS;
		chunk0 = (warp * 32) / ROWS;
		const int row = tile_row0 + row_local;
		mask.clear();
		int r0, r1, r2;
		split3((uint32_t)(row < op.rows ? row : 0), op.rd12, op.rd2, r0, r1, r2);
		const int hr = r1 * op.ah + op.ch, wr = r2 * op.aw + op.cw;
		poff = r0 * op.rs0 + hr * op.Wd + wr + tap_min(op.R, op.S, op.bh * op.Wd, op.bw);
		if (row < op.rows) {
			for (int r = 0; r < op.R; r++) {
				const bool okh = (unsigned)(hr + r * op.bh) < (unsigned)op.H;
				for (int s = 0; s < op.S; s++)
					if (okh && (unsigned)(wr + s * op.bw) < (unsigned)op.W) mask.set(r * op.S + s);
			}
		}
	}
	__device__ __forceinline__ void load(const Operand& op, const float* __restrict__ base, int kb, float (&v)[NV])
	{
		// offsets are kept relative to the first channel of the stage so that they fit the entry's offset field
		const int k = kb * BK + lane;
		const uint32_t cfirst = fdiv((uint32_t)(kb * BK), op.kd12);
		const char* __restrict__ sb = reinterpret_cast<const char*>(base + (long long)cfirst * op.ks0 + poff);
		uint32_t entry = 0u;                         // tap TOP, never valid
		if (k < op.kdim) {
			const uint32_t c = fdiv((uint32_t)k, op.kd12);
			const uint32_t t = (uint32_t)k - c * op.kd12.d;
			const uint32_t r = fdiv(t, op.kd2);
			const uint32_t s = t - r * op.kd2.d;
			const int koff = (int)(c - cfirst) * op.ks0 + (int)r * op.bh * op.Wd + (int)s * op.bw - tap_min(op.R, op.S, op.bh * op.Wd, op.bw);
			entry = ((uint32_t)koff << TM::SH) | (TM::TOP - t);
		}
		__syncwarp();
		sts32(tab + lane * 4, entry);
		__syncwarp();
		#pragma unroll
		for (int i = 0; i < NCH; i++) {
			uint32_t e4[4];
			asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
						 : "=r"(e4[0]), "=r"(e4[1]), "=r"(e4[2]), "=r"(e4[3]) : "r"(tab + (chunk0 + i * CSTEP) * 16));
			#pragma unroll
			for (int e = 0; e < 4; e++) v[i * 4 + e] = ldg_pred(sb + ((unsigned long long)(e4[e] >> TM::SH) << 2), mask.test(e4[e]));
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NCH; i++) {
			const int chunk = chunk0 + i * CSTEP;
			const uint32_t addr = tile + row_local * 128 + ((chunk ^ (row_local & 7)) << 4);
			sts128(addr, to_tf32(v[i * 4]), to_tf32(v[i * 4 + 1]), to_tf32(v[i * 4 + 2]), to_tf32(v[i * 4 + 3]));
		}
	}
};

// rows = (c, r, s) taps, k = spatial positions (K-contiguous in memory).  wgrad's activation operand.
// lane = k; warp w covers tile rows w, w+8, ...; the per-row tap entries live in a small smem table.
template <int ROWS, bool WIDE>
struct KTapProducer {
	using TM = TapMask<WIDE>;
	static constexpr int NR = ROWS / NPROD_WARPS;
	static constexpr int NV = NR;
	int warp, lane;
	uint32_t table;

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp_, int lane_, uint32_t table_)
	{
		warp = warp_;
		lane = lane_;
		table = table_;
		const int t = warp * 32 + lane;
		if (t < ROWS) {
			const int row = tile_row0 + t;
			uint32_t entry = 0u;
			if (row < op.rows) {
				const uint32_t c = fdiv((uint32_t)row, op.rd12);
				const uint32_t tp = (uint32_t)row - c * op.rd12.d;
				const uint32_t r = fdiv(tp, op.rd2);
				const uint32_t s = tp - r * op.rd2.d;
				const int roff = (int)c * op.rs0 + (int)r * op.ah * op.Wd + (int)s * op.aw - tap_min(op.R, op.S, op.ah * op.Wd, op.aw);
				entry = ((uint32_t)roff << TM::SH) | (TM::TOP - tp);
			}
			sts32(table + t * 4, entry);
		}
	}
	__device__ __forceinline__ void load(const Operand& op, const float* __restrict__ base, int kb, float (&v)[NV])
	{
		const int k = kb * BK + lane;
		TM mask;
		mask.clear();
		int k0, k1, k2;
		split3((uint32_t)(k < op.kdim ? k : 0), op.kd12, op.kd2, k0, k1, k2);
		const int hk = k1 * op.bh + op.ch, wk = k2 * op.bw + op.cw;
		const char* __restrict__ sb = reinterpret_cast<const char*>(
			base + ((long long)k0 * op.ks0 + hk * op.Wd + wk + tap_min(op.R, op.S, op.ah * op.Wd, op.aw)));
		if (k < op.kdim) {
			for (int r = 0; r < op.R; r++) {
				const bool okh = (unsigned)(hk + r * op.ah) < (unsigned)op.H;
				for (int s = 0; s < op.S; s++)
					if (okh && (unsigned)(wk + s * op.aw) < (unsigned)op.W) mask.set(r * op.S + s);
			}
		}
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			uint32_t ent;
			asm volatile("ld.shared.b32 %0, [%1];" : "=r"(ent) : "r"(table + (warp + i * NPROD_WARPS) * 4));
			v[i] = ldg_pred(sb + ((unsigned long long)(ent >> TM::SH) << 2), mask.test(ent));
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			const uint32_t row = (uint32_t)(warp + i * NPROD_WARPS);
			sts32(tile + row * 128 + (((((uint32_t)lane >> 2) ^ (row & 7)) << 4) | (((uint32_t)lane & 3) << 2)), to_tf32(v[i]));
		}
	}
};

// dense K-contiguous rows: element (row, k) at base[row*rs0 + (k / D)*ks0 + k % D]  (D = kd12.d, 0 = no batch split).
// weights [K][C*R*S], transposed-B GEMM operands, and wgrad's dy operand ([ko][(n, pq)], D = PQ).
// lane = k; warp w covers tile rows w, w+8, ... by stepping one 64-bit pointer.
template <int ROWS>
struct KDenseProducer {
	static constexpr int NR = ROWS / NPROD_WARPS;
	static constexpr int NV = NR;
	int warp, lane, nvalid;
	long long rowoff0, step;

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp_, int lane_, uint32_t)
	{
		warp = warp_;
		lane = lane_;
		const int row0 = tile_row0 + warp_;
		rowoff0 = (long long)row0 * op.rs0;
		step = (long long)NPROD_WARPS * op.rs0 * 4;
		const int left = op.rows - row0;
		nvalid = left <= 0 ? 0 : min(NR, (left + NPROD_WARPS - 1) / NPROD_WARPS);
	}
	__device__ __forceinline__ void load(const Operand& op, const float* __restrict__ base, int kb, float (&v)[NV])
	{
		const int k = kb * BK + lane;
		int koff = k;
		if (op.kd12.d > 1) {
			const uint32_t k0 = fdiv((uint32_t)k, op.kd12);
			koff = (int)k0 * op.ks0 + (int)((uint32_t)k - k0 * op.kd12.d);
		}
		const int n = k < op.kdim ? nvalid : 0;
		const char* __restrict__ p = reinterpret_cast<const char*>(base + rowoff0 + koff);
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			v[i] = ldg_pred(p, i < n);
			p += step;
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			const uint32_t row = (uint32_t)(warp + i * NPROD_WARPS);
			sts32(tile + row * 128 + (((((uint32_t)lane >> 2) ^ (row & 7)) << 4) | (((uint32_t)lane & 3) << 2)), to_tf32(v[i]));
		}
	}
};

// rows = spatial positions (MN-contiguous in memory), k ordered (tap, channel): k-block kb = t * cblocks + cb covers channels
// cb*32 .. cb*32+31 of ONE filter tap t = (r, s).  Tap validity is then a single per-thread predicate per stage and the 32
// channel addresses are the tap address + j * channel-stride: 3 instructions per element (address, load, tf32 rounding)
// instead of a table lookup + mask test per element.  fprop / dgrad whenever the channel count is not tiny; the prepared
// filter (TMA operand) is laid out in the same k order by the prep kernels.
template <int ROWS, bool WIDE>
struct MnChanProducer {
	using TM = TapMask<WIDE>;
	static constexpr int NCH = 8 * ROWS / NPROD;
	static constexpr int CSTEP = NPROD / ROWS;
	static constexpr int NV = NCH * 4;
	TM mask;
	int poff, row_local, chunk0;

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp, int lane, uint32_t)
	{
		const int t = warp * 32 + lane;
		row_local = t % ROWS;
		chunk0 = (warp * 32) / ROWS;
		const int row = tile_row0 + row_local;
		mask.clear();
		int r0, r1, r2;
		split3((uint32_t)(row < op.rows ? row : 0), op.rd12, op.rd2, r0, r1, r2);
		const int hr = r1 * op.ah + op.ch, wr = r2 * op.aw + op.cw;
		poff = r0 * op.rs0 + hr * op.Wd + wr;
		if (row < op.rows) {
			for (int r = 0; r < op.R; r++) {
				const bool okh = (unsigned)(hr + r * op.bh) < (unsigned)op.H;
				for (int s = 0; s < op.S; s++)
					if (okh && (unsigned)(wr + s * op.bw) < (unsigned)op.W) mask.set((int)TM::TOP - (r * op.S + s));
			}
		}
	}
	__device__ __forceinline__ void load(const Operand& op, const float* __restrict__ base, int kb, float (&v)[NV])
	{
		const uint32_t t = fdiv((uint32_t)kb, op.kbdiv);             // warp-uniform
		const uint32_t cb = (uint32_t)kb - t * op.kbdiv.d;
		const uint32_t r = fdiv(t, op.kd2);
		const uint32_t sx = t - r * op.kd2.d;
		const int tapoff = (int)r * op.bh * op.Wd + (int)sx * op.bw;
		const bool ok = mask.test(t);                                 // bit (TOP - t) of the mask, see init
		const int c0 = (int)cb * 32 + chunk0 * 4;
		const int cleft = op.chans - c0;
		// one 64-bit base per stage; channel j is a 32x32 -> 64-bit multiply-add on it (a single IMAD.WIDE each)
		const char* __restrict__ ptr = reinterpret_cast<const char*>(base + ((long long)c0 * op.ks0 + (poff + tapoff)));
		const unsigned ksb = (unsigned)op.ks0 * 4u;
		constexpr int JMAX = (NCH - 1) * CSTEP * 4 + 3;
		if (cleft > JMAX) {
			// every channel of the block exists (channel counts that are multiples of 32): one predicate for all loads
			#pragma unroll
			for (int i = 0; i < NCH; i++) {
				#pragma unroll
				for (int e = 0; e < 4; e++) {
					const unsigned j = (unsigned)(i * CSTEP * 4 + e);
					v[i * 4 + e] = ldg_pred(ptr + (unsigned long long)ksb * j, ok);
				}
			}
		} else {
			#pragma unroll
			for (int i = 0; i < NCH; i++) {
				#pragma unroll
				for (int e = 0; e < 4; e++) {
					const int j = i * CSTEP * 4 + e;
					v[i * 4 + e] = ldg_pred(ptr + (unsigned long long)ksb * (unsigned)j, ok && j < cleft);
				}
			}
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NCH; i++) {
			const int chunk = chunk0 + i * CSTEP;
			const uint32_t addr = tile + row_local * 128 + ((chunk ^ (row_local & 7)) << 4);
			sts128(addr, to_tf32(v[i * 4]), to_tf32(v[i * 4 + 1]), to_tf32(v[i * 4 + 2]), to_tf32(v[i * 4 + 3]));
		}
	}
};

// K-contiguous operands of wgrad with the reduction index ordered (image, position block): k-block kb = n * kbpi + pb covers
// positions pb*32 .. pb*32+31 of ONE image n (positions past the plane are zero), so the image / position decode is
// warp-uniform and a lane only derives (p, q) from its position.  lane = position; warp w covers tile rows w, w+16, ...
//
// KPosDense: rows are channels with a constant stride (dy, or x of a 1x1 convolution with unit stride and no padding).
template <int ROWS>
struct KPosDenseProducer {
	static constexpr int NR = ROWS / NPROD_WARPS;
	static constexpr int NV = NR;
	int warp, lane, nvalid;
	long long rowoff0, step;

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp_, int lane_, uint32_t)
	{
		warp = warp_;
		lane = lane_;
		const int row0 = tile_row0 + warp_;
		rowoff0 = (long long)row0 * op.rs0;
		step = (long long)NPROD_WARPS * op.rs0 * 4;
		const int left = op.rows - row0;
		nvalid = left <= 0 ? 0 : min(NR, (left + NPROD_WARPS - 1) / NPROD_WARPS);
	}
	__device__ __forceinline__ void load(const Operand& op, const float* __restrict__ base, int kb, float (&v)[NV])
	{
		const uint32_t n = fdiv((uint32_t)kb, op.kbdiv);              // warp-uniform
		const int pos = (int)((uint32_t)kb - n * op.kbdiv.d) * BK + lane;
		const int cnt = pos < op.plane ? nvalid : 0;
		const char* __restrict__ p = reinterpret_cast<const char*>(base + (rowoff0 + (long long)n * op.ks0 + pos));
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			v[i] = ldg_pred(p, i < cnt);
			p += step;
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			const uint32_t row = (uint32_t)(warp + i * NPROD_WARPS);
			sts32(tile + row * 128 + (((((uint32_t)lane >> 2) ^ (row & 7)) << 4) | (((uint32_t)lane & 3) << 2)), to_tf32(v[i]));
		}
	}
};

// KPosTap: rows = (c, r, s) filter taps of x (any stride / padding / dilation); per-row offsets + tap ids in a smem table.
template <int ROWS, bool WIDE>
struct KPosTapProducer {
	using TM = TapMask<WIDE>;
	static constexpr int NR = ROWS / NPROD_WARPS;
	static constexpr int NV = NR;
	int warp, lane;
	uint32_t table;

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp_, int lane_, uint32_t table_)
	{
		warp = warp_;
		lane = lane_;
		table = table_;
		const int t = warp * 32 + lane;
		if (t < ROWS) {
			const int row = tile_row0 + t;
			uint32_t entry = 0u;
			if (row < op.rows) {
				const uint32_t c = fdiv((uint32_t)row, op.rd12);
				const uint32_t tp = (uint32_t)row - c * op.rd12.d;
				const uint32_t r = fdiv(tp, op.rd2);
				const uint32_t s = tp - r * op.rd2.d;
				const int roff = (int)c * op.rs0 + (int)r * op.ah * op.Wd + (int)s * op.aw - tap_min(op.R, op.S, op.ah * op.Wd, op.aw);
				entry = ((uint32_t)roff << TM::SH) | (TM::TOP - tp);
			}
			sts32(table + t * 4, entry);
		}
	}
	__device__ __forceinline__ void load(const Operand& op, const float* __restrict__ base, int kb, float (&v)[NV])
	{
		const uint32_t n = fdiv((uint32_t)kb, op.kbdiv);              // warp-uniform
		const int pos = (int)((uint32_t)kb - n * op.kbdiv.d) * BK + lane;
		const bool pvalid = pos < op.plane;
		const uint32_t pp = fdiv((uint32_t)(pvalid ? pos : 0), op.kd2);
		const int qq = (pvalid ? pos : 0) - (int)pp * (int)op.kd2.d;
		const int hk = (int)pp * op.bh + op.ch, wk = qq * op.bw + op.cw;
		TM mask;
		mask.clear();
		if (op.R == 3 && op.S == 3) {
			// the common 3x3 filter: straight-line code (no loop / branch overhead per k-block)
			const uint32_t w0 = (unsigned)wk < (unsigned)op.W, w1 = (unsigned)(wk + op.aw) < (unsigned)op.W,
						   w2 = (unsigned)(wk + 2 * op.aw) < (unsigned)op.W;
			const uint32_t wm = w0 | (w1 << 1) | (w2 << 2);
			const uint32_t h0 = (unsigned)hk < (unsigned)op.H, h1 = (unsigned)(hk + op.ah) < (unsigned)op.H,
						   h2 = (unsigned)(hk + 2 * op.ah) < (unsigned)op.H;
			const uint32_t m9 = (h0 ? wm : 0u) | (h1 ? wm << 3 : 0u) | (h2 ? wm << 6 : 0u);
			mask.m = pvalid ? m9 : 0u;
		} else if (pvalid) {
			// valid(r, s) = vh(r) & vw(s): build the row of column bits once, then place it for every valid filter row
			TM wm;
			wm.clear();
			for (int s = 0; s < op.S; s++)
				if ((unsigned)(wk + s * op.aw) < (unsigned)op.W) wm.set(s);
			for (int r = 0; r < op.R; r++)
				if ((unsigned)(hk + r * op.ah) < (unsigned)op.H) mask.m |= wm.m << (r * op.S);
		}
		const char* __restrict__ sb = reinterpret_cast<const char*>(
			base + ((long long)n * op.ks0 + hk * op.Wd + wk + tap_min(op.R, op.S, op.ah * op.Wd, op.aw)));
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			uint32_t ent;
			asm volatile("ld.shared.b32 %0, [%1];" : "=r"(ent) : "r"(table + (warp + i * NPROD_WARPS) * 4));
			v[i] = ldg_pred(sb + ((unsigned long long)(ent >> TM::SH) << 2), mask.test(ent));
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			const uint32_t row = (uint32_t)(warp + i * NPROD_WARPS);
			sts32(tile + row * 128 + (((((uint32_t)lane >> 2) ^ (row & 7)) << 4) | (((uint32_t)lane & 3) << 2)), to_tf32(v[i]));
		}
	}
};

// ---- 16-byte lanes (float tensors whose planes are a multiple of four elements and 16-byte aligned: 28 x 28, 14 x 14, ...)
//
// A stage holds the same bytes as with the 4-byte producers; a thread issues 8 LDG.128 instead of 32 LDG.32 for them, and the
// shared-memory image is written with STS.128 in both cases.
//
// MnVec: the MnChan operand of a 1 x 1 / stride-1 / un-padded filter (rows = positions, contiguous in memory; k = channels).  A
// thread owns FOUR consecutive rows (one 16-byte vector per channel) and 8 channels: the 4 x 4 blocks it loads are transposed by
// register naming alone -- vector e of the load is channel e, word j of the store is channel j.  lane = (row quad % 8, channel
// quad % 4): a load instruction reads 4 channels x 128 contiguous bytes, a store instruction hits every 16-byte bank group 4 times.
template <int ROWS>
struct MnVecProducer {
	static_assert(ROWS == 128, "one row quad per thread");
	static constexpr int NV = 32;
	int poff, row0, cq0;
	bool rvalid;

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp, int lane, uint32_t)
	{
		row0 = (warp * 8 + (lane & 7)) * 4;
		cq0 = lane >> 3;
		const int row = tile_row0 + row0;
		rvalid = row < op.rows;                                       // rows % 4 == 0: a quad is valid as a whole
		const uint32_t r0 = fdiv((uint32_t)(rvalid ? row : 0), op.rd12);
		poff = (int)r0 * op.rs0 + (int)((uint32_t)(rvalid ? row : 0) - r0 * op.rd12.d);
	}
	__device__ __forceinline__ void load(const Operand& op, const float* __restrict__ base, int kb, float (&v)[NV])
	{
		const int c0 = kb * 32 + cq0 * 4;
		const char* __restrict__ ptr = reinterpret_cast<const char*>(base + ((long long)c0 * op.ks0 + poff));
		const unsigned long long ksb = (unsigned long long)(unsigned)op.ks0 * 4ull;
		#pragma unroll
		for (int it = 0; it < 2; it++) {
			#pragma unroll
			for (int e = 0; e < 4; e++) {
				const int j = it * 16 + e;
				ldg128_pred(ptr + ksb * (unsigned)j, rvalid && c0 + j < op.chans, v[it * 16 + e * 4], v[it * 16 + e * 4 + 1], v[it * 16 + e * 4 + 2],
							v[it * 16 + e * 4 + 3]);
			}
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int it = 0; it < 2; it++) {
			const int chunk = it * 4 + cq0;
			#pragma unroll
			for (int j = 0; j < 4; j++) {
				const int row = row0 + j;
				sts128(tile + row * 128 + ((chunk ^ (row & 7)) << 4), to_tf32(v[it * 16 + j]), to_tf32(v[it * 16 + 4 + j]), to_tf32(v[it * 16 + 8 + j]),
					   to_tf32(v[it * 16 + 12 + j]));
			}
		}
	}
};

// KPosVec: the KPosDense operand (rows = channel planes, k = positions of one image per k-block).  lane = (row % 4, 16-byte chunk of
// the 128-byte k-block): a load instruction reads 4 rows x 128 contiguous bytes, the vector goes to shared memory as it is.
template <int ROWS>
struct KPosVecProducer {
	static constexpr int NR = ROWS / 16;
	static constexpr int NV = NR * 4;
	int row0, chunk, nvalid;
	long long rowoff0, step;

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp, int lane, uint32_t)
	{
		row0 = warp * 4 + (lane >> 3);
		chunk = lane & 7;
		const int row = tile_row0 + row0;
		rowoff0 = (long long)row * op.rs0;
		step = 16ll * op.rs0 * 4;
		const int left = op.rows - row;
		nvalid = left <= 0 ? 0 : min(NR, (left + 15) / 16);
	}
	__device__ __forceinline__ void load(const Operand& op, const float* __restrict__ base, int kb, float (&v)[NV])
	{
		const uint32_t n = fdiv((uint32_t)kb, op.kbdiv);              // warp-uniform
		const int pos = (int)((uint32_t)kb - n * op.kbdiv.d) * BK + chunk * 4;
		const int cnt = pos < op.plane ? nvalid : 0;                  // plane % 4 == 0: a vector is valid as a whole
		const char* __restrict__ p = reinterpret_cast<const char*>(base + (rowoff0 + (long long)n * op.ks0 + pos));
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			ldg128_pred(p, i < cnt, v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
			p += step;
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			const uint32_t row = (uint32_t)(row0 + i * 16);
			sts128(tile + row * 128 + ((((uint32_t)chunk) ^ (row & 7)) << 4), to_tf32(v[i * 4]), to_tf32(v[i * 4 + 1]), to_tf32(v[i * 4 + 2]),
				   to_tf32(v[i * 4 + 3]));
		}
	}
};

// ------------------------------------------------------------------------------------------ 16-bit producers
// half / bfloat16 operands: a k-block is 64 elements (one 128-byte swizzle row), a 16-byte chunk 8 elements.  Values travel
// through the producers' registers as packed pairs (bit patterns held in `float` registers, never touched by arithmetic) and are
// stored unchanged: the tensor core reads 16-bit inputs exactly.
constexpr int BK16 = 64;
__device__ __forceinline__ uint32_t ldg16(const uint16_t* p) { return (uint32_t)__ldg(p); }
__device__ __forceinline__ float pack16(uint32_t lo, uint32_t hi) { return __uint_as_float(lo | (hi << 16)); }

template <bool CDIV>
__device__ __forceinline__ uint32_t fetch16(const Operand& op, const uint16_t* __restrict__ base, const RowInfo& ri, const KInfo& ki)
{
	int hh = ri.hr + ki.hk;
	int ww = ri.wr + ki.wk;
	bool ok = true;
	if (CDIV) {
		ok = (hh % op.cdh == 0) && (ww % op.cdw == 0);
		hh /= op.cdh;
		ww /= op.cdw;
	}
	ok = ok && ((unsigned)hh < (unsigned)op.H) && ((unsigned)ww < (unsigned)op.W);
	const int off = ri.rbase + ki.kbase + hh * op.Wd + ww;
	return ldg16_pred(base + off, ok);
}

// general MN-contiguous producer (any geometry; the slow path, e.g. a first layer with 3 input channels)
template <int ROWS, bool CDIV>
struct MnProducer16 {
	static constexpr int NCH = 8 * ROWS / NPROD;
	static constexpr int CSTEP = NPROD / ROWS;
	static constexpr int NV = NCH * 4;
	RowInfo ri;
	int row_local, chunk0;

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp, int lane, uint32_t)
	{
		const int t = warp * 32 + lane;
		row_local = t % ROWS;
		chunk0 = (warp * 32) / ROWS;
		ri = make_rowinfo(op, tile_row0 + row_local);
	}
	__device__ __forceinline__ void load(const Operand& op, const uint16_t* __restrict__ base, int kb, float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NCH; i++) {
			const int kc = kb * BK16 + (chunk0 + i * CSTEP) * 8;
			#pragma unroll
			for (int w = 0; w < 4; w++) {
				const KInfo k0 = make_kinfo(op, kc + 2 * w), k1 = make_kinfo(op, kc + 2 * w + 1);
				v[i * 4 + w] = pack16(fetch16<CDIV>(op, base, ri, k0), fetch16<CDIV>(op, base, ri, k1));
			}
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NCH; i++) {
			const int chunk = chunk0 + i * CSTEP;
			const uint32_t addr = tile + row_local * 128 + ((chunk ^ (row_local & 7)) << 4);
			sts128(addr, __float_as_uint(v[i * 4]), __float_as_uint(v[i * 4 + 1]), __float_as_uint(v[i * 4 + 2]), __float_as_uint(v[i * 4 + 3]));
		}
	}
};

// MnChanProducer for 16-bit elements: k-block kb = tap * cblocks + cb covers channels cb*64 .. cb*64+63 of one tap
template <int ROWS, bool WIDE>
struct MnChanProducer16 {
	using TM = TapMask<WIDE>;
	static constexpr int NCH = 8 * ROWS / NPROD;
	static constexpr int CSTEP = NPROD / ROWS;
	static constexpr int NV = NCH * 4;
	TM mask;
	int poff, row_local, chunk0;

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp, int lane, uint32_t)
	{
		const int t = warp * 32 + lane;
		row_local = t % ROWS;
		chunk0 = (warp * 32) / ROWS;
		const int row = tile_row0 + row_local;
		mask.clear();
		int r0, r1, r2;
		split3((uint32_t)(row < op.rows ? row : 0), op.rd12, op.rd2, r0, r1, r2);
		const int hr = r1 * op.ah + op.ch, wr = r2 * op.aw + op.cw;
		poff = r0 * op.rs0 + hr * op.Wd + wr;
		if (row < op.rows) {
			for (int r = 0; r < op.R; r++) {
				const bool okh = (unsigned)(hr + r * op.bh) < (unsigned)op.H;
				for (int s = 0; s < op.S; s++)
					if (okh && (unsigned)(wr + s * op.bw) < (unsigned)op.W) mask.set((int)TM::TOP - (r * op.S + s));
			}
		}
	}
	__device__ __forceinline__ void load(const Operand& op, const uint16_t* __restrict__ base, int kb, float (&v)[NV])
	{
		const uint32_t t = fdiv((uint32_t)kb, op.kbdiv);
		const uint32_t cb = (uint32_t)kb - t * op.kbdiv.d;
		const uint32_t r = fdiv(t, op.kd2);
		const uint32_t sx = t - r * op.kd2.d;
		const int tapoff = (int)r * op.bh * op.Wd + (int)sx * op.bw;
		const bool ok = mask.test(t);
		const int c0 = (int)cb * BK16 + chunk0 * 8;
		const int cleft = op.chans - c0;
		const uint16_t* __restrict__ ptr = base + ((long long)c0 * op.ks0 + (poff + tapoff));
		#pragma unroll
		for (int i = 0; i < NCH; i++) {
			#pragma unroll
			for (int w = 0; w < 4; w++) {
				const int j = i * CSTEP * 8 + 2 * w;
				const uint32_t lo = ldg16_pred(ptr + (size_t)((unsigned)j * (unsigned)op.ks0), ok && j < cleft);
				const uint32_t hi = ldg16_pred(ptr + (size_t)((unsigned)(j + 1) * (unsigned)op.ks0), ok && j + 1 < cleft);
				v[i * 4 + w] = pack16(lo, hi);
			}
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NCH; i++) {
			const int chunk = chunk0 + i * CSTEP;
			const uint32_t addr = tile + row_local * 128 + ((chunk ^ (row_local & 7)) << 4);
			sts128(addr, __float_as_uint(v[i * 4]), __float_as_uint(v[i * 4 + 1]), __float_as_uint(v[i * 4 + 2]), __float_as_uint(v[i * 4 + 3]));
		}
	}
};

// K-contiguous 16-bit producers: lane l owns k = 2l, 2l+1 of the 64-element k-block -- the same 4 bytes of the swizzled row that
// lane l writes for fp32 operands, so the store code is unchanged.
template <int ROWS>
struct KDenseProducer16 {
	static constexpr int NR = ROWS / NPROD_WARPS;
	static constexpr int NV = NR;
	int warp, lane, nvalid;
	long long rowoff0, step;

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp_, int lane_, uint32_t)
	{
		warp = warp_;
		lane = lane_;
		const int row0 = tile_row0 + warp_;
		rowoff0 = (long long)row0 * op.rs0;
		step = (long long)NPROD_WARPS * op.rs0;
		const int left = op.rows - row0;
		nvalid = left <= 0 ? 0 : min(NR, (left + NPROD_WARPS - 1) / NPROD_WARPS);
	}
	__device__ __forceinline__ void load(const Operand& op, const uint16_t* __restrict__ base, int kb, float (&v)[NV])
	{
		const int k = kb * BK16 + 2 * lane;
		int koff0 = k, koff1 = k + 1;
		if (op.kd12.d > 1) {
			const uint32_t q0 = fdiv((uint32_t)k, op.kd12), q1 = fdiv((uint32_t)(k + 1), op.kd12);
			koff0 = (int)q0 * op.ks0 + (int)((uint32_t)k - q0 * op.kd12.d);
			koff1 = (int)q1 * op.ks0 + (int)((uint32_t)(k + 1) - q1 * op.kd12.d);
		}
		const int n0 = k < op.kdim ? nvalid : 0, n1 = k + 1 < op.kdim ? nvalid : 0;
		const uint16_t* __restrict__ p = base + rowoff0;
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			v[i] = pack16(ldg16_pred(p + koff0, i < n0), ldg16_pred(p + koff1, i < n1));
			p += step;
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			const uint32_t row = (uint32_t)(warp + i * NPROD_WARPS);
			sts32(tile + row * 128 + (((((uint32_t)lane >> 2) ^ (row & 7)) << 4) | (((uint32_t)lane & 3) << 2)), __float_as_uint(v[i]));
		}
	}
};

template <int ROWS>
struct KPosDenseProducer16 {
	static constexpr int NR = ROWS / NPROD_WARPS;
	static constexpr int NV = NR;
	int warp, lane, nvalid;
	long long rowoff0, step;

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp_, int lane_, uint32_t)
	{
		warp = warp_;
		lane = lane_;
		const int row0 = tile_row0 + warp_;
		rowoff0 = (long long)row0 * op.rs0;
		step = (long long)NPROD_WARPS * op.rs0;
		const int left = op.rows - row0;
		nvalid = left <= 0 ? 0 : min(NR, (left + NPROD_WARPS - 1) / NPROD_WARPS);
	}
	__device__ __forceinline__ void load(const Operand& op, const uint16_t* __restrict__ base, int kb, float (&v)[NV])
	{
		const uint32_t n = fdiv((uint32_t)kb, op.kbdiv);
		const int pos = (int)((uint32_t)kb - n * op.kbdiv.d) * BK16 + 2 * lane;
		const int n0 = pos < op.plane ? nvalid : 0, n1 = pos + 1 < op.plane ? nvalid : 0;
		const uint16_t* __restrict__ p = base + (rowoff0 + (long long)n * op.ks0 + pos);
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			v[i] = pack16(ldg16_pred(p, i < n0), ldg16_pred(p + 1, i < n1));
			p += step;
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			const uint32_t row = (uint32_t)(warp + i * NPROD_WARPS);
			sts32(tile + row * 128 + (((((uint32_t)lane >> 2) ^ (row & 7)) << 4) | (((uint32_t)lane & 3) << 2)), __float_as_uint(v[i]));
		}
	}
};

template <int ROWS, bool WIDE>
struct KPosTapProducer16 {
	using TM = TapMask<WIDE>;
	static constexpr int NR = ROWS / NPROD_WARPS;
	static constexpr int NV = NR;
	int warp, lane;
	uint32_t table;

	__device__ __forceinline__ void init(const Operand& op, int tile_row0, int warp_, int lane_, uint32_t table_)
	{
		warp = warp_;
		lane = lane_;
		table = table_;
		const int t = warp * 32 + lane;
		if (t < ROWS) {
			const int row = tile_row0 + t;
			uint32_t entry = 0u;
			if (row < op.rows) {
				const uint32_t c = fdiv((uint32_t)row, op.rd12);
				const uint32_t tp = (uint32_t)row - c * op.rd12.d;
				const uint32_t r = fdiv(tp, op.rd2);
				const uint32_t s = tp - r * op.rd2.d;
				const int roff = (int)c * op.rs0 + (int)r * op.ah * op.Wd + (int)s * op.aw - tap_min(op.R, op.S, op.ah * op.Wd, op.aw);
				entry = ((uint32_t)roff << TM::SH) | (TM::TOP - tp);
			}
			sts32(table + t * 4, entry);
		}
	}
	__device__ __forceinline__ void position(const Operand& op, uint32_t n, int pos, TM& mask, const uint16_t* __restrict__ base,
											 const uint16_t* __restrict__& sb)
	{
		const bool pvalid = pos < op.plane;
		const uint32_t pp = fdiv((uint32_t)(pvalid ? pos : 0), op.kd2);
		const int qq = (pvalid ? pos : 0) - (int)pp * (int)op.kd2.d;
		const int hk = (int)pp * op.bh + op.ch, wk = qq * op.bw + op.cw;
		mask.clear();
		if (op.R == 3 && op.S == 3) {
			const uint32_t w0 = (unsigned)wk < (unsigned)op.W, w1 = (unsigned)(wk + op.aw) < (unsigned)op.W,
						   w2 = (unsigned)(wk + 2 * op.aw) < (unsigned)op.W;
			const uint32_t wm = w0 | (w1 << 1) | (w2 << 2);
			const uint32_t h0 = (unsigned)hk < (unsigned)op.H, h1 = (unsigned)(hk + op.ah) < (unsigned)op.H,
						   h2 = (unsigned)(hk + 2 * op.ah) < (unsigned)op.H;
			const uint32_t m9 = (h0 ? wm : 0u) | (h1 ? wm << 3 : 0u) | (h2 ? wm << 6 : 0u);
			mask.m = pvalid ? m9 : 0u;
		} else if (pvalid) {
			TM wm;
			wm.clear();
			for (int s = 0; s < op.S; s++)
				if ((unsigned)(wk + s * op.aw) < (unsigned)op.W) wm.set(s);
			for (int r = 0; r < op.R; r++)
				if ((unsigned)(hk + r * op.ah) < (unsigned)op.H) mask.m |= wm.m << (r * op.S);
		}
		sb = base + ((long long)n * op.ks0 + hk * op.Wd + wk + tap_min(op.R, op.S, op.ah * op.Wd, op.aw));
	}
	__device__ __forceinline__ void load(const Operand& op, const uint16_t* __restrict__ base, int kb, float (&v)[NV])
	{
		const uint32_t n = fdiv((uint32_t)kb, op.kbdiv);
		const int pos = (int)((uint32_t)kb - n * op.kbdiv.d) * BK16 + 2 * lane;
		TM m0, m1;
		const uint16_t* __restrict__ sb0;
		const uint16_t* __restrict__ sb1;
		position(op, n, pos, m0, base, sb0);
		position(op, n, pos + 1, m1, base, sb1);
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			uint32_t ent;
			asm volatile("ld.shared.b32 %0, [%1];" : "=r"(ent) : "r"(table + (warp + i * NPROD_WARPS) * 4));
			const uint32_t off = ent >> TM::SH;
			v[i] = pack16(ldg16_pred(sb0 + off, m0.test(ent)), ldg16_pred(sb1 + off, m1.test(ent)));
		}
	}
	__device__ __forceinline__ void store(uint32_t tile, const float (&v)[NV])
	{
		#pragma unroll
		for (int i = 0; i < NR; i++) {
			const uint32_t row = (uint32_t)(warp + i * NPROD_WARPS);
			sts32(tile + row * 128 + (((((uint32_t)lane >> 2) ^ (row & 7)) << 4) | (((uint32_t)lane & 3) << 2)), __float_as_uint(v[i]));
		}
	}
};

// filter operand fetched by TMA from the prepared (tf32-rounded, zero-padded, K-major) copy: no per-thread state
template <int ROWS>
struct TmaProducer {
	static constexpr int NV = 1;
	__device__ __forceinline__ void init(const Operand&, int, int, int, uint32_t) {}
	template <typename P>
	__device__ __forceinline__ void load(const Operand&, const P* __restrict__, int, float (&)[NV]) {}
	__device__ __forceinline__ void store(uint32_t, const float (&)[NV]) {}
};

template <int ROWS, int MODE, bool CDIV, bool H16> struct ProducerSel;
template <int ROWS, bool CDIV> struct ProducerSel<ROWS, MODE_K_GENERAL, CDIV, false> { using type = KProducer<ROWS, false>; };
template <int ROWS, bool CDIV> struct ProducerSel<ROWS, MODE_K_SIMPLE, CDIV, false> { using type = KProducer<ROWS, true>; };
template <int ROWS, bool CDIV> struct ProducerSel<ROWS, MODE_MN_GENERAL, CDIV, false> { using type = MnProducer<ROWS, false, CDIV>; };
template <int ROWS, bool CDIV> struct ProducerSel<ROWS, MODE_MN_SIMPLE, CDIV, false> { using type = MnProducer<ROWS, true, false>; };
// for the fast producers the CDIV slot selects the 64-bit tap mask (more than 32 taps, e.g. a 7x7 filter)
template <int ROWS, bool WIDE> struct ProducerSel<ROWS, MODE_MN_TAP, WIDE, false> { using type = MnTapProducer<ROWS, WIDE>; };
template <int ROWS, bool WIDE> struct ProducerSel<ROWS, MODE_K_TAP, WIDE, false> { using type = KTapProducer<ROWS, WIDE>; };
template <int ROWS, bool WIDE> struct ProducerSel<ROWS, MODE_K_DENSE, WIDE, false> { using type = KDenseProducer<ROWS>; };
template <int ROWS, bool WIDE, bool H16> struct ProducerSel<ROWS, MODE_TMA, WIDE, H16> { using type = TmaProducer<ROWS>; };
// MODE_K_POS_TMA: the KPosDense operand over 16-byte aligned planes, fetched by the copy engine through a 3-d tensor map
// (positions, channels, images): one box of 32 positions x ROWS channels of one image per k-block, zero-filled past the plane
template <int ROWS, bool WIDE, bool H16> struct ProducerSel<ROWS, MODE_K_POS_TMA, WIDE, H16> { using type = TmaProducer<ROWS>; };
// MODE_MN_TMA: the activation operand of a 1x1 / stride-1 convolution over 16-byte aligned planes, fetched as it lies in memory
// (positions contiguous = M-major) by four boxes of 32 positions x 32 channels per k-block; the MMA reads it through an MN-major
// descriptor, so nobody transposes it.  A box never leaves its image, a tile may: rows are chunked per image (GemmParams::img_chunks)
template <int ROWS, bool WIDE> struct ProducerSel<ROWS, MODE_MN_TMA, WIDE, false> { using type = TmaProducer<ROWS>; };
template <int ROWS, bool WIDE> struct ProducerSel<ROWS, MODE_MN_CHAN, WIDE, false> { using type = MnChanProducer<ROWS, WIDE>; };
template <int ROWS, bool WIDE> struct ProducerSel<ROWS, MODE_K_POS_TAP, WIDE, false> { using type = KPosTapProducer<ROWS, WIDE>; };
template <int ROWS, bool WIDE> struct ProducerSel<ROWS, MODE_K_POS_DENSE, WIDE, false> { using type = KPosDenseProducer<ROWS>; };
template <int ROWS, bool WIDE> struct ProducerSel<ROWS, MODE_MN_VEC, WIDE, false> { using type = MnVecProducer<ROWS>; };
template <int ROWS, bool WIDE> struct ProducerSel<ROWS, MODE_K_POS_VEC, WIDE, false> { using type = KPosVecProducer<ROWS>; };
// half / bfloat16 operands
template <int ROWS, bool CDIV> struct ProducerSel<ROWS, MODE_MN_GENERAL, CDIV, true> { using type = MnProducer16<ROWS, CDIV>; };
template <int ROWS, bool WIDE> struct ProducerSel<ROWS, MODE_MN_CHAN, WIDE, true> { using type = MnChanProducer16<ROWS, WIDE>; };
template <int ROWS, bool WIDE> struct ProducerSel<ROWS, MODE_K_DENSE, WIDE, true> { using type = KDenseProducer16<ROWS>; };
template <int ROWS, bool WIDE> struct ProducerSel<ROWS, MODE_K_POS_TAP, WIDE, true> { using type = KPosTapProducer16<ROWS, WIDE>; };
template <int ROWS, bool WIDE> struct ProducerSel<ROWS, MODE_K_POS_DENSE, WIDE, true> { using type = KPosDenseProducer16<ROWS>; };

template <int BN> struct Cfg {
	static constexpr int STAGE_BYTES = (BM + BN) * 128;
	static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;          // 8 / 6 / 4 stages for BN = 64 / 128 / 256
	static constexpr int TABLE_BYTES = (BM + (BN > 128 ? BN : 128)) * 16;  // one set of producer tables per producer group
	static constexpr int BAR_BYTES = 256;
	static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NGROUPS * TABLE_BYTES + BAR_BYTES + 1024;  // + align slack
	static constexpr int TMEM_COLS = 2 * BN;                           // double-buffered accumulator (power of two >= 32)
};

// one unit of work of the persistent schedule
struct Work {
	int m_tile, n_tile, group, split, kb_begin, kb_end;
};

__device__ __forceinline__ Work decode_work(const GemmParams& p, int t)
{
	Work w;
	// n fastest: CTAs running together share the activation tile through L2
	uint32_t u = (uint32_t)t, q = fdiv(u, p.fd_tiles_n);
	w.n_tile = (int)(u - q * (uint32_t)p.tiles_n);
	u = q;
	q = fdiv(u, p.fd_tiles_m);
	w.m_tile = (int)(u - q * (uint32_t)p.tiles_m);
	u = q;
	q = fdiv(u, p.fd_splits);
	w.split = (int)(u - q * (uint32_t)p.splits);
	w.group = (int)q;
	w.kb_begin = w.split * p.kb_per_split;
	w.kb_end = min(p.kblocks, w.kb_begin + p.kb_per_split);
	return w;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads)
{
	asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, uint32_t bar)
{
	asm volatile(
		"cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
		::"r"(dst), "l"((unsigned long long)tmap), "r"(c0), "r"(c1), "r"(bar)
		: "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar)
{
	asm volatile(
		"cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
		::"r"(dst), "l"((unsigned long long)tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
		: "memory");
}

// output element access by kind (OUT_F32 / OUT_F16 / OUT_BF16); `idx` in elements
__device__ __forceinline__ void out_store(void* base, size_t idx, float r, int kind)
{
	if (kind == OUT_F32) ((float*)base)[idx] = r;
	else if (kind == OUT_F16) ((__half*)base)[idx] = __float2half_rn(r);
	else ((__nv_bfloat16*)base)[idx] = __float2bfloat16_rn(r);
}
__device__ __forceinline__ float out_load(const void* base, size_t idx, int kind)
{
	if (kind == OUT_F32) return ((const float*)base)[idx];
	if (kind == OUT_F16) return __half2float(((const __half*)base)[idx]);
	return __bfloat162float(((const __nv_bfloat16*)base)[idx]);
}

// one 32-lane x 16-column chunk of the accumulator -> global memory.  `fast`: 1 = store alpha*acc + bias[m], 2 = store
// alpha*acc + bias[n], 3 = red.add alpha*acc, 4 = col2im scatter, 0 = generic (beta, bias with split-K, 16-bit outputs, ...).
// Lanes are the contiguous output dimension, so every store instruction of a warp writes consecutive addresses.
// `outp` points at the row's first element (element offsets are scaled by the element size here).
__device__ __forceinline__ void epilogue_chunk(const Epilogue& E, const uint32_t (&v)[EPI_COLS], char* outp, const void* biasp, float bias_m,
											   bool mvalid, bool addbias, int n0, int fast, int hb, int wb)
{
	if (!mvalid) return;
	const unsigned ncs = (unsigned)E.ncs;
	const float alpha = E.alpha;
	if (fast == 4) {
		// col2im scatter: the column decode is warp-uniform, the bounds test and the address are per lane
		const int HW = E.c2i_H * E.c2i_W;
		float* o = (float*)outp;
		#pragma unroll
		for (int j = 0; j < EPI_COLS; j++) {
			const int col = n0 + j;
			if (col < E.N) {
				const uint32_t c = fdiv((uint32_t)col, E.c2i_rs);
				const uint32_t t = (uint32_t)col - c * E.c2i_rs.d;
				const uint32_t r = fdiv(t, E.c2i_s);
				const uint32_t sx = t - r * E.c2i_s.d;
				const int h = hb + (int)r * E.c2i_dh, w = wb + (int)sx * E.c2i_dw;
				if ((unsigned)h < (unsigned)E.c2i_H && (unsigned)w < (unsigned)E.c2i_W)
					atomicAdd(o + ((size_t)c * HW + h * E.c2i_W + w), alpha * __uint_as_float(v[j]));
			}
		}
		return;
	}
	if (fast != 0 && n0 + EPI_COLS <= E.N) {
		float* dst = (float*)outp + (size_t)n0 * ncs;
		if (fast == 1) {
			#pragma unroll
			for (int j = 0; j < EPI_COLS; j++) dst[(size_t)j * ncs] = fmaf(alpha, __uint_as_float(v[j]), bias_m);
		} else if (fast == 2) {
			const float* bn = (const float*)biasp + n0;
			#pragma unroll
			for (int j = 0; j < EPI_COLS; j++) dst[(size_t)j * ncs] = fmaf(alpha, __uint_as_float(v[j]), __ldg(bn + j));
		} else {
			#pragma unroll
			for (int j = 0; j < EPI_COLS; j++) atomicAdd(dst + (size_t)j * ncs, alpha * __uint_as_float(v[j]));
		}
		return;
	}
	const int kind = E.out_kind;
	if (kind != OUT_F32 && !E.atomic && E.beta == 0.0f && n0 + EPI_COLS <= E.N) {
		// 16-bit store path (fprop / dgrad of half / bfloat16 tensors)
		#pragma unroll
		for (int j = 0; j < EPI_COLS; j++) {
			float r = alpha * __uint_as_float(v[j]);
			if (E.bias_mode == 1) r += out_load(biasp, (size_t)(n0 + j), kind);
			else r += bias_m;
			const size_t idx = (size_t)(n0 + j) * ncs;
			if (kind == OUT_F16) ((__half*)outp)[idx] = __float2half_rn(r);
			else ((__nv_bfloat16*)outp)[idx] = __float2bfloat16_rn(r);
		}
		return;
	}
	#pragma unroll
	for (int j = 0; j < EPI_COLS; j++) {
		if (n0 + j < E.N) {
			float r = alpha * __uint_as_float(v[j]);
			if (addbias) {
				if (E.bias_mode == 1) r += out_load(biasp, (size_t)(n0 + j), kind);
				else if (E.bias_mode == 2) r += bias_m;
			}
			const size_t idx = (size_t)(n0 + j) * ncs;
			if (E.atomic) {
				atomicAdd((float*)outp + idx, r);            // out (fp32) was pre-scaled by beta on the host side
			} else {
				if (E.beta != 0.0f) r += E.beta * out_load(outp, idx, kind);
				out_store(outp, idx, r, kind);
			}
		}
	}
}

// Staged epilogue of one tile (Epilogue::staged; the 8 epilogue warps together).  Rounds of 32 columns: phase 1 moves each warp's
// 16 columns x 32 rows from TMEM to the staging tile stg[column][row] (rows contiguous, 16 KB); phase 2 lets warp ew write
// columns 4*ew .. 4*ew+3, each as the 512 contiguous bytes of 128 rows -- one 16-byte-lane request (VEC: row offsets contiguous
// and 16-byte aligned) or four adjacent 128-byte requests issued back to back.  Measured on the store pattern alone
// (tools/ubench/store_pattern.cu): 4.9 TB/s against 3.0 TB/s for 128-byte requests scattered over the channel planes.
template <bool VEC>
__device__ __forceinline__ void staged_tile(const Epilogue& E, int row0, int col0, int group, int ncols, uint32_t tmem_d, uint32_t stg, int ew,
											 int lane)
{
	const int half = ew >> 2, lg = ew & 3;
	constexpr int NOFF = VEC ? 1 : 4;
	int roff[NOFF];
	bool rval[NOFF];
	#pragma unroll
	for (int e = 0; e < NOFF; e++) {
		const int mm = row0 + (VEC ? 4 * lane : 32 * e + lane);
		rval[e] = mm < E.M;
		int a0, a1, a2;
		split3((uint32_t)(rval[e] ? mm : 0), E.md12, E.md2, a0, a1, a2);
		roff[e] = a0 * E.ms0 + a1 * E.ms1 + a2 * E.ms2;
	}
	float* outg = (float*)E.out + (long long)group * E.group_stride + (long long)col0 * E.ncs;
	const float* biasg = (E.bias && E.bias_mode == 1) ? (const float*)E.bias + (long long)group * E.bias_group_stride + col0 : nullptr;
	const float alpha = E.alpha;
	const uint32_t wr = stg + (uint32_t)(((half * EPI_COLS) * BM + lg * 32 + lane) * 4);
	const uint32_t rd = stg + (uint32_t)(((ew * 4) * BM + (VEC ? 4 * lane : lane)) * 4);

	uint32_t v[EPI_COLS];
	if (half * EPI_COLS < ncols) tmem_ld16(tmem_d + (uint32_t)(half * EPI_COLS), v);
	#pragma unroll 1
	for (int c0 = 0; c0 < ncols; c0 += 32) {
		if (c0 + half * EPI_COLS < ncols) {
			tmem_wait_ld(v);
			#pragma unroll
			for (int j = 0; j < EPI_COLS; j++) sts32(wr + j * BM * 4, v[j]);
		}
		named_bar_sync(5, NEPI_WARPS_GEMM * 32);
		if (c0 + 32 + half * EPI_COLS < ncols) tmem_ld16(tmem_d + (uint32_t)(c0 + 32 + half * EPI_COLS), v);
		#pragma unroll 1
		for (int i = 0; i < 4; i++) {
			const int cl = c0 + ew * 4 + i;
			if (cl >= ncols) break;
			const float b = biasg ? __ldg(biasg + cl) : 0.0f;
			float* dstc = outg + (long long)cl * E.ncs;
			if (VEC) {
				float x0, x1, x2, x3;
				asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3) : "r"(rd + i * BM * 4));
				if (rval[0])
					*reinterpret_cast<float4*>(dstc + roff[0]) = make_float4(fmaf(alpha, x0, b), fmaf(alpha, x1, b), fmaf(alpha, x2, b), fmaf(alpha, x3, b));
			} else {
				float x[4];
				#pragma unroll
				for (int e = 0; e < 4; e++) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[e]) : "r"(rd + (i * BM + 32 * e) * 4));
				#pragma unroll
				for (int e = 0; e < 4; e++)
					if (rval[e]) dstc[roff[e]] = fmaf(alpha, x[e], b);
			}
		}
		named_bar_sync(5, NEPI_WARPS_GEMM * 32);
	}
}

// ------------------------------------------------------------------------------------------ the kernel
template <int BN, int AMODE, int BMODE, bool CDIV, bool H16>
__global__ void __launch_bounds__(NTHREADS, 1) umma_gemm_kernel(const __grid_constant__ GemmParams p, const __grid_constant__ CUtensorMap tmapB,
																const __grid_constant__ CUtensorMap tmapA)
{
	using C = Cfg<BN>;
	using EL = typename std::conditional<H16, uint16_t, float>::type;      // operand element as the producers see it
	constexpr int BKE = H16 ? BK16 : BK;                                   // elements per k-block (one 128-byte row)
	constexpr bool B_TMA = BMODE == MODE_TMA;
	constexpr bool A_KPT = AMODE == MODE_K_POS_TMA, B_KPT = BMODE == MODE_K_POS_TMA;   // plane operands through the copy engine
	constexpr bool A_MNT = AMODE == MODE_MN_TMA;
	static_assert(!A_MNT || !H16, "MODE_MN_TMA: float tensors only");
	extern __shared__ uint8_t smem_raw[];
	const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
	const uint32_t tables = smem0 + C::STAGES * C::STAGE_BYTES;
	const uint32_t bars = tables + NGROUPS * C::TABLE_BYTES;      // full[STAGES], empty[STAGES], acc_full[2], acc_empty[2], tmem ptr
	const uint32_t bar_full = bars, bar_empty = bars + 8 * C::STAGES;
	const uint32_t bar_accfull = bars + 16 * C::STAGES, bar_accempty = bar_accfull + 16;
	const uint32_t tmem_slot = bar_accempty + 16;
	const uint32_t bar_tma = tmem_slot + 16;                  // [STAGES], FIXUP only: the copied activation tiles have landed
	static_assert(16 * C::STAGES + 48 + 8 * C::STAGES <= C::BAR_BYTES, "barrier area too small");
	// Float operands that come through the copy engine would reach the tensor core un-rounded (the tf32 MMA ignores the low 13
	// mantissa bits: truncation, -3.5e-4 per operand on average).  The producer group that owns the k-block therefore rounds the
	// landed tile in place before it hands the stage to the MMA thread: bits + 0x1000, i.e. round to nearest, ties away from zero,
	// once the hardware drops the low bits -- the value cvt.rna.tf32.f32 gives the gathering producers, so both routes multiply
	// the same numbers.  16-bit operands are exact and skip this.
	constexpr bool FIXUP = !H16 && (A_KPT || B_KPT || A_MNT);

	const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
	const int lane = threadIdx.x & 31;
	const int total_work = p.tiles_m * p.tiles_n * p.groups * p.splits;

	if (warp == MMA_WARP) {
		if (lane == 0) {
			for (int s = 0; s < C::STAGES; s++) {
				mbar_init(bar_full + 8 * s, NPROD_WARPS + (B_TMA ? 1 : 0) + (FIXUP ? 0 : (A_KPT ? 1 : 0) + (B_KPT ? 1 : 0) + (A_MNT ? 1 : 0)));
				if (FIXUP) mbar_init(bar_tma + 8 * s, (A_KPT ? 1 : 0) + (B_KPT ? 1 : 0) + (A_MNT ? 1 : 0));
				mbar_init(bar_empty + 8 * s, 1);
			}
			for (int a = 0; a < 2; a++) {
				mbar_init(bar_accfull + 8 * a, 1);
				mbar_init(bar_accempty + 8 * a, NEPI_WARPS_GEMM);
			}
			fence_barrier_init();
		}
		__syncwarp();
		tmem_alloc(tmem_slot, C::TMEM_COLS);
	}

	tc_fence_before();
	__syncthreads();
	tc_fence_after();

	uint32_t tmem_base;
	asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

	if (warp < NPROD_WARPS_ALL) {
		// ===================== producers =====================
		setmaxnreg_inc<REGS_PROD>();
		const int grp = warp / NPROD_WARPS, gw = warp % NPROD_WARPS;       // producer group, warp within it
		typename ProducerSel<BM, AMODE, CDIV, H16>::type prodA;
		typename ProducerSel<BN, BMODE, false, H16>::type prodB;
		using PA = decltype(prodA);
		using PB = decltype(prodB);
		float va[PA::NV];
		float vb[PB::NV];
		const uint32_t tset = tables + (uint32_t)grp * C::TABLE_BYTES;

		// cursor over the CTA's k-block sequence (work units x their k-blocks); this group owns every NGROUPS-th k-block
		int lwork = blockIdx.x, lkb = 0, inited = -1;
		int stage = 0;
		uint32_t phase = 0;
		Work lw{};
		bool lvalid = lwork < total_work;
		if (lvalid) { lw = decode_work(p, lwork); lkb = lw.kb_begin; }
		auto advance = [&](int n) {
			// n k-blocks further: whole work units are skipped without decoding them
			while (lvalid && n > 0) {
				const int left = lw.kb_end - lkb;
				if (n < left) { lkb += n; break; }
				n -= left;
				lwork += gridDim.x;
				lvalid = lwork < total_work;
				if (lvalid) { lw = decode_work(p, lwork); lkb = lw.kb_begin; }
			}
		};
		auto advance_stage = [&](int n) {
			stage += n;
			if (stage >= C::STAGES) { stage -= C::STAGES; phase ^= 1; }
		};
		// A group may run ahead of the MMAs by less than one lap of the stage ring only (the empty-barrier wait is a parity
		// test): its consecutive k-blocks are NGROUPS apart, so NGROUPS <= STAGES.  (Measured: pairs of k-blocks per group --
		// 64 loads in flight per thread -- change nothing, and need an 8-stage ring to be safe.)
		static_assert(C::STAGES >= NGROUPS, "a producer group would lap the stage ring");
		advance(grp);
		advance_stage(grp);
		const EL* baseA = (const EL*)p.A.ptr;
		const EL* baseB = (const EL*)p.B.ptr;
		PZ_TL_DECL(8)
		while (lvalid) {
			PZ_TL(6)
			if (inited != lwork) {
				// first k-block this group sees of a work unit: per-tile producer state + the group's smem tables
				named_bar_sync(1 + grp, NPROD);                       // every warp of the group is done reading the old tables
				prodA.init(p.A, lw.m_tile * BM, gw, lane, tset);
				prodB.init(p.B, lw.n_tile * BN, gw, lane, tset + BM * 16);
				baseA = (const EL*)p.A.ptr + (long long)lw.group * p.A.group_stride;
				baseB = (const EL*)p.B.ptr + (long long)lw.group * p.B.group_stride;
				named_bar_sync(1 + grp, NPROD);
				inited = lwork;
			}
			PZ_TL(0)
			prodA.load(p.A, baseA, lkb, va);
			prodB.load(p.B, baseB, lkb, vb);
			PZ_TL(1)
			mbar_wait(bar_empty + 8 * stage, phase ^ 1);
			PZ_TL(2)
			PZ_TL_TOUCH(va[PA::NV - 1])
			PZ_TL(3)
			const uint32_t tileA = smem0 + stage * C::STAGE_BYTES;
			if (B_TMA) {
				if (gw == 0 && lane == 0) {
					if (p.debug_skip & 2) mbar_arrive(bar_full + 8 * stage);
					else {
						mbar_arrive_expect_tx(bar_full + 8 * stage, BN * 128);
						tma_load_2d(tileA + BM * 128, &tmapB, lkb * BKE, lw.group * p.tma_rows_per_group + lw.n_tile * BN, bar_full + 8 * stage);
					}
				}
			}
			if (A_KPT || B_KPT) {
				if (gw == 0 && lane == 0) {
					// k-block -> (image, first position); rows = channels of the (only) group
					const uint32_t img = fdiv((uint32_t)lkb, p.A.kbdiv);
					const int pos = (int)((uint32_t)lkb - img * p.A.kbdiv.d) * BKE;
					const uint32_t bar_act = (FIXUP ? bar_tma : bar_full) + 8 * stage;
					if (A_KPT) {
						mbar_arrive_expect_tx(bar_act, BM * 128);
						tma_load_3d(tileA, &tmapA, pos, lw.m_tile * BM, (int)img, bar_act);
					}
					if (B_KPT) {
						mbar_arrive_expect_tx(bar_act, BN * 128);
						tma_load_3d(tileA + BM * 128, &tmapB, pos, lw.n_tile * BN, (int)img, bar_act);
					}
				}
			}
			if (A_MNT) {
				if (gw == 0 && lane == 0) {
					const int chunk0 = lw.m_tile * (BM / 32);
					const int nch = min(BM / 32, p.total_chunks - chunk0);       // chunks past the last image are not fetched (nor stored)
					const uint32_t bar_act = (FIXUP ? bar_tma : bar_full) + 8 * stage;
					mbar_arrive_expect_tx(bar_act, (uint32_t)nch * 4096u);
					#pragma unroll
					for (int j = 0; j < BM / 32; j++) {
						if (j < nch) {
							const uint32_t img = fdiv((uint32_t)(chunk0 + j), p.fd_img_chunks);
							const int pix0 = (int)((uint32_t)(chunk0 + j) - img * (uint32_t)p.img_chunks) * 32;
							tma_load_3d(tileA + j * 4096, &tmapA, pix0, lkb * BKE, (int)img, bar_act);
						}
					}
				}
			}
			if (FIXUP) {
				// the copied tiles of this stage: wait for them, round them in place (16 bytes per thread per step, conflict-free)
				mbar_wait(bar_tma + 8 * stage, phase);
				const uint32_t lo = (A_KPT || A_MNT) ? 0u : (uint32_t)BM * 128u;
				const uint32_t hi = B_KPT ? (uint32_t)(BM + BN) * 128u : (uint32_t)BM * 128u;
				#pragma unroll 4
				for (uint32_t off = lo + (uint32_t)(gw * 32 + lane) * 16u; off < hi; off += NPROD * 16u) {
					uint32_t a, b, c, d;
					asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(tileA + off) : "memory");
					sts128(tileA + off, a + 0x1000u, b + 0x1000u, c + 0x1000u, d + 0x1000u);
				}
			}
			prodA.store(tileA, va);
			prodB.store(tileA + BM * 128, vb);
			PZ_TL(4)
			fence_async_smem();
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_full + 8 * stage);
			PZ_TL(5)
			advance(NGROUPS);
			advance_stage(NGROUPS);
		}
		PZ_TL_FLUSH(0, 8)
	} else if (warp < EPI_WARP0) {
		// ===================== MMA issuer (one thread of warp 16) =====================
		setmaxnreg_dec<REGS_MMA>();
		if (warp == MMA_WARP) {
			const uint32_t idesc = H16 ? make_idesc_f16(BM, BN, p.ab_bf16) : (make_idesc_tf32(BM, BN) | (A_MNT ? 1u << 15 : 0u));
			int stage = 0;
			uint32_t phase = 0;
			int as = 0;
			uint32_t aphase = 0;
			PZ_TL_DECL(4)
			for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
				const Work w = decode_work(p, work);
				PZ_TL(0)
				mbar_wait(bar_accempty + 8 * as, aphase ^ 1);        // epilogue has drained this accumulator buffer
				tc_fence_after();
				PZ_TL(1)
				const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
				for (int kb = w.kb_begin; kb < w.kb_end; kb++) {
					PZ_TL(3)
					mbar_wait(bar_full + 8 * stage, phase);
					tc_fence_after();
					PZ_TL(2)
					if (lane == 0) {
						const uint32_t tileA = smem0 + stage * C::STAGE_BYTES;
						const uint64_t da = A_MNT ? make_smem_desc_mn(tileA, p.debug_skip) : make_smem_desc(tileA);
						const uint64_t db = make_smem_desc(tileA + BM * 128);
						constexpr int ASTEP = A_MNT ? 64 : 2;       // MN-major A: the next 8 k are the next 1024 bytes
						#pragma unroll
						for (int kk = 0; kk < 4; kk++) {       // 8 tf32 / 16 halves = 32 bytes per MMA: +2 in the (addr >> 4) field
							if (p.debug_skip & 4) continue;
							if (H16) umma_f16(tmem_d, da + 2 * kk, db + 2 * kk, idesc, (kb > w.kb_begin || kk > 0) ? 1u : 0u);
							else umma_tf32(tmem_d, da + ASTEP * kk, db + 2 * kk, idesc, (kb > w.kb_begin || kk > 0) ? 1u : 0u);
						}
						umma_commit(bar_empty + 8 * stage);
					}
					__syncwarp();
					if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
				}
				if (lane == 0) umma_commit(bar_accfull + 8 * as);
				__syncwarp();
				if (++as == 2) { as = 0; aphase ^= 1; }
				PZ_TL(3)
			}
			PZ_TL_FLUSH(8, 4)
			tc_fence_before();
		}
	} else {
		// ===================== epilogue (4 warps; warp w may touch TMEM lanes 32*(w%4) .. +31) =====================
		setmaxnreg_dec<REGS_EPI>();
		const Epilogue& E = p.E;
		const int lg = warp & 3;
		int as = 0;
		uint32_t aphase = 0;
		// staged stores reuse the producer tables, which the channel-ordered / 16-byte producers and the TMA operand do not use
		constexpr bool STAGED_OK = (AMODE == MODE_MN_CHAN || AMODE == MODE_MN_VEC) && BMODE == MODE_TMA && !H16;
		static_assert(NGROUPS * C::TABLE_BYTES >= 32 * BM * 4, "the staging tile does not fit the table area");
		PZ_TL_DECL(2)
		// (two separate tile loops: sharing one would keep the state of both paths live across it)
		if (STAGED_OK && E.staged) {
			for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
				const Work w = decode_work(p, work);
				PZ_TL(1)
				mbar_wait(bar_accfull + 8 * as, aphase);
				tc_fence_after();
				PZ_TL(0)
				const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN) + ((uint32_t)(lg * 32) << 16);
				const int ncols = (p.debug_skip & 1) ? 0 : min(BN, E.N - w.n_tile * BN);
				if (E.staged == 2) staged_tile<true>(E, w.m_tile * BM, w.n_tile * BN, w.group, ncols, tmem_d, tables, warp - EPI_WARP0, lane);
				else staged_tile<false>(E, w.m_tile * BM, w.n_tile * BN, w.group, ncols, tmem_d, tables, warp - EPI_WARP0, lane);
				tc_fence_before();
				__syncwarp();
				if (lane == 0) mbar_arrive(bar_accempty + 8 * as);
				if (++as == 2) { as = 0; aphase ^= 1; }
			}
		} else
		for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
			const Work w = decode_work(p, work);
			PZ_TL(1)
			int m = w.m_tile * BM + lg * 32 + lane;
			bool mvalid = m < E.M;
			int m0, m1, m2;
			if (A_MNT) {
				// chunked rows over dense planes: this warp's 32 rows are one chunk = 32 positions of one image, contiguous in the output too
				const int chunk = w.m_tile * (BM / 32) + lg;
				const uint32_t img = fdiv((uint32_t)chunk, p.fd_img_chunks);
				const int pix = (int)((uint32_t)chunk - img * (uint32_t)p.img_chunks) * 32 + lane;
				mvalid = chunk < p.total_chunks && pix < (int)E.md12.d;
				m0 = (int)img; m1 = 0; m2 = mvalid ? pix : 0;
				m = (int)img * (int)E.md12.d + m2;
			} else
				split3((uint32_t)(mvalid ? m : 0), E.md12, E.md2, m0, m1, m2);
			const int oes = E.out_kind == OUT_F32 ? 4 : 2;       // bytes per output / bias element
			char* outp = (char*)E.out + ((long long)w.group * E.group_stride + ((long long)m0 * E.ms0 + (long long)m1 * E.ms1 + (long long)m2 * (A_MNT ? 1 : E.ms2))) * oes;
			const char* biasp = E.bias ? (const char*)E.bias + (long long)w.group * E.bias_group_stride * oes : nullptr;
			const bool addbias = !E.atomic || w.split == 0;      // with split-K the bias is contributed once
			const float bias_m = (E.bias_mode == 2 && mvalid && addbias) ? out_load(biasp, (size_t)m, E.out_kind) : 0.0f;

			mbar_wait(bar_accfull + 8 * as, aphase);
			tc_fence_after();
			PZ_TL(0)
			const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN) + ((uint32_t)(lg * 32) << 16);
			const int ncols = min(BN, E.N - w.n_tile * BN);       // valid columns of this tile (> 0)
			// fast paths (straight-line, 2 - 3 instructions per element): plain store and split-K red.add
			const int fast = E.c2i ? 4 : (E.out_kind != OUT_F32 ? 0 : (E.atomic ? ((addbias && E.bias_mode) ? 0 : 3) : (E.beta != 0.0f ? 0 : (E.bias_mode == 1 ? 2 : 1))));
			const int hb = m1 * E.c2i_sh - E.c2i_ph, wb = m2 * E.c2i_sw - E.c2i_pw;

			// this warp's chunks: columns half*16 + 32*i (the quadrant's other warp takes the chunks in between)
			const int half = (warp - EPI_WARP0) >> 2;
			constexpr int CSTRIDE = 2 * EPI_COLS;
			uint32_t v0[EPI_COLS], v1[EPI_COLS];
			const int cfirst = half * EPI_COLS;
			if (cfirst < ncols) tmem_ld16(tmem_d + (uint32_t)cfirst, v0);
			#pragma unroll 1
			for (int c0 = cfirst; c0 < ((p.debug_skip & 1) ? 0 : ncols); c0 += 2 * CSTRIDE) {
				tmem_wait_ld(v0);
				if (c0 + CSTRIDE < ncols) tmem_ld16(tmem_d + (uint32_t)(c0 + CSTRIDE), v1);
				epilogue_chunk(E, v0, outp, biasp, bias_m, mvalid, addbias, w.n_tile * BN + c0, fast, hb, wb);
				if (c0 + CSTRIDE < ncols) {
					tmem_wait_ld(v1);
					if (c0 + 2 * CSTRIDE < ncols) tmem_ld16(tmem_d + (uint32_t)(c0 + 2 * CSTRIDE), v0);
					epilogue_chunk(E, v1, outp, biasp, bias_m, mvalid, addbias, w.n_tile * BN + c0 + CSTRIDE, fast, hb, wb);
				}
			}
			tc_fence_before();
			__syncwarp();
			if (lane == 0) mbar_arrive(bar_accempty + 8 * as);
			if (++as == 2) { as = 0; aphase ^= 1; }
		}
		PZ_TL(1)
		PZ_TL_FLUSH(12, 2)
	}

	tc_fence_before();
	__syncthreads();
	if (warp == MMA_WARP) {
		tc_fence_after();
		tmem_dealloc(tmem_base, C::TMEM_COLS);
	}
}

// host-side launcher (defined in pz_gemm.cu).  `bn` = 0 lets the launcher pick the tile width; tmap_src (MODE_TMA only) is
// the prepared filter: fp32 [tma_rows][tma_kpad], tf32-rounded, zero-padded, 16-byte aligned.
struct TmaSource {
	const void* ptr;
	long long rows, kpad;
};
// dtype: PZ_F32 (tf32 products), PZ_F16 or PZ_BF16 -- the element type of both operands
// MODE_K_POS_TMA / MODE_MN_TMA operand: planes [images][chans][plane] of whole 16-byte units on a 16-byte aligned base
struct PlaneTma {
	const void* ptr;
	long long plane, chans, images;
};
int launch(GemmParams& p, int dtype, int bn, int amode, int bmode, bool cdiv, int groups, const TmaSource* tma, cudaStream_t stream,
		   const PlaneTma* planeA = nullptr, const PlaneTma* planeB = nullptr);
inline int elems_per_kblock(int dtype) { return dtype == PZ_F32 ? BK : BK16; }
// ---- halo path (pz_halo.cu): stride-1 R x S convolution with the activation operand staged once per channel block
struct HaloGeometry {
	const void* src;             // (N, C_total, Hs, Ws) source of the correlation: x (fprop) or dy (dgrad)
	int N, Hs, Ws, ph, pw;       // zero padding added on every side of the source
	int R, S;
	int chans, out_chans;        // reduction / output channels per group
	int out_h, out_w;            // must equal Hs + 2*ph - R + 1, Ws + 2*pw - S + 1
	int groups;
	long long img_stride, chan_stride, group_stride;   // source strides in elements
};
int launch_halo(const HaloGeometry& g, int dtype, const TmaSource& tsrc, const Epilogue& E, double alg_flops, double alg_bytes,
				cudaStream_t stream);
int make_filter_tmap(CUtensorMap* tmap, int dtype, const TmaSource& tma, int bn);
int finalize16(int dtype, void* out, int64_t ldo, const float* acc, int64_t rows, int64_t cols, float beta, cudaStream_t stream);
inline int out_kind_of(int dtype) { return dtype == PZ_F32 ? OUT_F32 : (dtype == PZ_F16 ? OUT_F16 : OUT_BF16); }
int pick_bn(int n, long long m_rows, int kblocks, int groups, int max_bn);
float* scratch(size_t bytes);                   // library-owned, stream-ordered scratch (prepared filters)

// helpers to build operands
Operand dense_k(const void* ptr, int rows, int kdim, long long ld);      // element (row,k) at ptr[row*ld + k]
Operand dense_mn(const void* ptr, int rows, int kdim, long long ld, int bke);   // element (row,k) at ptr[k*ld + row] (MODE_MN_CHAN)

}  // namespace pzumma
