// pz_gemm.cu -- launcher of the tcgen05 GEMM engine (pz_umma.cuh) and the pz_gemm entry point.
// Replaces cublasGemmEx as used by CuBlas_Context_gemm (reference Cuda/Source/Libs/CuBlas.c:327-403).
#include "pz_umma.cuh"

#include <cstdlib>

namespace pzumma {

Operand dense_k(const void* ptr, int rows, int kdim, long long ld)
{
	Operand o{};
	o.ptr = ptr;
	o.rd12 = make_fastdiv(1);  // r0 = row
	o.rd2 = make_fastdiv(0);
	o.kd12 = make_fastdiv(0);  // k2 = k
	o.kd2 = make_fastdiv(0);
	o.rs0 = (int)ld;
	o.ks0 = 0;
	o.ah = o.bh = o.ch = 0;
	o.aw = 0; o.bw = 1; o.cw = 0;
	o.H = 1; o.W = kdim; o.Wd = 0;
	o.cdh = o.cdw = 1;
	o.rows = rows; o.kdim = kdim;
	o.R = o.S = 1;
	o.group_stride = 0;
	return o;
}

// MN-contiguous dense operand for the channel-ordered producer: rows are "positions" of a 1 x rows map, k are "channels" with
// stride ld, one filter tap
Operand dense_mn(const void* ptr, int rows, int kdim, long long ld, int bke)
{
	Operand o{};
	o.ptr = ptr;
	o.rd12 = make_fastdiv(0);  // r2 = row
	o.rd2 = make_fastdiv(0);
	o.kd12 = make_fastdiv(1);
	o.kd2 = make_fastdiv(1);   // one tap: s = t
	o.rs0 = 0;
	o.ks0 = (int)ld;
	o.ah = o.bh = o.ch = 0;
	o.aw = 1; o.bw = 0; o.cw = 0;
	o.H = 1; o.W = rows; o.Wd = 0;
	o.cdh = o.cdw = 1;
	o.rows = rows; o.kdim = kdim;
	o.R = o.S = 1;
	o.group_stride = 0;
	o.chans = kdim;
	o.kbdiv = make_fastdiv((uint32_t)pz_cdiv(kdim, bke));
	return o;
}

// Tile width: the widest tile reads the gathered operand the fewest times, but few wide tiles leave SMs idle in the last
// round of the persistent schedule.  Cost model (clock cycles per SM, rough): a k-block costs max(MMA, gather) and a tile
// ends with a TMEM drain + stores; the launch takes ceil(units / SMs) rounds.
int pick_bn(int n, long long m_rows, int kblocks, int groups, int max_bn)
{
	static const int forced = [] { const char* e = getenv("PZ_FORCE_BN"); return e ? atoi(e) : 0; }();
	if (forced == 64 || forced == 128 || forced == 256) return forced <= max_bn ? forced : max_bn;
	const long long sms = pz_num_sms();
	int best = 64;
	double best_cost = 1e300;
	for (int bn = 64; bn <= max_bn; bn *= 2) {
		const long long units = pz_cdiv(m_rows, BM) * pz_cdiv(n, bn) * groups;
		const long long rounds = pz_cdiv(units, sms);
		const double per_kb = 2.0 * bn > 320.0 ? 2.0 * bn : 320.0;
		const double cost = (double)rounds * ((double)kblocks * per_kb + 10.0 * bn + 1500.0);
		if (cost < best_cost * 0.97) { best_cost = cost; best = bn; }
		if (bn >= n) break;
	}
	return best;
}

float* scratch(size_t bytes) { return (float*)pz_scratch(bytes); }

template <int BN, int AM, int BMODE, bool CDIV, bool H16>
static int launch_inst(const GemmParams& p, const CUtensorMap& tmap, const CUtensorMap& tmapA, int grid, cudaStream_t stream)
{
	auto kern = umma_gemm_kernel<BN, AM, BMODE, CDIV, H16>;
	static bool configured = false;
	if (!configured) {
		PZ_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<BN>::SMEM_BYTES));
		configured = true;
	}
	{
		const bool hbm = p.alg_bytes > 0.0 && p.alg_flops / p.alg_bytes < kPzRidgeFlopPerByte;
		PzProfScope prof(hbm ? PZ_PROF_GEMM_HBM : PZ_PROF_GEMM, stream, p.alg_flops, p.alg_bytes);
		kern<<<grid, NTHREADS, Cfg<BN>::SMEM_BYTES, stream>>>(p, tmap, tmapA);
	}
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
								 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
								 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmapEncodeFn tmap_encoder()
{
	// resolved through the runtime so that the library has no link-time dependency on libcuda.so
	static TmapEncodeFn encode = nullptr;
	if (!encode) {
		void* fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
			encode = (TmapEncodeFn)fn;
	}
	return encode;
}

// 3-d tensor map over float planes [images][chans][plane] (MODE_K_POS_TMA): box = 32 positions x `rows` channels of one image,
// 128-byte swizzle -- the K-major tile the MMA descriptors expect, positions past the plane read as zero
static int make_plane_tmap(CUtensorMap* tmap, int dtype, const PlaneTma& t, int rows, bool atom32 = false)
{
	const size_t es = dtype == PZ_F32 ? 4 : 2;
	const CUtensorMapDataType dt = dtype == PZ_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
								   : (dtype == PZ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
	PZ_REQUIRE(t.ptr != nullptr && ((uintptr_t)t.ptr & 15) == 0 && (t.plane * es) % 16 == 0 && t.chans >= rows && rows <= 256, "bad plane TMA source");
	TmapEncodeFn encode = tmap_encoder();
	PZ_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available in this driver");
	cuuint64_t dims[3] = {(cuuint64_t)t.plane, (cuuint64_t)t.chans, (cuuint64_t)t.images};
	cuuint64_t strides[2] = {(cuuint64_t)t.plane * es, (cuuint64_t)t.plane * (cuuint64_t)t.chans * es};
	cuuint32_t box[3] = {(cuuint32_t)elems_per_kblock(dtype), (cuuint32_t)rows, 1};
	cuuint32_t estr[3] = {1, 1, 1};
	memset(tmap, 0, sizeof(*tmap));
	// atom32: 32-byte chunks swizzled inside the 128-byte rows (the layout of an MN-major tf32 MMA operand)
	CUresult r = encode(tmap, dt, 3, (void*)t.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
						atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
						CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		pz_set_error(PZ_ERR_CUDA, "cuTensorMapEncodeTiled (planes) failed (%d)", (int)r);
		return PZ_ERR_CUDA;
	}
	return PZ_OK;
}

// 2-d tensor map over the prepared filter: rows of kpad elements, box = one k-block x bn rows, 128-byte swizzle
int make_filter_tmap(CUtensorMap* tmap, int dtype, const TmaSource& tma, int bn)
{
	const int bke = elems_per_kblock(dtype);
	const size_t es = dtype == PZ_F32 ? 4 : 2;
	PZ_REQUIRE(tma.ptr != nullptr && ((uintptr_t)tma.ptr & 15) == 0 && (tma.kpad * es) % 16 == 0, "bad TMA source");
	cuuint64_t dims[2] = {(cuuint64_t)tma.kpad, (cuuint64_t)tma.rows};
	cuuint64_t strides[1] = {(cuuint64_t)tma.kpad * es};
	cuuint32_t box[2] = {(cuuint32_t)bke, (cuuint32_t)bn};
	cuuint32_t estr[2] = {1, 1};
	const CUtensorMapDataType dt = dtype == PZ_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
								   : (dtype == PZ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
	TmapEncodeFn encode = tmap_encoder();      // (the library must load, and export its symbols, on a machine without a driver)
	PZ_REQUIRE(encode != nullptr, "cuTensorMapEncodeTiled is not available in this driver");
	memset(tmap, 0, sizeof(*tmap));
	CUresult r = encode(tmap, dt, 2, (void*)tma.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
						CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) {
		pz_set_error(PZ_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
		return PZ_ERR_CUDA;
	}
	return PZ_OK;
}

int launch(GemmParams& p, int dtype, int bn, int amode, int bmode, bool cdiv, int groups, const TmaSource* tma, cudaStream_t stream,
		   const PlaneTma* planeA, const PlaneTma* planeB)
{
	const int M = p.E.M, N = p.E.N;
	if (M <= 0 || N <= 0) return PZ_OK;
	PZ_REQUIRE(dtype == PZ_F32 || dtype == PZ_F16 || dtype == PZ_BF16, "unsupported operand dtype %d", dtype);
	PZ_REQUIRE(p.kblocks > 0 && p.splits > 0 && p.kb_per_split > 0, "empty contraction");
	PZ_REQUIRE((long long)(p.splits - 1) * p.kb_per_split < p.kblocks, "empty split");
	const bool h16 = dtype != PZ_F32;
	p.ab_bf16 = dtype == PZ_BF16 ? 1 : 0;
	{
		static const int skip = [] { const char* e = getenv("PZ_DEBUG_SKIP"); return e ? atoi(e) : 0; }();
		p.debug_skip = skip;
	}
	p.tiles_m = (int)pz_cdiv(M, BM);
	p.img_chunks = p.total_chunks = 0;
	p.fd_img_chunks = make_fastdiv(1);
	if (amode == MODE_MN_TMA) {
		// rows in chunks of 32 positions of one image (the boxes of the activation tensor map are per image), 4 chunks per tile
		PZ_REQUIRE(planeA != nullptr && planeA->plane * planeA->images == M && dtype == PZ_F32 && groups == 1 && p.E.md12.d == (uint32_t)planeA->plane,
				   "bad MN-major TMA operand");
		p.img_chunks = (int)pz_cdiv(planeA->plane, 32);
		p.fd_img_chunks = make_fastdiv((uint32_t)p.img_chunks);
		PZ_REQUIRE(planeA->images * (long long)p.img_chunks < (1ll << 31), "tile grid too large");
		p.total_chunks = (int)(planeA->images * p.img_chunks);
		p.tiles_m = (int)pz_cdiv(p.total_chunks, BM / 32);
	}
	p.tiles_n = (int)pz_cdiv(N, bn);
	p.groups = groups;
	p.fd_tiles_n = make_fastdiv((uint32_t)p.tiles_n);
	p.fd_tiles_m = make_fastdiv((uint32_t)p.tiles_m);
	p.fd_splits = make_fastdiv((uint32_t)p.splits);
	const long long units = (long long)p.tiles_m * p.tiles_n * groups * p.splits;
	PZ_REQUIRE(units < (1ll << 31), "tile grid too large");
	const int grid = (int)(units < pz_num_sms() ? units : pz_num_sms());

	alignas(64) CUtensorMap tmap;
	memset(&tmap, 0, sizeof(tmap));
	if (bmode == MODE_TMA) {
		PZ_REQUIRE(tma != nullptr, "bad TMA source");
		int st = make_filter_tmap(&tmap, dtype, *tma, bn);
		if (st != PZ_OK) return st;
	}
	alignas(64) CUtensorMap tmapA;
	memset(&tmapA, 0, sizeof(tmapA));
	if (amode == MODE_MN_TMA) {
		int st = make_plane_tmap(&tmapA, dtype, *planeA, BK, true);
		if (st != PZ_OK) return st;
	}
	if (amode == MODE_K_POS_TMA || bmode == MODE_K_POS_TMA) {
		PZ_REQUIRE(groups == 1, "plane TMA operands: one group");
		if (amode == MODE_K_POS_TMA) {
			PZ_REQUIRE(planeA != nullptr, "bad plane TMA source");
			int st = make_plane_tmap(&tmapA, dtype, *planeA, BM);
			if (st != PZ_OK) return st;
		}
		if (bmode == MODE_K_POS_TMA) {
			PZ_REQUIRE(planeB != nullptr, "bad plane TMA source");
			int st = make_plane_tmap(&tmap, dtype, *planeB, bn);
			if (st != PZ_OK) return st;
		}
	}

#define PZ_INST(BNV, AMV, BMV, CD, H)                                                  \
	if (bn == BNV && amode == AMV && bmode == BMV && cdiv == CD && h16 == H)           \
		return launch_inst<BNV, AMV, BMV, CD, H>(p, tmap, tmapA, grid, stream);
#define PZ_INST_BN(AMV, BMV, CD, H) PZ_INST(64, AMV, BMV, CD, H) PZ_INST(128, AMV, BMV, CD, H)
#define PZ_INST_BN3(AMV, BMV, CD, H) PZ_INST_BN(AMV, BMV, CD, H) PZ_INST(256, AMV, BMV, CD, H)
	// ---- float32 storage, tf32 products
	PZ_INST_BN3(MODE_MN_CHAN, MODE_TMA, false, false)          // fprop, dgrad with k ordered (tap, channel)
	PZ_INST_BN3(MODE_MN_CHAN, MODE_TMA, true, false)           //   ... with more than 31 taps
	PZ_INST_BN3(MODE_MN_TAP, MODE_TMA, false, false)           // fprop, dgrad with very few channels (table-driven taps)
	PZ_INST_BN3(MODE_MN_TAP, MODE_TMA, true, false)            //   ... with more than 31 taps (7x7)
	PZ_INST_BN3(MODE_MN_GENERAL, MODE_TMA, false, false)       // fallback: offsets too large for the packed tap entries
	PZ_INST_BN3(MODE_MN_GENERAL, MODE_TMA, true, false)        // strided dgrad with dilation (exact-division gather)
	PZ_INST_BN(MODE_MN_CHAN, MODE_K_DENSE, false, false)       // GEMM NN
	PZ_INST_BN(MODE_MN_CHAN, MODE_MN_CHAN, false, false)       // GEMM TN
	PZ_INST_BN(MODE_K_DENSE, MODE_K_DENSE, false, false)       // GEMM NT
	PZ_INST_BN(MODE_K_POS_TAP, MODE_K_POS_DENSE, false, false) // wgrad
	PZ_INST_BN(MODE_K_POS_TAP, MODE_K_POS_DENSE, true, false)  //   ... with more than 31 taps
	PZ_INST_BN(MODE_K_POS_DENSE, MODE_K_POS_DENSE, false, false)   // wgrad of a 1x1 / stride-1 / un-padded filter
	PZ_INST_BN(MODE_K_GENERAL, MODE_K_DENSE, false, false)     // fallback for wgrad
	PZ_INST_BN3(MODE_MN_VEC, MODE_TMA, false, false)           // 1x1 fprop / dgrad over 16-byte aligned planes: 16-byte producer lanes
	PZ_INST_BN(MODE_K_POS_VEC, MODE_K_POS_VEC, false, false)   // wgrad of a 1x1 filter over 16-byte aligned planes
	PZ_INST_BN(MODE_K_POS_TAP, MODE_K_POS_VEC, false, false)   // wgrad: dy planes 16-byte aligned
	PZ_INST_BN(MODE_K_POS_TAP, MODE_K_POS_VEC, true, false)
	PZ_INST_BN3(MODE_MN_TMA, MODE_TMA, false, false)           // 1x1 fprop / dgrad over 16-byte aligned planes: MN-major A by the copy engine
	PZ_INST_BN(MODE_K_POS_TMA, MODE_K_POS_TMA, false, false)   // wgrad of a 1x1 filter: both operands by the copy engine (3-d tensor maps)
	PZ_INST_BN(MODE_K_POS_TAP, MODE_K_POS_TMA, false, false)   // wgrad: dy by the copy engine next to a tap-gathered x
	// ---- half / bfloat16 storage
	PZ_INST_BN3(MODE_MN_CHAN, MODE_TMA, false, true)
	PZ_INST_BN3(MODE_MN_CHAN, MODE_TMA, true, true)
	PZ_INST_BN3(MODE_MN_GENERAL, MODE_TMA, false, true)
	PZ_INST_BN3(MODE_MN_GENERAL, MODE_TMA, true, true)
	PZ_INST_BN(MODE_MN_CHAN, MODE_K_DENSE, false, true)
	PZ_INST_BN(MODE_MN_CHAN, MODE_MN_CHAN, false, true)
	PZ_INST_BN(MODE_K_DENSE, MODE_K_DENSE, false, true)
	PZ_INST_BN(MODE_K_POS_TAP, MODE_K_POS_DENSE, false, true)
	PZ_INST_BN(MODE_K_POS_TAP, MODE_K_POS_DENSE, true, true)
	PZ_INST_BN(MODE_K_POS_DENSE, MODE_K_POS_DENSE, false, true)
	PZ_INST_BN(MODE_K_POS_TMA, MODE_K_POS_TMA, false, true)
	PZ_INST_BN(MODE_K_POS_TAP, MODE_K_POS_TMA, false, true)
#undef PZ_INST_BN3
#undef PZ_INST_BN
#undef PZ_INST
	pz_set_error(PZ_ERR_UNSUPPORTED, "no GEMM instantiation for dtype=%d bn=%d amode=%d bmode=%d cdiv=%d", dtype, bn, amode, bmode, (int)cdiv);
	return PZ_ERR_UNSUPPORTED;
}

// out = beta * out + acc, converting the fp32 accumulator `acc` ([rows][cols], dense) to the 16-bit output (pitch ldo):
// the second half of a split-K contraction with half / bfloat16 storage
template <typename T>
__global__ void finalize16_kernel(T* __restrict__ out, int64_t ldo, const float* __restrict__ acc, int64_t rows, int64_t cols, float beta)
{
	const int64_t total = rows * cols;
	for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
		const int64_t r = i / cols, c = i - r * cols;
		float v = acc[i];
		T* o = out + r * ldo + c;
		if (beta != 0.0f) v += beta * (float)*o;
		*o = (T)v;
	}
}

int finalize16(int dtype, void* out, int64_t ldo, const float* acc, int64_t rows, int64_t cols, float beta, cudaStream_t stream)
{
	if (rows <= 0 || cols <= 0) return PZ_OK;
	int64_t blocks = pz_cdiv(rows * cols, 256);
	if (blocks > (int64_t)pz_num_sms() * 16) blocks = (int64_t)pz_num_sms() * 16;
	if (dtype == PZ_F16) finalize16_kernel<__half><<<(unsigned)blocks, 256, 0, stream>>>((__half*)out, ldo, acc, rows, cols, beta);
	else finalize16_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>((__nv_bfloat16*)out, ldo, acc, rows, cols, beta);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

}  // namespace pzumma

using namespace pzumma;

// Contractions with a tiny K (e.g. Linear wgrad at batch 2) gain nothing from tensor cores and would expose the raw
// 2^-11 tf32 rounding of single products; they run in exact fp32 FMAs on the CUDA cores instead.
template <typename T>
__global__ void __launch_bounds__(256) gemm_smallk_kernel(const T* __restrict__ A, const T* __restrict__ B, T* C, int64_t M,
														  int64_t N, int K, int64_t lda, int64_t ldb, int64_t ldc, int transA, int transB,
														  float alpha, float beta, const T* __restrict__ bias)
{
	const int64_t total = M * N;
	for (int64_t idx = (int64_t)blockIdx.x * 256 + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * 256) {
		const int64_t m = idx / N, n = idx - m * N;
		float acc = 0.0f;
		for (int k = 0; k < K; k++) {
			const float a = (float)(transA ? A[k * lda + m] : A[m * lda + k]);
			const float b = (float)(transB ? B[n * ldb + k] : B[k * ldb + n]);
			acc = fmaf(a, b, acc);
		}
		float r = alpha * acc;
		if (bias) r += (float)bias[n];
		if (beta != 0.0f) r += beta * (float)C[m * ldc + n];
		C[m * ldc + n] = (T)r;
	}
}

constexpr int kSmallK = 8;

extern "C" int pz_gemm(int dtype, const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda,
					   int64_t ldb, int64_t ldc, int transA, int transB, float alpha, float beta, const void* bias,
					   void* stream)
{
	PZ_REQUIRE(dtype == PZ_F32 || dtype == PZ_F16 || dtype == PZ_BF16, "pz_gemm: unsupported dtype %d", dtype);
	PZ_REQUIRE(M >= 0 && N >= 0 && K > 0, "pz_gemm: invalid sizes M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
	PZ_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "pz_gemm: dimension too large");
	// A is stored as (transA ? K : M) rows of lda elements, B as (transB ? N : K) rows of ldb elements (32-bit index algebra)
	PZ_REQUIRE((transA ? K : M) * lda < (1ll << 31) && (transB ? N : K) * ldb < (1ll << 31) && M * ldc < (1ll << 31),
			   "pz_gemm: operand exceeds 2^31 elements");
	if (M == 0 || N == 0) return PZ_OK;
	cudaStream_t s = pz_stream(stream);
	const size_t es = dtype == PZ_F32 ? 4 : 2;

	if (K <= kSmallK || (dtype == PZ_F32 && pz_exact_fp32())) {      // tiny K, or the exact mode: fp32 FMAs on the CUDA cores
		int64_t blocks = pz_cdiv(M * N, 256);
		if (blocks > (int64_t)pz_num_sms() * 16) blocks = (int64_t)pz_num_sms() * 16;
		if (dtype == PZ_F32)
			gemm_smallk_kernel<float><<<(unsigned)blocks, 256, 0, s>>>((const float*)A, (const float*)B, (float*)C, M, N, (int)K, lda, ldb, ldc,
																	   transA, transB, alpha, beta, (const float*)bias);
		else if (dtype == PZ_F16)
			gemm_smallk_kernel<__half><<<(unsigned)blocks, 256, 0, s>>>((const __half*)A, (const __half*)B, (__half*)C, M, N, (int)K, lda, ldb,
																		ldc, transA, transB, alpha, beta, (const __half*)bias);
		else
			gemm_smallk_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, s>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)B, (__nv_bfloat16*)C,
																			   M, N, (int)K, lda, ldb, ldc, transA, transB, alpha, beta,
																			   (const __nv_bfloat16*)bias);
		pz_count_launch(1);
		PZ_LAUNCH_CHECK();
		return PZ_OK;
	}

	// TMEM lanes (engine rows) are mapped to C's contiguous dimension: D[n][m] = sum_k opB[k][n] * opA[m][k]
	GemmParams p{};
	int amode, bmode;
	const int bke = elems_per_kblock(dtype);
	PZ_REQUIRE(!(transA && transB), "pz_gemm: at most one operand may be transposed");
	PZ_REQUIRE(ldb < (1ll << 20) && lda < (1ll << 20), "pz_gemm: row pitch too large");
	if (transB) { p.A = dense_k(B, (int)N, (int)K, ldb); amode = MODE_K_DENSE; }
	else        { p.A = dense_mn(B, (int)N, (int)K, ldb, bke); amode = MODE_MN_CHAN; }
	if (transA) { p.B = dense_mn(A, (int)M, (int)K, lda, bke); bmode = MODE_MN_CHAN; }
	else        { p.B = dense_k(A, (int)M, (int)K, lda); bmode = MODE_K_DENSE; }

	Epilogue& E = p.E;
	E.out = C;
	E.bias = bias;
	E.out_kind = out_kind_of(dtype);
	E.md12 = make_fastdiv(0);
	E.md2 = make_fastdiv(0);
	E.ms0 = 0; E.ms1 = 0; E.ms2 = 1;
	E.ncs = (int)ldc;
	E.M = (int)N;
	E.N = (int)M;
	E.alpha = alpha;
	E.beta = beta;
	E.bias_mode = bias ? 2 : 0;
	E.atomic = 0;
	E.group_stride = 0;
	E.bias_group_stride = 0;

	p.kblocks = (int)pz_cdiv(K, bke);
	p.alg_flops = 2.0 * (double)M * (double)N * (double)K;
	p.alg_bytes = (double)es * ((double)M * K + (double)K * N + (double)M * N * (beta != 0.0f ? 2.0 : 1.0));
	const int bn = pick_bn((int)M, N, p.kblocks, 1, 128);
	const long long tiles = pz_cdiv(N, BM) * pz_cdiv(M, bn);
	int splits = 1;
	if (tiles < pz_num_sms() && p.kblocks >= 16 && !(dtype != PZ_F32 && (bias != nullptr || ldc != N))) {
		splits = (int)((2ll * pz_num_sms()) / tiles);
		if (splits > p.kblocks / 4) splits = p.kblocks / 4;
		if (splits < 1) splits = 1;
	}
	p.kb_per_split = (int)pz_cdiv(p.kblocks, splits);
	p.splits = (int)pz_cdiv(p.kblocks, p.kb_per_split);

	if (p.splits > 1) {
		E.atomic = 1;
		if (dtype != PZ_F32) {
			// split-K with 16-bit storage: red.add into an fp32 scratch (bias included), then out = beta*out + acc
			float* acc = (float*)pz_scratch((size_t)M * N * sizeof(float));
			if (!acc) { pz_set_error(PZ_ERR_MEMORY, "pz_gemm: cannot allocate the split-K accumulator"); return PZ_ERR_MEMORY; }
			int st = pz_memset8(acc, 0, (size_t)M * N * 4, stream);
			if (st != PZ_OK) return st;
			E.out = acc;
			E.out_kind = OUT_F32;
			E.ncs = (int)N;
			E.beta = 0.0f;
			st = launch(p, dtype, bn, amode, bmode, false, 1, nullptr, s);
			if (st != PZ_OK) return st;
			return finalize16(dtype, C, ldc, acc, M, N, beta, s);
		}
		// split-K accumulates with red.add: bring C to beta*C first (rows of C may be pitched)
		if (ldc == N) {
			int st = beta == 0.0f ? pz_memset8(C, 0, (size_t)M * N * 4, stream)
								  : (beta == 1.0f ? PZ_OK : pz_scale_shift(PZ_F32, C, C, beta, 0.0f, M * N, stream));
			if (st != PZ_OK) return st;
		} else {
			for (int64_t r = 0; r < M; r++) {
				float* row = (float*)C + r * ldc;
				int st = beta == 0.0f ? pz_memset8(row, 0, (size_t)N * 4, stream)
									  : (beta == 1.0f ? PZ_OK : pz_scale_shift(PZ_F32, row, row, beta, 0.0f, N, stream));
				if (st != PZ_OK) return st;
			}
		}
	}
	return launch(p, dtype, bn, amode, bmode, false, 1, nullptr, s);
}

// instrumented builds (-DPZ_TIMELINE): copy the 32 phase counters of umma_gemm_kernel to `out` and reset them; PZ_ERR_UNSUPPORTED otherwise
extern "C" int pz_debug_timeline(unsigned long long* out)
{
#ifdef PZ_TIMELINE
	PZ_CHECK_CUDA(cudaDeviceSynchronize());
	PZ_CHECK_CUDA(cudaMemcpyFromSymbol(out, pzumma::pz_timeline, sizeof(unsigned long long) * 32));
	unsigned long long zero[32] = {};
	PZ_CHECK_CUDA(cudaMemcpyToSymbol(pzumma::pz_timeline, zero, sizeof(zero)));
	return PZ_OK;
#else
	(void)out;
	pz_set_error(PZ_ERR_UNSUPPORTED, "pz_debug_timeline: the library was built without -DPZ_TIMELINE");
	return PZ_ERR_UNSUPPORTED;
#endif
}
