// pz_conv.cu -- 2-D convolution forward / backward-data / backward-filter as implicit GEMMs on the
// tcgen05 engine of pz_umma.cuh, operating directly on the reference's NCHW tensors.
//
// Replaces cudnnConvolutionForward / BackwardData / BackwardFilter / BackwardBias as driven by
// CuDnn_Context_convNd, _convNdBackwardData, _convNdBackwardParams (reference
// Cuda/Source/Libs/CuDnn.c:397-449, 517-571, 652-712, 375-394).
//
// GEMM mapping (TMEM lanes = the contiguous dimension of the output, so epilogue stores coalesce):
//   fprop : D[(n,p,q)][k]     = sum_{c,r,s}  x[n,c,p*s-pad+r*dil,...] * w[k,c,r,s]        -> y   NCHW
//   dgrad : D[(n,h,w)][c]     = sum_{k,r,s}  dy[n,k,(h+pad-r*dil)/s,...] * w[k,c,r,s]     -> dx  NCHW
//   wgrad : D[(c,r,s)][k]     = sum_{n,p,q}  x[n,c,p*s-pad+r*dil,...] * dy[n,k,p,q]       -> dw  KCRS
// The activations are gathered by the producer warps with the index algebra of pzumma::Operand, so no
// im2col matrix, NHWC copy or zero-inserted gradient is ever materialised in HBM.
#include "pz_umma.cuh"

using namespace pzumma;

namespace {

struct Geo {
	int N, C, H, W, K, R, S, P, Q, sh, sw, ph, pw, dh, dw, G, Cg, Kg;
};

int check_desc(int dtype, const pz_conv2d_desc* d, Geo& g)
{
	PZ_REQUIRE(dtype == PZ_F32 || dtype == PZ_F16 || dtype == PZ_BF16, "conv2d: unsupported dtype %d", dtype);
	PZ_REQUIRE(d != nullptr, "conv2d: null descriptor");
	g = Geo{d->N, d->C, d->H, d->W, d->K, d->R, d->S, d->P, d->Q, d->stride_h, d->stride_w, d->pad_h, d->pad_w,
			d->dil_h, d->dil_w, d->groups, 0, 0};
	PZ_REQUIRE(g.N > 0 && g.C > 0 && g.H > 0 && g.W > 0 && g.K > 0 && g.R > 0 && g.S > 0, "conv2d: invalid tensor dims");
	PZ_REQUIRE(g.sh > 0 && g.sw > 0 && g.dh > 0 && g.dw > 0 && g.ph >= 0 && g.pw >= 0 && g.G > 0, "conv2d: invalid conv params");
	PZ_REQUIRE(g.C % g.G == 0 && g.K % g.G == 0, "conv2d: maps not divisible by groups");
	PZ_REQUIRE(g.P > 0 && g.Q > 0, "conv2d: empty output");
	// output size as the reference computes it (CuDnn.c:242-266); a deconv "postpad" < stride keeps it exact
	PZ_REQUIRE(g.H + 2 * g.ph - g.dh * (g.R - 1) - 1 >= 0 && g.W + 2 * g.pw - g.dw * (g.S - 1) - 1 >= 0,
			   "conv2d: filter does not fit the padded input");
	PZ_REQUIRE(g.P == (g.H + 2 * g.ph - g.dh * (g.R - 1) - 1) / g.sh + 1 &&
			   g.Q == (g.W + 2 * g.pw - g.dw * (g.S - 1) - 1) / g.sw + 1,
			   "conv2d: output size %dx%d inconsistent with input %dx%d", g.P, g.Q, g.H, g.W);
	g.Cg = g.C / g.G;
	g.Kg = g.K / g.G;
	long long xin = (long long)g.N * g.C * g.H * g.W, yout = (long long)g.N * g.K * g.P * g.Q;
	PZ_REQUIRE(xin < (1ll << 31) && yout < (1ll << 31), "conv2d: tensor exceeds 2^31 elements");
	return PZ_OK;
}

// ================================================================================================ exact fp32 mode
// `dnn.enableTensorOps(False)` (reference: CuDnn_Context_enableTensorOps, CuDnn.c:61-74): float32 convolutions as plain fp32
// FMAs on the CUDA cores -- the accuracy cuDNN gives the reference on this stack (its fp32 convolutions do not use TF32
// there), which the reference's own unit tests assume (np.allclose at 1e-5 relative).  A verification path, not tuned: one
// thread per output element, direct loops; the filter gradient reduces over (n, p, q) with one CTA per filter element.
__global__ void __launch_bounds__(256) exact_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w,
														  const float* __restrict__ bias, float* __restrict__ y, Geo g, long long total)
{
	for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
		const int q = (int)(i % g.Q);
		long long r0 = i / g.Q;
		const int p = (int)(r0 % g.P);
		r0 /= g.P;
		const int k = (int)(r0 % g.K), n = (int)(r0 / g.K);
		const int grp = k / g.Kg;
		float acc = 0.0f;
		for (int c = 0; c < g.Cg; c++) {
			const float* xp = x + ((long long)n * g.C + grp * g.Cg + c) * g.H * g.W;
			const float* wp = w + ((long long)k * g.Cg + c) * g.R * g.S;
			for (int r = 0; r < g.R; r++) {
				const int h = p * g.sh - g.ph + r * g.dh;
				if ((unsigned)h >= (unsigned)g.H) continue;
				for (int sx = 0; sx < g.S; sx++) {
					const int ww = q * g.sw - g.pw + sx * g.dw;
					if ((unsigned)ww < (unsigned)g.W) acc = fmaf(xp[h * g.W + ww], wp[r * g.S + sx], acc);
				}
			}
		}
		y[i] = acc + (bias ? bias[k] : 0.0f);
	}
}

__global__ void __launch_bounds__(256) exact_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w,
														  const float* __restrict__ bias, float* __restrict__ dx, Geo g, long long total)
{
	for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
		const int ww = (int)(i % g.W);
		long long r0 = i / g.W;
		const int h = (int)(r0 % g.H);
		r0 /= g.H;
		const int c = (int)(r0 % g.C), n = (int)(r0 / g.C);
		const int grp = c / g.Cg, cl = c - grp * g.Cg;
		float acc = 0.0f;
		for (int kl = 0; kl < g.Kg; kl++) {
			const int k = grp * g.Kg + kl;
			const float* dyp = dy + ((long long)n * g.K + k) * g.P * g.Q;
			const float* wp = w + ((long long)k * g.Cg + cl) * g.R * g.S;
			for (int r = 0; r < g.R; r++) {
				const int hp = h + g.ph - r * g.dh;
				if (hp < 0 || hp % g.sh != 0 || hp / g.sh >= g.P) continue;
				for (int sx = 0; sx < g.S; sx++) {
					const int wq = ww + g.pw - sx * g.dw;
					if (wq < 0 || wq % g.sw != 0 || wq / g.sw >= g.Q) continue;
					acc = fmaf(dyp[(hp / g.sh) * g.Q + wq / g.sw], wp[r * g.S + sx], acc);
				}
			}
		}
		dx[i] = acc + (bias ? bias[c] : 0.0f);
	}
}

// one CTA per filter element (k, c, r, s): fixed-order strided partial sums + tree reduction (deterministic)
__global__ void __launch_bounds__(256) exact_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
														  Geo g, float alpha, float beta)
{
	__shared__ float red[256];
	const long long e = blockIdx.x;
	const int sx = (int)(e % g.S);
	long long r0 = e / g.S;
	const int r = (int)(r0 % g.R);
	r0 /= g.R;
	const int cl = (int)(r0 % g.Cg), k = (int)(r0 / g.Cg);
	const int c = (k / g.Kg) * g.Cg + cl;
	const long long PQ = (long long)g.P * g.Q, total = (long long)g.N * PQ;
	float acc = 0.0f;
	for (long long i = threadIdx.x; i < total; i += 256) {
		const int n = (int)(i / PQ);
		const int pq = (int)(i - (long long)n * PQ);
		const int p = pq / g.Q, q = pq - p * g.Q;
		const int h = p * g.sh - g.ph + r * g.dh, ww = q * g.sw - g.pw + sx * g.dw;
		if ((unsigned)h < (unsigned)g.H && (unsigned)ww < (unsigned)g.W)
			acc = fmaf(x[(((long long)n * g.C + c) * g.H + h) * g.W + ww], dy[((long long)n * g.K + k) * PQ + pq], acc);
	}
	red[threadIdx.x] = acc;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) {
		if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
		__syncthreads();
	}
	if (threadIdx.x == 0) dw[e] = alpha * red[0] + (beta != 0.0f ? beta * dw[e] : 0.0f);
}

int g_exact_fp32 = 0;

unsigned exact_grid(long long total)
{
	long long blocks = pz_cdiv(total, 256);
	const long long cap = (long long)pz_num_sms() * 16;
	return (unsigned)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

// the fast producers pack (offset << 6 | tap) into 32 bits: `span` elements of channel / row offsets plus the tap
// offsets of one filter window must stay below 2^26, and a filter may have at most 63 taps
bool tap_entries_fit(long long span, const Geo& g)
{
	const long long taps = (long long)(g.R - 1) * g.dh * (g.W > g.Q ? g.W : g.Q) + (long long)(g.S - 1) * g.dw;
	return span + 2 * taps < (1ll << 26) && g.R * g.S <= 63;
}

// true when no tap of any output position can fall outside the un-padded input
bool taps_in_bounds(const Geo& g)
{
	return g.ph == 0 && g.pw == 0 && (long long)(g.P - 1) * g.sh + (long long)g.dh * (g.R - 1) < g.H &&
		   (long long)(g.Q - 1) * g.sw + (long long)g.dw * (g.S - 1) < g.W;
}

void set_splits(GemmParams& p, long long tiles, int min_kb_per_split)
{
	int splits = 1;
	if (tiles < 2ll * pz_num_sms()) {
		splits = (int)((3ll * pz_num_sms()) / (tiles > 0 ? tiles : 1));
		if (splits > p.kblocks / min_kb_per_split) splits = p.kblocks / min_kb_per_split;
		if (splits < 1) splits = 1;
	}
	p.kb_per_split = (int)pz_cdiv(p.kblocks, splits);
	p.splits = (int)pz_cdiv(p.kblocks, p.kb_per_split);
}

// fprop / dgrad of small maps (14x14, 7x7 at batch 64) have fewer output tiles than the GPU has SMs.  Splitting the reduction
// (red.add epilogue into a zeroed output) keeps every SM busy: the output is a few MB, the reduction is where the work is.
void fill_machine_splits(GemmParams& p, long long units)
{
	const long long sms = pz_num_sms();
	p.splits = 1;
	p.kb_per_split = p.kblocks;
	if (units * 5 >= sms * 4 || units <= 0) return;   // >= 80% of one wave already
	// rounds of the persistent schedule x k-blocks per unit (+ a fixed per-unit cost: pipeline fill, red.add epilogue)
	const long long overhead = 6;
	long long best = 1, best_cost = (p.kblocks + overhead) * pz_cdiv(units, sms);
	for (long long sp = 2; sp <= 16 && sp * 6 <= p.kblocks; sp++) {
		const long long cost = pz_cdiv(units * sp, sms) * (pz_cdiv(p.kblocks, sp) + overhead);
		if (cost * 100 < best_cost * 92) { best = sp; best_cost = cost; }
	}
	if (best < 2) return;
	p.kb_per_split = (int)pz_cdiv(p.kblocks, best);
	p.splits = (int)pz_cdiv(p.kblocks, p.kb_per_split);
}

// algorithmic work of one conv pass: 2*MACs; every operand read once + the result written once
void set_alg(GemmParams& p, const Geo& g, int dtype)
{
	const double x = (double)g.N * g.C * g.H * g.W, y = (double)g.N * g.K * g.P * g.Q, w = (double)g.K * g.Cg * g.R * g.S;
	p.alg_flops = 2.0 * y * g.Cg * g.R * g.S;
	p.alg_bytes = (double)pz_dtype_size(dtype) * (x + y + w);
}


// Prepared filters (the TMA-fetched operand of fprop / dgrad): fp32 rows of `kpad` (multiple of 32) elements, K-major,
// rounded to tf32 (the tensor core would otherwise truncate) and zero-padded, in the library scratch.

// fprop: wp[ko_total][k] = w[ko_total][k], k = (c, r, s) < kdim
// chan != 0: k ordered (tap, channel block, 32 channels) for MnChanProducer, kpad = RS * ceil(Cg / 32) * 32
template <typename T> __device__ __forceinline__ T prep_round(T v) { return v; }
template <> __device__ __forceinline__ float prep_round<float>(float v) { return __uint_as_float(to_tf32(v)); }

template <typename T>
__global__ void prep_filter_fprop(const T* __restrict__ w, T* __restrict__ wp, int kdim, int kpad, long long total, int chan, int Cg,
								   int RS)
{
	// (prepared filters hold < 2^31 elements: 32-bit index arithmetic, a 64-bit division costs ~100 instructions per element)
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if ((long long)i >= total) return;
	const unsigned urow = i / (unsigned)kpad;
	const int k = (int)(i - urow * (unsigned)kpad);
	const long long row = urow;
	T v = T(0.0f);
	if (chan) {
		const int cpad = kpad / RS;
		const int t = k / cpad, c = k % cpad;
		if (c < Cg) v = w[row * kdim + (long long)c * RS + t];
	} else if (k < kdim) {
		v = w[row * kdim + k];
	}
	wp[i] = prep_round<T>(v);
}

// dgrad: wt[g*Cg + c][(ko, r', s')] = w[g*Kg + ko][c][r0 + sh*r'][s0 + sw*s'] -- the sub-filter of one output-parity class of
// a strided transposed convolution (r0 = s0 = 0, sh = sw = 1, Rc = R, Sc = S: the whole filter, stride-1 dgrad)
// chan != 0: k ordered (tap (r', s'), ko block, 32 ko) for MnChanProducer, kpad = Rc * Sc * ceil(Kg / 32) * 32
template <typename T>
__global__ void prep_filter_dgrad(const T* __restrict__ w, T* __restrict__ wt, int Kg, int Cg, int R, int S, int r0, int s0,
								   int sh, int sw, int Rc, int Sc, int kpad, long long total, int chan, int flip)
{
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if ((long long)i >= total) return;
	const int row = (int)(i / (unsigned)kpad);
	int k = (int)(i - (unsigned)row * (unsigned)kpad);
	int ko = -1, rc = 0, sc = 0;
	if (chan) {
		const int kopad = kpad / (Rc * Sc);
		const int t = k / kopad;
		ko = k % kopad;
		if (ko >= Kg) ko = -1;
		rc = t / Sc;
		sc = t % Sc;
		if (flip) { rc = Rc - 1 - rc; sc = Sc - 1 - sc; }     // halo dgrad: tap t' reads the filter mirrored
	} else if (k < Kg * Rc * Sc) {
		sc = k % Sc;
		k /= Sc;
		rc = k % Rc;
		ko = k / Rc;
	}
	T v = T(0.0f);
	if (ko >= 0) {
		const int c = row % Cg, g = row / Cg;
		v = prep_round<T>(w[((((long long)g * Kg + ko) * Cg + c) * R + (r0 + sh * rc)) * S + (s0 + sw * sc)]);
	}
	wt[i] = v;
}

// launch a templated prep kernel for the storage type (16-bit types are moved as raw bits)
#define PZ_PREP_LAUNCH(dtype, kernel, total, stream, w, wt, ...)                                                                      \
	do {                                                                                                                              \
		PZ_REQUIRE((long long)(total) < (1ll << 32), "conv2d: prepared filter exceeds 2^32 elements");                                \
		if ((dtype) == PZ_F32)                                                                                                        \
			kernel<float><<<(unsigned)pz_cdiv(total, 256), 256, 0, stream>>>((const float*)(w), (float*)(wt), __VA_ARGS__);           \
		else                                                                                                                          \
			kernel<__half><<<(unsigned)pz_cdiv(total, 256), 256, 0, stream>>>((const __half*)(w), (__half*)(wt), __VA_ARGS__);        \
		pz_count_launch(1);                                                                                                           \
	} while (0)

// col2im dgrad: wt[(c, r, s)][ko] = w[ko][c][r][s] (the filter transposed), rows of kpad elements
__global__ void prep_filter_col2im(const float* __restrict__ w, float* __restrict__ wt, int K, int CRS, int kpad, long long total)
{
	const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
	if ((long long)i >= total) return;
	const unsigned urow = i / (unsigned)kpad;
	const int ko = (int)(i - urow * (unsigned)kpad);
	const long long row = urow;
	wt[i] = ko < K ? __uint_as_float(to_tf32(w[(long long)ko * CRS + row])) : 0.0f;
}

// the (tap, channel) k order pads the channels of every tap to a multiple of 32: worth it unless the channel count is tiny
inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
// bke = elements per k-block: 32 (float) or 64 (half / bfloat16)
inline bool use_chan_order(int chans, int bke) { return round_up(chans, bke) * 3 <= chans * 4; }

// row length of the prepared dgrad filter: `taps` taps of Kg output channels
inline int dgrad_kpad(int Kg, int taps, bool may_chan, int bke)
{
	return may_chan && use_chan_order(Kg, bke) ? taps * round_up(Kg, bke) : round_up(Kg * taps, bke);
}


template <int THREADS, typename T>
__global__ void __launch_bounds__(THREADS) bias_grad_kernel(const T* __restrict__ t, float* __restrict__ db, long long N,
															 long long C, long long S, float alpha, int nsplit)
{
	// grid = (C, nsplit): each CTA reduces a slice of the images of one channel, then one red.add
	const long long c = blockIdx.x;
	float acc = 0.0f;
	for (long long n = blockIdx.y; n < N; n += nsplit) {
		const T* plane = t + (n * C + c) * S;
		for (long long s = threadIdx.x; s < S; s += THREADS) acc += (float)plane[s];
	}
	__shared__ float part[THREADS / 32];
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x < 32) {
		acc = threadIdx.x < THREADS / 32 ? part[threadIdx.x] : 0.0f;
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
		if (threadIdx.x == 0) atomicAdd(db + c, alpha * acc);
	}
}

int prescale(float* buf, long long n, float beta, void* stream)
{
	if (beta == 0.0f) return pz_memset8(buf, 0, (size_t)n * 4, stream);
	if (beta == 1.0f) return PZ_OK;
	return pz_scale_shift(PZ_F32, buf, buf, beta, 0.0f, n, stream);
}


// staged epilogue stores (Epilogue::staged): plain fp32 stores only; 16-byte lanes when the rows of a tile are contiguous in
// memory (row offset = image * ms0 + position) and everything is a multiple of four elements.  OFF by default: measured on the
// ResNet-50 layers (profiles/r02_conv_epilogue_gather_experiments.md) the two barriers per 32 columns cost more than the wider
// requests gain (64 -> 256 channels at 55 x 55: 0.068 -> 0.158 ms).  PZ_STAGED_EPI=1 enables it for experiments.
void set_staged(Epilogue& E, int dtype)
{
	static const bool on = [] { const char* e = getenv("PZ_STAGED_EPI"); return e && atoi(e); }();
	E.staged = 0;
	if (!on || dtype != PZ_F32 || E.atomic || E.beta != 0.0f || E.bias_mode == 2 || E.c2i || E.out_kind != OUT_F32) return;
	const bool contiguous = E.ms2 == 1 && (E.md2.d == 0 || (long long)E.ms1 == (long long)E.md2.d);
	const bool vec = contiguous && E.md12.d % 4 == 0 && E.ms0 % 4 == 0 && E.ncs % 4 == 0 && E.group_stride % 4 == 0 && ((uintptr_t)E.out & 15) == 0;
	E.staged = vec ? 2 : 1;
}

// 16-byte producer lanes (MODE_MN_VEC / MODE_K_POS_VEC in pz_umma.cuh) need float tensors whose planes, image and group strides are
// multiples of four elements on a 16-byte aligned base.  PZ_VEC_GATHER: 0 = never, 1 = where it measured faster (default: the dy
// operand of wgrad next to a tap-gathered x, and 1x1 wgrad over small planes), 2 = wherever it is legal (experiments).  Four times
// fewer load instructions do NOT make the fprop / dgrad gather faster (28 x 28, 128 -> 512 channels: 0.043 -> 0.048 ms): the
// gather is bound behind the LSU, not by issue (profiles/r02_conv_epilogue_gather_experiments.md).
// PZ_TMA_WGRAD: 0 = producers gather every wgrad operand, 1 = both operands of a 1x1 / stride-1 wgrad over aligned planes come
// through the copy engine (MODE_K_POS_TMA: 3-d tensor maps, no producer instructions), 2 (default) = also the dy operand next to a
// tap-gathered x.  Measured (N = 64): 1x1 wgrad over 28 x 28 planes 0.052 -> 0.028 ms (4.6 TB/s), over 14 x 14 planes 0.040 -> 0.031 ms,
// 3x3 wgrad 0.071 -> 0.060 / 0.080 -> 0.065 ms; ResNet-50 step 16.64 -> 15.99 ms.  The copy engine hands over raw float bits and the
// tf32 MMA ignores the low 13 mantissa bits: left alone, a copied operand is truncated, which shrinks the result by 3.5e-4 per
// copied operand on average (max error against a float64 contraction 8.5e-4 with two copied operands -- one teacher-forced
// ResNet-50 layer crossed 1e-3).  The producer group therefore rounds the landed tile in place (umma_gemm_kernel, FIXUP) to the
// value cvt.rna gives the gathering producers: 3.3e-4 again (tools/check_tma_wgrad.py).
int tma_wgrad_level()
{
	static const int level = [] { const char* e = getenv("PZ_TMA_WGRAD"); return e ? atoi(e) : 2; }();
	return level;
}
// planes of `chans` channels per image, dense (N, chans, plane): tile rows must not run past the channels
bool kpos_tma_ok(const Operand& op, const void* ptr, int chans, int rows, int groups, int dtype)
{
	const int es = dtype == PZ_F32 ? 4 : 2;
	return groups == 1 && (op.plane * es) % 16 == 0 && op.plane >= elems_per_kblock(dtype) && chans % rows == 0 && ((uintptr_t)ptr & 15) == 0 &&
		   op.rs0 == op.plane && op.ks0 == (long long)chans * op.plane;
}

// What the copy engine cannot do for these tensors (tools/ubench/tma4d.cu, profiles/r02_tma_operands.txt): a box must start on a
// 16-byte boundary of global memory, so the tap-shifted windows of a 3x3 filter (one element left / right in a row) cannot be
// fetched as boxes of an NCHW tensor -- a coordinate of -1 or +1 along the row raises "illegal instruction"; shifts by whole rows
// would be legal.  That is why only 1x1 filters (and the un-shifted dy operand of any wgrad) take the tensor-map paths.
// PZ_TMA_FPROP: the activation operand of 1x1 / stride-1 fprop and dgrad over dense, 16-byte aligned planes comes through the copy
// engine in its memory order and is multiplied through an MN-major descriptor (MODE_MN_TMA).  0 = off, 1 (default) = planes of at
// least 512 positions, 2 = every legal plane.  The landed tile is rounded to tf32 in place like the wgrad operands (see
// PZ_TMA_WGRAD).  Rows come in chunks of 32 positions of one image, so a 14 x 14 plane fills 196 of 224 rows and loses what the
// copy engine gains: ResNet-50 step 15.91 (off) / 15.80 (1) / 15.84 ms (2), family of the 1x1 convolutions 5.12 / 5.01 / 5.05 ms
// (profiles/r02_tma_operands.txt).
int tma_fprop_level()
{
	static const int level = [] { const char* e = getenv("PZ_TMA_FPROP"); return e ? atoi(e) : 1; }();
	return level;
}
bool mn_tma_ok(const Operand& A, const Epilogue& E, int groups, int dtype)
{
	const int lvl = tma_fprop_level();
	const long long plane = A.rd12.d;
	if (lvl < 1 || (lvl < 2 && plane < 512)) return false;
	const bool dense_out = E.ms2 == 1 && (E.md2.d == 0 || (long long)E.ms1 == (long long)E.md2.d) && (long long)E.md12.d == plane && !E.c2i &&
						   E.bias_mode != 2 && E.out_kind == OUT_F32;
	return dtype == PZ_F32 && groups == 1 && dense_out && A.R == 1 && A.S == 1 && A.ah == 1 && A.aw == 1 && A.ch == 0 && A.cw == 0 && A.cdh == 1 &&
		   A.cdw == 1 && (long long)A.H * A.W == plane && A.Wd == A.W && (int)A.rd2.d == A.W && plane % 4 == 0 && plane >= 32 &&
		   A.chans % 32 == 0 && A.ks0 == plane && (long long)A.rs0 == (long long)A.chans * plane && ((uintptr_t)A.ptr & 15) == 0;
}

int vec_gather_level()
{
	static const int level = [] { const char* e = getenv("PZ_VEC_GATHER"); return e ? atoi(e) : 1; }();
	return level;
}
// the MnChan operand of a 1 x 1 / stride-1 / un-padded filter: position offset = row index inside the plane
bool mn_vec_ok(const Operand& A, int dtype)
{
	return vec_gather_level() >= 2 && dtype == PZ_F32 && A.R == 1 && A.S == 1 && A.ah == 1 && A.aw == 1 && A.ch == 0 && A.cw == 0 && A.cdh == 1 &&
		   A.cdw == 1 && (long long)A.H * A.W == (long long)A.rd12.d && A.Wd == A.W && (int)A.rd2.d == A.W && A.rd12.d % 4 == 0 && A.rs0 % 4 == 0 &&
		   A.ks0 % 4 == 0 && A.group_stride % 4 == 0 && ((uintptr_t)A.ptr & 15) == 0;
}
// the KPosDense operand (wgrad): rows = planes, k = positions of one image
bool kpos_vec_ok(const Operand& B, int dtype)
{
	return vec_gather_level() >= 1 && dtype == PZ_F32 && B.plane % 4 == 0 && B.rs0 % 4 == 0 && B.ks0 % 4 == 0 && B.group_stride % 4 == 0 &&
		   ((uintptr_t)B.ptr & 15) == 0;
}
}  // namespace

extern "C" {

// float32 math mode of the contractions (convolutions here, pz_gemm through pz_exact_fp32()): 0 = TF32 tensor-core products
// (default), 1 = exact fp32 FMAs on the CUDA cores.  Process-wide, like the reference's per-context enableTensorOps.
int pz_set_exact_fp32(int on)
{
	g_exact_fp32 = on ? 1 : 0;
	return PZ_OK;
}
int pz_exact_fp32(void) { return g_exact_fp32; }

int pz_conv2d_fprop(int dtype, const pz_conv2d_desc* d, const void* x, const void* w, const void* bias, void* y, void* stream)
{
	Geo g;
	int st = check_desc(dtype, d, g);
	if (st != PZ_OK) return st;
	if (dtype == PZ_F32 && g_exact_fp32) {
		const long long total = (long long)g.N * g.K * g.P * g.Q;
		exact_fprop_kernel<<<exact_grid(total), 256, 0, pz_stream(stream)>>>((const float*)x, (const float*)w, (const float*)bias, (float*)y, g, total);
		pz_count_launch(1);
		PZ_LAUNCH_CHECK();
		return PZ_OK;
	}
	const int RS = g.R * g.S, PQ = g.P * g.Q, HW = g.H * g.W;

	GemmParams p{};
	Operand& A = p.A;   // im2col view of x: rows (n,p,q), k (c,r,s)
	const int bke = elems_per_kblock(dtype);
	const bool h16 = dtype != PZ_F32;
	A.ptr = x;
	A.rd12 = make_fastdiv(PQ); A.rd2 = make_fastdiv(g.Q);
	A.kd12 = make_fastdiv(RS); A.kd2 = make_fastdiv(g.S);
	A.rs0 = g.C * HW; A.ks0 = HW;
	A.ah = g.sh; A.bh = g.dh; A.ch = -g.ph;
	A.aw = g.sw; A.bw = g.dw; A.cw = -g.pw;
	A.H = g.H; A.W = g.W; A.Wd = g.W;
	A.cdh = A.cdw = 1;
	A.rows = g.N * PQ; A.kdim = g.Cg * RS;
	A.R = g.R; A.S = g.S;
	A.group_stride = (long long)g.Cg * HW;

	// filter: prepared copy [K][kpad], fetched by TMA
	const bool fast = tap_entries_fit(34ll * HW, g);
	// 16-bit tensors have no table-driven tap producer: even a 3-channel first layer takes the channel-ordered path (its padded
	// channels cost predicated-off loads and cheap MMA work, far less than the general gather)
	const bool chan = fast && (use_chan_order(g.Cg, bke) || h16);
	const int kdim = g.Cg * RS, kpad = chan ? RS * round_up(g.Cg, bke) : round_up(kdim, bke);
	const long long wtotal = (long long)g.K * kpad;
	void* wp = pz_scratch((size_t)wtotal * pz_dtype_size(dtype));
	if (!wp) { pz_set_error(PZ_ERR_MEMORY, "conv2d fprop: cannot allocate %lld bytes of filter scratch", wtotal * 4); return PZ_ERR_MEMORY; }
	PZ_PREP_LAUNCH(dtype, prep_filter_fprop, wtotal, pz_stream(stream), w, wp, kdim, kpad, wtotal, chan ? 1 : 0, g.Cg, RS);
	PZ_LAUNCH_CHECK();
	const TmaSource tsrc{wp, g.K, kpad};
	p.tma_rows_per_group = g.Kg;

	Epilogue& E = p.E;
	E.out = y;
	E.bias = bias;
	E.out_kind = out_kind_of(dtype);
	E.md12 = make_fastdiv(PQ); E.md2 = make_fastdiv(0);
	E.ms0 = g.K * PQ; E.ms1 = 0; E.ms2 = 1;
	E.ncs = PQ;
	E.M = g.N * PQ; E.N = g.Kg;
	E.alpha = 1.0f; E.beta = 0.0f;
	E.bias_mode = bias ? 1 : 0;
	E.atomic = 0;
	E.group_stride = (long long)g.Kg * PQ;
	E.bias_group_stride = g.Kg;

	A.chans = g.Cg;
	A.kbdiv = make_fastdiv((uint32_t)(round_up(g.Cg, bke) / bke));
	p.kblocks = kpad / bke;
	set_alg(p, g, dtype);
	if (chan && RS > 1 && g.sh == 1 && g.sw == 1 && g.dh == 1 && g.dw == 1) {
		// stride-1 filters with more than one tap: halo path (x staged once per channel block, taps = shifted windows)
		HaloGeometry hg{};
		hg.src = x;
		hg.N = g.N; hg.Hs = g.H; hg.Ws = g.W; hg.ph = g.ph; hg.pw = g.pw;
		hg.R = g.R; hg.S = g.S;
		hg.chans = g.Cg; hg.out_chans = g.Kg;
		hg.out_h = g.P; hg.out_w = g.Q;
		hg.groups = g.G;
		hg.img_stride = (long long)g.C * HW; hg.chan_stride = HW; hg.group_stride = (long long)g.Cg * HW;
		Epilogue HE = E;
		HE.ms0 = g.K * PQ; HE.ms1 = g.Q; HE.ms2 = 1;
		HE.M = 0;
		const int hst = launch_halo(hg, dtype, tsrc, HE, p.alg_flops, p.alg_bytes, pz_stream(stream));
		if (hst != PZ_ERR_UNSUPPORTED) return hst;
	}
	const int bn = pick_bn(g.Kg, (long long)g.N * PQ, p.kblocks, g.G, 256);
	fill_machine_splits(p, pz_cdiv(E.M, BM) * pz_cdiv(E.N, bn) * g.G);
	if (h16) { p.splits = 1; p.kb_per_split = p.kblocks; }       // red.add needs an fp32 output
	if (p.splits > 1) {
		E.atomic = 1;
		int st2 = pz_memset8(y, 0, (size_t)g.N * g.K * PQ * 4, stream);
		if (st2 != PZ_OK) return st2;
	}
	// the table-driven tap producer exists for float only; 16-bit tensors with very few channels take the general gather
	int amode = chan ? MODE_MN_CHAN : (fast && !h16 ? MODE_MN_TAP : MODE_MN_GENERAL);
	if (amode == MODE_MN_CHAN && mn_tma_ok(A, E, g.G, dtype)) {
		const PlaneTma px{x, PQ, g.C, g.N};
		return launch(p, dtype, bn, MODE_MN_TMA, MODE_TMA, false, g.G, &tsrc, pz_stream(stream), &px);
	}
	if (amode == MODE_MN_CHAN && mn_vec_ok(A, dtype)) amode = MODE_MN_VEC;
	set_staged(E, dtype);
	return launch(p, dtype, bn, amode, MODE_TMA, amode == MODE_MN_GENERAL ? false : RS > 31, g.G, &tsrc, pz_stream(stream));
}

// the dgrad filter repack lives in the library's own scratch; callers need not supply a workspace any more
size_t pz_conv2d_dgrad_workspace(int dtype, const pz_conv2d_desc* d)
{
	(void)dtype;
	(void)d;
	return 0;
}

int pz_conv2d_dgrad(int dtype, const pz_conv2d_desc* d, const void* dy, const void* w, const void* bias, void* dx,
					void* workspace, size_t workspace_bytes, void* stream)
{
	Geo g;
	int st = check_desc(dtype, d, g);
	if (st != PZ_OK) return st;
	if (dtype == PZ_F32 && g_exact_fp32) {
		const long long total = (long long)g.N * g.C * g.H * g.W;
		exact_dgrad_kernel<<<exact_grid(total), 256, 0, pz_stream(stream)>>>((const float*)dy, (const float*)w, (const float*)bias, (float*)dx, g, total);
		pz_count_launch(1);
		PZ_LAUNCH_CHECK();
		return PZ_OK;
	}
	const int RS = g.R * g.S, PQ = g.P * g.Q, HW = g.H * g.W;
	const bool is1x1 = g.R == 1 && g.S == 1;
	const bool strided = g.sh > 1 || g.sw > 1;

	const int bke = elems_per_kblock(dtype);
	const bool h16 = dtype != PZ_F32;
	const size_t es = pz_dtype_size(dtype);
	Epilogue E{};
	E.out = dx;
	E.bias = bias;
	E.out_kind = out_kind_of(dtype);
	E.alpha = 1.0f; E.beta = 0.0f;
	E.bias_mode = bias ? 1 : 0;
	E.atomic = 0;
	E.ncs = HW;
	E.N = g.Cg;
	E.group_stride = (long long)g.Cg * HW;
	E.bias_group_stride = g.Cg;

	(void)workspace;
	(void)workspace_bytes;

	// launches one stride-1 transposed-convolution problem over the sub-filter (r0 + sh*r', s0 + sw*s') whose outputs are the
	// input-gradient positions (a_h + sh*h', a_w + sw*w'); the whole tensor for stride 1 (a = 0, sh = sw = 1).
	// mode 0: parity class of a strided, un-dilated filter (fast tap producer); mode 1: stride 1, any dilation (fast tap
	// producer); mode 2: any stride and dilation through the exact-division gather (csh = csw = 1, whole filter).
	auto run_class = [&](char* wt, int a_h, int a_w, int r0, int s0, int csh, int csw, int Rc, int Sc, int Hc, int Wc, int mode) -> int {
		if (mode != 2 && h16 && !use_chan_order(g.Kg, bke)) {
			// no table-driven tap producer for 16-bit tensors: few output channels take the general gather, which handles the
			// whole (possibly strided) problem in one launch -- only legal when called for the whole tensor
			PZ_REQUIRE(a_h == 0 && a_w == 0 && csh == 1 && csw == 1, "conv2d dgrad: 16-bit strided filters need >= 48 output channels");
			mode = 2;
		}
		const bool chan = mode != 2 && use_chan_order(g.Kg, bke);
		const int kdim = g.Kg * Rc * Sc, kpad = dgrad_kpad(g.Kg, Rc * Sc, mode != 2, bke);
		const long long total = (long long)g.C * kpad;
		// stride-1 filters with more than one tap: halo path (dy staged once per channel block, taps = shifted windows)
		const bool halo = mode == 1 && chan && g.dh == 1 && g.dw == 1 && Rc * Sc > 1 && g.R - 1 - g.ph >= 0 && g.S - 1 - g.pw >= 0;
		PZ_PREP_LAUNCH(dtype, prep_filter_dgrad, total, pz_stream(stream), w, wt, g.Kg, g.Cg, g.R, g.S, r0, s0, csh, csw, Rc, Sc, kpad, total,
					   chan ? 1 : 0, halo ? 1 : 0);
		PZ_LAUNCH_CHECK();
		if (halo) {
			HaloGeometry hg{};
			hg.src = dy;
			hg.N = g.N; hg.Hs = g.P; hg.Ws = g.Q; hg.ph = g.R - 1 - g.ph; hg.pw = g.S - 1 - g.pw;
			hg.R = g.R; hg.S = g.S;
			hg.chans = g.Kg; hg.out_chans = g.Cg;
			hg.out_h = g.H; hg.out_w = g.W;
			hg.groups = g.G;
			hg.img_stride = (long long)g.K * PQ; hg.chan_stride = PQ; hg.group_stride = (long long)g.Kg * PQ;
			Epilogue HE = E;
			HE.ms0 = g.C * HW; HE.ms1 = g.W; HE.ms2 = 1;
			HE.M = 0; HE.N = g.Cg;
			const TmaSource htsrc{wt, g.C, kpad};
			const double flops = 2.0 * (double)g.N * HW * g.K * g.Cg * Rc * Sc;
			const double bytes = (double)es * ((double)g.N * g.K * PQ + (double)g.K * g.Cg * Rc * Sc + (double)g.N * g.C * HW);
			const int hst = launch_halo(hg, dtype, htsrc, HE, flops, bytes, pz_stream(stream));
			if (hst != PZ_ERR_UNSUPPORTED) return hst;
			// geometry does not fit the halo kernel: re-prepare the filter un-mirrored and take the gather path
			PZ_PREP_LAUNCH(dtype, prep_filter_dgrad, total, pz_stream(stream), w, wt, g.Kg, g.Cg, g.R, g.S, r0, s0, csh, csw, Rc, Sc, kpad,
						   total, chan ? 1 : 0, 0);
			PZ_LAUNCH_CHECK();
		}

		GemmParams q{};
		Operand& QA = q.A;                 // rows (n, h', w') of the class, k (ko, r', s')
		QA.ptr = dy;
		QA.rd12 = make_fastdiv(Hc * Wc); QA.rd2 = make_fastdiv(Wc);
		QA.kd12 = make_fastdiv(Rc * Sc); QA.kd2 = make_fastdiv(Sc);
		QA.rs0 = g.K * PQ; QA.ks0 = PQ;
		QA.H = g.P; QA.W = g.Q; QA.Wd = g.Q;
		if (mode == 0) {
			QA.ah = 1; QA.bh = -1; QA.ch = (a_h + g.ph - r0) / csh;
			QA.aw = 1; QA.bw = -1; QA.cw = (a_w + g.pw - s0) / csw;
			QA.cdh = QA.cdw = 1;
		} else {
			// hh = (h + pad - r*dil) / stride when divisible (mode 1: stride 1)
			QA.ah = 1; QA.bh = -g.dh; QA.ch = g.ph;
			QA.aw = 1; QA.bw = -g.dw; QA.cw = g.pw;
			QA.cdh = g.sh; QA.cdw = g.sw;
		}
		QA.rows = g.N * Hc * Wc; QA.kdim = kdim;
		QA.R = Rc; QA.S = Sc;
		QA.chans = g.Kg;
		QA.kbdiv = make_fastdiv((uint32_t)(round_up(g.Kg, bke) / bke));
		QA.group_stride = (long long)g.Kg * PQ;

		q.E = E;
		q.E.out = (char*)dx + ((long long)a_h * g.W + a_w) * es;
		q.E.md12 = make_fastdiv(Hc * Wc); q.E.md2 = make_fastdiv(Wc);
		q.E.ms0 = g.C * HW; q.E.ms1 = csh * g.W; q.E.ms2 = csw;
		q.E.M = g.N * Hc * Wc;
		q.kblocks = kpad / bke;
		q.splits = 1;
		q.kb_per_split = q.kblocks;
		q.tma_rows_per_group = g.Cg;
		q.alg_flops = 2.0 * (double)q.E.M * g.K * g.Cg * Rc * Sc / g.G;
		q.alg_bytes = (double)es * ((double)g.N * g.K * PQ / (csh * csw) + (double)g.K * g.Cg * Rc * Sc + (double)q.E.M * g.C);
		const TmaSource tsrc{wt, g.C, kpad};
		const int bn = pick_bn(g.Cg, q.E.M, q.kblocks, g.G, 256);
		if (!h16 && a_h == 0 && a_w == 0 && csh == 1 && csw == 1) {
			// whole-tensor problem: split the reduction when there are too few tiles to fill the machine
			fill_machine_splits(q, pz_cdiv(q.E.M, BM) * pz_cdiv(q.E.N, bn) * g.G);
			if (q.splits > 1) {
				q.E.atomic = 1;
				int st2 = pz_memset8(dx, 0, (size_t)g.N * g.C * HW * es, stream);
				if (st2 != PZ_OK) return st2;
			}
		}
		const bool cdiv = mode == 2 ? strided : Rc * Sc > 31;
		int amode = mode == 2 ? MODE_MN_GENERAL : (chan ? MODE_MN_CHAN : MODE_MN_TAP);
		if (amode == MODE_MN_CHAN && a_h == 0 && a_w == 0 && csh == 1 && csw == 1 && mn_tma_ok(QA, q.E, g.G, dtype)) {
			const PlaneTma pdy{dy, PQ, g.K, g.N};
			return launch(q, dtype, bn, MODE_MN_TMA, MODE_TMA, false, g.G, &tsrc, pz_stream(stream), &pdy);
		}
		if (amode == MODE_MN_CHAN && mn_vec_ok(QA, dtype)) amode = MODE_MN_VEC;
		set_staged(q.E, dtype);
		return launch(q, dtype, bn, amode, MODE_TMA, cdiv, g.G, &tsrc, pz_stream(stream));
	};

	if (!h16 && g.G == 1 && g.C <= 4 && g.C * RS <= 256 && RS > 1 && bias == nullptr && use_chan_order(g.K, bke) && tap_entries_fit(34ll * PQ, g)) {
		// Very few input channels (the first layer of a network): the implicit GEMM over (n, h, w) x c would gather every dy
		// element R*S times for a handful of output columns.  Instead  D[(n,p,q)][(c,r,s)] = sum_k dy[n,k,p,q] * w[k,c,r,s]
		// reads dy exactly once (a 1x1-convolution-shaped GEMM) and the epilogue scatters D into dx (col2im) with red.add.
		st = pz_memset8(dx, 0, (size_t)g.N * g.C * HW * 4, stream);
		if (st != PZ_OK) return st;
		const int crs = g.C * RS, kpad = round_up(g.K, 32);
		const long long total = (long long)crs * kpad;
		float* wt = scratch((size_t)total * sizeof(float));
		if (!wt) { pz_set_error(PZ_ERR_MEMORY, "conv2d dgrad: cannot allocate filter scratch"); return PZ_ERR_MEMORY; }
		prep_filter_col2im<<<(unsigned)pz_cdiv(total, 256), 256, 0, pz_stream(stream)>>>((const float*)w, wt, g.K, crs, kpad, total);
		pz_count_launch(1);
		PZ_LAUNCH_CHECK();

		GemmParams q{};
		Operand& QA = q.A;                 // rows (n, p, q) of dy, k = ko: a 1x1 "convolution" over dy
		QA.ptr = (const float*)dy;
		QA.rd12 = make_fastdiv(PQ); QA.rd2 = make_fastdiv(g.Q);
		QA.kd12 = make_fastdiv(1); QA.kd2 = make_fastdiv(1);
		QA.rs0 = g.K * PQ; QA.ks0 = PQ;
		QA.ah = 1; QA.bh = 1; QA.ch = 0;
		QA.aw = 1; QA.bw = 1; QA.cw = 0;
		QA.H = g.P; QA.W = g.Q; QA.Wd = g.Q;
		QA.cdh = QA.cdw = 1;
		QA.rows = g.N * PQ; QA.kdim = g.K;
		QA.R = 1; QA.S = 1;
		QA.chans = g.K;
		QA.kbdiv = make_fastdiv((uint32_t)(kpad / 32));
		QA.group_stride = 0;

		q.E = E;
		q.E.out = (float*)dx;
		q.E.md12 = make_fastdiv(PQ); q.E.md2 = make_fastdiv(g.Q);
		q.E.ms0 = g.C * HW; q.E.ms1 = 0; q.E.ms2 = 0;
		q.E.ncs = 0;
		q.E.M = g.N * PQ; q.E.N = crs;
		q.E.c2i = 1;
		q.E.c2i_sh = g.sh; q.E.c2i_sw = g.sw; q.E.c2i_ph = g.ph; q.E.c2i_pw = g.pw; q.E.c2i_dh = g.dh; q.E.c2i_dw = g.dw;
		q.E.c2i_H = g.H; q.E.c2i_W = g.W;
		q.E.c2i_rs = make_fastdiv((uint32_t)RS); q.E.c2i_s = make_fastdiv((uint32_t)g.S);
		q.splits = 1;
		q.kblocks = kpad / BK;
		q.kb_per_split = q.kblocks;
		q.tma_rows_per_group = crs;
		q.alg_flops = 2.0 * (double)g.N * PQ * g.K * crs;
		q.alg_bytes = 4.0 * ((double)g.N * g.K * PQ + (double)g.K * crs + (double)g.N * g.C * HW);
		const TmaSource tsrc{wt, crs, kpad};
		const int bn = crs <= 64 ? 64 : (crs <= 128 ? 128 : 256);
		return launch(q, dtype, bn, mn_vec_ok(QA, dtype) ? MODE_MN_VEC : MODE_MN_CHAN, MODE_TMA, false, 1, &tsrc, pz_stream(stream));
	}

	if (is1x1 && g.ph == 0 && g.pw == 0 && strided && bias == nullptr && (!h16 || use_chan_order(g.Kg, bke))) {
		// 1x1 strided: dx[n,c,p*sh,q*sw] = sum_k dy[n,k,p,q] * w[k,c]; all other positions of dx are zero.  Small 16-bit
		// filters and the deconvolution-forward bias take the general gather path at the end of this function instead.
		st = pz_memset8(dx, 0, (size_t)g.N * g.C * HW * es, stream);
		if (st != PZ_OK) return st;
		char* wt = (char*)pz_scratch((size_t)g.C * dgrad_kpad(g.Kg, 1, true, bke) * es);
		if (!wt) { pz_set_error(PZ_ERR_MEMORY, "conv2d dgrad: cannot allocate filter scratch"); return PZ_ERR_MEMORY; }
		PZ_REQUIRE(tap_entries_fit(34ll * PQ, g), "conv2d dgrad: tensor too large");
		return run_class(wt, 0, 0, 0, 0, g.sh, g.sw, 1, 1, g.P, g.Q, 0);
	}

	if (strided && g.dh == 1 && g.dw == 1 && tap_entries_fit(34ll * PQ, g) && (!h16 || use_chan_order(g.Kg, bke))) {
		// A strided transposed convolution splits into sh*sw independent STRIDE-1 problems, one per parity class
		// (h mod sh, w mod sw) of the input-gradient positions: only the taps r = r0 + sh*r' with r0 = (a_h + pad_h) mod sh
		// can reach such a position.  No tap is ever evaluated on a zero, and each class runs on the fast tap producer.
		size_t need = 0;
		for (int a_h = 0; a_h < g.sh && a_h < g.H; a_h++)
			for (int a_w = 0; a_w < g.sw && a_w < g.W; a_w++) {
				const int r0 = (a_h + g.ph) % g.sh, s0 = (a_w + g.pw) % g.sw;
				const int Rc = r0 < g.R ? (g.R - r0 + g.sh - 1) / g.sh : 0, Sc = s0 < g.S ? (g.S - s0 + g.sw - 1) / g.sw : 0;
				need += (size_t)g.C * dgrad_kpad(g.Kg, Rc * Sc, true, bke) * es;
			}
		char* wsp = (char*)pz_scratch(need);
		if (!wsp) { pz_set_error(PZ_ERR_MEMORY, "conv2d dgrad: cannot allocate %zu bytes of filter scratch", need); return PZ_ERR_MEMORY; }
		bool zeroed = false;
		for (int a_h = 0; a_h < g.sh && a_h < g.H; a_h++)
			for (int a_w = 0; a_w < g.sw && a_w < g.W; a_w++) {
				const int r0 = (a_h + g.ph) % g.sh, s0 = (a_w + g.pw) % g.sw;
				const int Rc = r0 < g.R ? (g.R - r0 + g.sh - 1) / g.sh : 0, Sc = s0 < g.S ? (g.S - s0 + g.sw - 1) / g.sw : 0;
				const int Hc = (g.H - a_h + g.sh - 1) / g.sh, Wc = (g.W - a_w + g.sw - 1) / g.sw;
				if (Rc == 0 || Sc == 0) {
					// no tap reaches this class (stride larger than the filter): those gradients are zero (+ bias)
					if (!zeroed) {
						PZ_REQUIRE(bias == nullptr, "conv2d dgrad: bias with stride > filter size is not supported");
						st = pz_memset8(dx, 0, (size_t)g.N * g.C * HW * es, stream);
						if (st != PZ_OK) return st;
						zeroed = true;      // the memset precedes every class launch in stream order: classes only overwrite their own positions
					}
					continue;
				}
				st = run_class(wsp, a_h, a_w, r0, s0, g.sh, g.sw, Rc, Sc, Hc, Wc, 0);
				if (st != PZ_OK) return st;
				wsp += (size_t)g.C * dgrad_kpad(g.Kg, Rc * Sc, true, bke) * es;
			}
		return PZ_OK;
	}

	{
		const bool tapmode = !strided && tap_entries_fit(34ll * PQ, g);
		char* wt = (char*)pz_scratch((size_t)g.C * round_up(g.Kg, bke) * RS * es);
		if (!wt) { pz_set_error(PZ_ERR_MEMORY, "conv2d dgrad: cannot allocate filter scratch"); return PZ_ERR_MEMORY; }
		return run_class(wt, 0, 0, 0, 0, 1, 1, g.R, g.S, g.H, g.W, tapmode ? 1 : 2);
	}
}

int pz_conv2d_wgrad(int dtype, const pz_conv2d_desc* d, const void* x, const void* dy, void* dw, float alpha, float beta,
					void* stream)
{
	Geo g;
	int st = check_desc(dtype, d, g);
	if (st != PZ_OK) return st;
	if (dtype == PZ_F32 && g_exact_fp32) {
		const long long elems = (long long)g.K * g.Cg * g.R * g.S;
		PZ_REQUIRE(elems < (1ll << 31), "conv2d wgrad: filter too large");
		exact_wgrad_kernel<<<(unsigned)elems, 256, 0, pz_stream(stream)>>>((const float*)x, (const float*)dy, (float*)dw, g, alpha, beta);
		pz_count_launch(1);
		PZ_LAUNCH_CHECK();
		return PZ_OK;
	}
	const int RS = g.R * g.S, PQ = g.P * g.Q, HW = g.H * g.W;
	PZ_REQUIRE((long long)g.N * PQ < (1ll << 31), "conv2d wgrad: reduction too long");

	// reduction index k ordered (image, block of positions): k-block kb = n * kbpi + pb (positions past PQ are zero)
	const int bke = elems_per_kblock(dtype);
	const bool h16 = dtype != PZ_F32;
	const int kbpi = (int)pz_cdiv(PQ, bke);
	const bool dense_x = RS == 1 && g.sh == 1 && g.sw == 1 && g.ph == 0 && g.pw == 0;
	const bool fast = tap_entries_fit((long long)g.Cg * HW, g);

	GemmParams p{};
	Operand& A = p.A;   // im2col view of x: rows (c,r,s), k (n,p,q)
	A.ptr = x;
	A.rd12 = make_fastdiv(RS); A.rd2 = make_fastdiv(g.S);
	A.kd12 = make_fastdiv(PQ); A.kd2 = make_fastdiv(g.Q);
	A.rs0 = HW; A.ks0 = g.C * HW;
	A.ah = g.dh; A.bh = g.sh; A.ch = -g.ph;
	A.aw = g.dw; A.bw = g.sw; A.cw = -g.pw;
	A.H = g.H; A.W = g.W; A.Wd = g.W;
	A.cdh = A.cdw = 1;
	A.rows = g.Cg * RS; A.kdim = g.N * PQ;
	A.R = g.R; A.S = g.S;
	A.group_stride = (long long)g.Cg * HW;
	A.kbdiv = make_fastdiv((uint32_t)kbpi);
	A.plane = PQ;

	Operand& B = p.B;   // dy: rows ko, k (n, pq)
	B.ptr = dy;
	B.rd12 = make_fastdiv(1); B.rd2 = make_fastdiv(0);
	B.kd12 = make_fastdiv(PQ); B.kd2 = make_fastdiv(0);
	B.rs0 = PQ; B.ks0 = g.K * PQ;
	B.ah = B.bh = B.ch = 0;
	B.aw = 0; B.bw = 1; B.cw = 0;
	B.H = 1; B.W = PQ; B.Wd = 0;
	B.cdh = B.cdw = 1;
	B.R = B.S = 1;
	B.rows = g.Kg; B.kdim = g.N * PQ;
	B.group_stride = (long long)g.Kg * PQ;
	B.kbdiv = make_fastdiv((uint32_t)kbpi);
	B.plane = PQ;

	Epilogue& E = p.E;
	E.out = dw;
	E.bias = nullptr;
	E.out_kind = out_kind_of(dtype);
	E.md12 = make_fastdiv(0); E.md2 = make_fastdiv(0);
	E.ms0 = 0; E.ms1 = 0; E.ms2 = 1;
	E.ncs = g.Cg * RS;
	E.M = g.Cg * RS; E.N = g.Kg;
	E.alpha = alpha; E.beta = beta;
	E.bias_mode = 0;
	E.group_stride = (long long)g.Kg * g.Cg * RS;
	E.bias_group_stride = 0;

	PZ_REQUIRE(fast || !h16, "conv2d wgrad: tensor too large for the 16-bit path");
	p.kblocks = fast ? g.N * kbpi : (int)pz_cdiv(A.kdim, bke);
	const int bn = g.Kg > 64 ? 128 : 64;        // the x operand is re-read once per column tile: keep those few
	set_splits(p, pz_cdiv(E.M, BM) * pz_cdiv(E.N, bn) * g.G, 8);
	E.atomic = p.splits > 1;
	set_alg(p, g, dtype);
	const long long wcount = (long long)g.K * g.Cg * RS;
	float* acc = nullptr;
	if (E.atomic && h16) {
		// split-K with 16-bit storage: red.add into an fp32 scratch, then dw = beta*dw + acc
		acc = (float*)pz_scratch((size_t)wcount * sizeof(float));
		if (!acc) { pz_set_error(PZ_ERR_MEMORY, "conv2d wgrad: cannot allocate the split-K accumulator"); return PZ_ERR_MEMORY; }
		st = pz_memset8(acc, 0, (size_t)wcount * 4, stream);
		if (st != PZ_OK) return st;
		E.out = acc;
		E.out_kind = OUT_F32;
		E.beta = 0.0f;
	} else if (E.atomic) {
		st = prescale((float*)dw, wcount, beta, stream);
		if (st != PZ_OK) return st;
	}
	if (!fast) st = launch(p, dtype, bn, MODE_K_GENERAL, MODE_K_DENSE, false, g.G, nullptr, pz_stream(stream));
	else {
		const bool vec_dy = kpos_vec_ok(B, dtype);
		const bool vec_x = dense_x && kpos_vec_ok(A, dtype) && (PQ < 512 || vec_gather_level() >= 2);    // 28 x 28 planes: measured slower
		const int amode = dense_x ? (vec_x && vec_dy ? MODE_K_POS_VEC : MODE_K_POS_DENSE) : MODE_K_POS_TAP;
		const int bmode = vec_dy && (!dense_x || vec_x) ? MODE_K_POS_VEC : MODE_K_POS_DENSE;
		// operands over 16-byte aligned planes can come through the copy engine instead (3-d tensor maps, no producer instructions)
		const int tl = tma_wgrad_level();
		const bool tma_dy = tl >= 1 && kpos_tma_ok(B, dy, g.K, bn, g.G, dtype);
		const bool tma_x = tma_dy && dense_x && kpos_tma_ok(A, x, g.C, BM, g.G, dtype);
		const PlaneTma px{x, PQ, g.C, g.N}, pdy{dy, PQ, g.K, g.N};
		if (tma_x)
			st = launch(p, dtype, bn, MODE_K_POS_TMA, MODE_K_POS_TMA, false, g.G, nullptr, pz_stream(stream), &px, &pdy);
		else if (tma_dy && !dense_x && tl >= 2 && RS <= 31)
			st = launch(p, dtype, bn, MODE_K_POS_TAP, MODE_K_POS_TMA, false, g.G, nullptr, pz_stream(stream), nullptr, &pdy);
		else
			st = launch(p, dtype, bn, amode, bmode, !dense_x && RS > 31, g.G, nullptr, pz_stream(stream));
	}
	if (st != PZ_OK || !acc) return st;
	return finalize16(dtype, dw, wcount, acc, 1, wcount, beta, pz_stream(stream));
}

int pz_bias_grad(int dtype, const void* t, void* db, int64_t N, int64_t C, int64_t S, float alpha, float beta, void* stream)
{
	PZ_REQUIRE(dtype == PZ_F32 || dtype == PZ_F16 || dtype == PZ_BF16, "bias_grad: unsupported dtype %d", dtype);
	if (C <= 0) return PZ_OK;
	const bool h16 = dtype != PZ_F32;
	float* acc = (float*)db;
	int st;
	if (h16) {
		// 16-bit bias gradient: reduce into an fp32 scratch, then db = beta*db + acc
		acc = (float*)pz_scratch((size_t)C * sizeof(float));
		if (!acc) { pz_set_error(PZ_ERR_MEMORY, "bias_grad: cannot allocate the accumulator"); return PZ_ERR_MEMORY; }
		st = pz_memset8(acc, 0, (size_t)C * 4, stream);
	} else
		st = prescale(acc, C, beta, stream);
	if (st != PZ_OK) return st;
	if (N > 0 && S > 0) {
		int nsplit = (int)(pz_cdiv(4ll * pz_num_sms(), C));
		if (nsplit > N) nsplit = (int)N;
		if (nsplit < 1) nsplit = 1;
		dim3 grid((unsigned)C, (unsigned)nsplit);
		if (dtype == PZ_F32) bias_grad_kernel<256, float><<<grid, 256, 0, pz_stream(stream)>>>((const float*)t, acc, N, C, S, alpha, nsplit);
		else if (dtype == PZ_F16) bias_grad_kernel<256, __half><<<grid, 256, 0, pz_stream(stream)>>>((const __half*)t, acc, N, C, S, alpha, nsplit);
		else bias_grad_kernel<256, __nv_bfloat16><<<grid, 256, 0, pz_stream(stream)>>>((const __nv_bfloat16*)t, acc, N, C, S, alpha, nsplit);
		pz_count_launch(1);
		PZ_LAUNCH_CHECK();
	}
	if (h16) return finalize16(dtype, db, C, acc, 1, C, beta, pz_stream(stream));
	return PZ_OK;
}

}  // extern "C"
