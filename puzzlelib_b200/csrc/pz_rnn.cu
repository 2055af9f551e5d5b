// pz_rnn.cu -- pointwise cells of the recurrent layers (LSTM, ReLU / tanh RNN), forward and backward.
//
// Replaces the cell math inside cudnnRNNForwardTraining / cudnnRNNBackwardData as driven by the reference's Rnn object
// (Cuda/Source/Libs/CuDnnRnn.c:565-1000; gate order and formulas pinned by the reference's own host loops,
// Cuda/Wrappers/CuDnnRnn.py:178-300).  The matrix products of a recurrent layer are ordinary GEMMs on the tcgen05 engine
// (pz_gemm): one (T*B x in) x (in x 4H) product for all time steps of the input projection, one (B x H) x (H x 4H) product
// per step for the recurrence; these kernels do everything between two GEMMs of a step in one pass over (B x 4H).
//
// Gate order in every 4H-wide row: i, f, c (candidate, tanh), o -- the cuDNN linear-layer order 0..3 that
// acquireLSTMParams names "wi/wf/wc/wo" (Cuda/Backend.py:264-306).
#include "pz_common.h"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// gates[b][4H]: pre-activations (input + recurrent projections) in, activations out (kept for the backward pass)
__global__ void __launch_bounds__(kThreads) lstm_cell_fwd_kernel(float* __restrict__ gates, const float* __restrict__ bw,
																  const float* __restrict__ br, const float* __restrict__ c_prev,
																  float* __restrict__ c_out, float* __restrict__ h_out, int B, int H)
{
	const int idx = blockIdx.x * kThreads + threadIdx.x;
	if (idx >= B * H) return;
	const int b = idx / H, j = idx - b * H;
	float* g = gates + (size_t)b * 4 * H;
	const float pi = g[j] + bw[j] + br[j];
	const float pf = g[H + j] + bw[H + j] + br[H + j];
	const float pc = g[2 * H + j] + bw[2 * H + j] + br[2 * H + j];
	const float po = g[3 * H + j] + bw[3 * H + j] + br[3 * H + j];
	const float i = sigmoidf_(pi), f = sigmoidf_(pf), cc = tanhf(pc), o = sigmoidf_(po);
	const float cp = c_prev ? c_prev[idx] : 0.0f;
	const float c = f * cp + i * cc;
	g[j] = i;
	g[H + j] = f;
	g[2 * H + j] = cc;
	g[3 * H + j] = o;
	c_out[idx] = c;
	h_out[idx] = o * tanhf(c);
}

// dh = dy (+ dh_next); dgates = gradient w.r.t. the pre-activations; dc_io: dc from step t+1 in, dc for step t-1 out
__global__ void __launch_bounds__(kThreads) lstm_cell_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dh_next,
																  float* __restrict__ dc_io, const float* __restrict__ acts,
																  const float* __restrict__ c, const float* __restrict__ c_prev,
																  float* __restrict__ dgates, int B, int H, int first)
{
	const int idx = blockIdx.x * kThreads + threadIdx.x;
	if (idx >= B * H) return;
	const int b = idx / H, j = idx - b * H;
	const float* a = acts + (size_t)b * 4 * H;
	float* dg = dgates + (size_t)b * 4 * H;
	const float i = a[j], f = a[H + j], cc = a[2 * H + j], o = a[3 * H + j];
	const float dh = dy[idx] + (dh_next ? dh_next[idx] : 0.0f);
	const float tc = tanhf(c[idx]);
	const float dc = dh * o * (1.0f - tc * tc) + (first ? 0.0f : dc_io[idx]);
	const float cp = c_prev ? c_prev[idx] : 0.0f;
	dg[j] = dc * cc * i * (1.0f - i);
	dg[H + j] = dc * cp * f * (1.0f - f);
	dg[2 * H + j] = dc * i * (1.0f - cc * cc);
	dg[3 * H + j] = dh * tc * o * (1.0f - o);
	dc_io[idx] = dc * f;
}

// plain RNN: h = act(pre + bw + br), act = relu (mode 0) / tanh (mode 1); pre[b][H] in place
__global__ void __launch_bounds__(kThreads) rnn_cell_fwd_kernel(float* __restrict__ h, const float* __restrict__ bw,
																 const float* __restrict__ br, int B, int H, int mode)
{
	const int idx = blockIdx.x * kThreads + threadIdx.x;
	if (idx >= B * H) return;
	const int j = idx % H;
	const float p = h[idx] + bw[j] + br[j];
	h[idx] = mode == 0 ? fmaxf(p, 0.0f) : tanhf(p);
}

__global__ void __launch_bounds__(kThreads) rnn_cell_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dh_next,
																 const float* __restrict__ h, float* __restrict__ dpre, int n, int mode)
{
	const int idx = blockIdx.x * kThreads + threadIdx.x;
	if (idx >= n) return;
	const float dh = dy[idx] + (dh_next ? dh_next[idx] : 0.0f);
	const float y = h[idx];
	dpre[idx] = mode == 0 ? (y > 0.0f ? dh : 0.0f) : dh * (1.0f - y * y);
}

// GRU (cuDNN formulation, reference host loop Cuda/Wrappers/CuDnnRnn.py:303-352): gate order r, i, h in every 3H row;
//   r = sigm(Wr x + Rr h + bwr + brr), i = sigm(Wi x + Ri h + bwi + bri), h~ = tanh(Wh x + bwh + r * (Rh h + brh)),
//   h' = (1 - i) * h~ + i * h.
// gx: input projections in, activations (r, i, h~) out; gh: recurrent projections in, its h part becomes q = Rh h + brh (kept
// for the backward pass).
__global__ void __launch_bounds__(kThreads) gru_cell_fwd_kernel(float* __restrict__ gx, float* __restrict__ gh, const float* __restrict__ bw,
																 const float* __restrict__ br, const float* __restrict__ h_prev,
																 float* __restrict__ h_out, int B, int H)
{
	const int idx = blockIdx.x * kThreads + threadIdx.x;
	if (idx >= B * H) return;
	const int b = idx / H, j = idx - b * H;
	float* x = gx + (size_t)b * 3 * H;
	float* h = gh + (size_t)b * 3 * H;
	const float r = sigmoidf_(x[j] + h[j] + bw[j] + br[j]);
	const float i = sigmoidf_(x[H + j] + h[H + j] + bw[H + j] + br[H + j]);
	const float q = h[2 * H + j] + br[2 * H + j];
	const float ht = tanhf(x[2 * H + j] + bw[2 * H + j] + r * q);
	const float hp = h_prev ? h_prev[idx] : 0.0f;
	x[j] = r;
	x[H + j] = i;
	x[2 * H + j] = ht;
	h[2 * H + j] = q;
	h_out[idx] = (1.0f - i) * ht + i * hp;
}

// dgx / dgh: gradients w.r.t. the input-side / recurrent-side pre-activations; dh_carry = dh * i (the direct path to h[t-1],
// the caller adds dgh * R to it)
__global__ void __launch_bounds__(kThreads) gru_cell_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ dh_next,
																 const float* __restrict__ acts, const float* __restrict__ gh,
																 const float* __restrict__ h_prev, float* __restrict__ dgx,
																 float* __restrict__ dgh, float* __restrict__ dh_carry, int B, int H)
{
	const int idx = blockIdx.x * kThreads + threadIdx.x;
	if (idx >= B * H) return;
	const int b = idx / H, j = idx - b * H;
	const float* a = acts + (size_t)b * 3 * H;
	const float r = a[j], i = a[H + j], ht = a[2 * H + j];
	const float q = gh[(size_t)b * 3 * H + 2 * H + j];
	const float hp = h_prev ? h_prev[idx] : 0.0f;
	const float dh = dy[idx] + (dh_next ? dh_next[idx] : 0.0f);
	const float dpre_h = dh * (1.0f - i) * (1.0f - ht * ht);
	const float dpre_i = dh * (hp - ht) * i * (1.0f - i);
	const float dpre_r = dpre_h * q * r * (1.0f - r);
	float* gx = dgx + (size_t)b * 3 * H;
	float* gr = dgh + (size_t)b * 3 * H;
	gx[j] = dpre_r;
	gx[H + j] = dpre_i;
	gx[2 * H + j] = dpre_h;
	gr[j] = dpre_r;
	gr[H + j] = dpre_i;
	gr[2 * H + j] = dpre_h * r;
	dh_carry[idx] = dh * i;
}

}  // namespace

extern "C" {

int pz_gru_cell_fwd(float* gx, float* gh, const float* bw, const float* br, const float* h_prev, float* h_out, int64_t B, int64_t H,
					void* stream)
{
	PZ_REQUIRE(B > 0 && H > 0 && B * H < (1ll << 31), "gru cell: invalid size %lld x %lld", (long long)B, (long long)H);
	gru_cell_fwd_kernel<<<(unsigned)pz_cdiv(B * H, kThreads), kThreads, 0, pz_stream(stream)>>>(gx, gh, bw, br, h_prev, h_out, (int)B, (int)H);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

int pz_gru_cell_bwd(const float* dy, const float* dh_next, const float* acts, const float* gh, const float* h_prev, float* dgx,
					float* dgh, float* dh_carry, int64_t B, int64_t H, void* stream)
{
	PZ_REQUIRE(B > 0 && H > 0 && B * H < (1ll << 31), "gru cell: invalid size %lld x %lld", (long long)B, (long long)H);
	gru_cell_bwd_kernel<<<(unsigned)pz_cdiv(B * H, kThreads), kThreads, 0, pz_stream(stream)>>>(dy, dh_next, acts, gh, h_prev, dgx, dgh,
																									dh_carry, (int)B, (int)H);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}


int pz_lstm_cell_fwd(float* gates, const float* bw, const float* br, const float* c_prev, float* c_out, float* h_out, int64_t B,
					 int64_t H, void* stream)
{
	PZ_REQUIRE(B > 0 && H > 0 && B * H < (1ll << 31), "lstm cell: invalid size %lld x %lld", (long long)B, (long long)H);
	lstm_cell_fwd_kernel<<<(unsigned)pz_cdiv(B * H, kThreads), kThreads, 0, pz_stream(stream)>>>(gates, bw, br, c_prev, c_out, h_out, (int)B,
																									 (int)H);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

int pz_lstm_cell_bwd(const float* dy, const float* dh_next, float* dc_io, const float* acts, const float* c, const float* c_prev,
					 float* dgates, int64_t B, int64_t H, int first, void* stream)
{
	PZ_REQUIRE(B > 0 && H > 0 && B * H < (1ll << 31), "lstm cell: invalid size %lld x %lld", (long long)B, (long long)H);
	lstm_cell_bwd_kernel<<<(unsigned)pz_cdiv(B * H, kThreads), kThreads, 0, pz_stream(stream)>>>(dy, dh_next, dc_io, acts, c, c_prev, dgates,
																									 (int)B, (int)H, first);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

int pz_rnn_cell_fwd(float* h, const float* bw, const float* br, int64_t B, int64_t H, int mode, void* stream)
{
	PZ_REQUIRE(B > 0 && H > 0 && B * H < (1ll << 31), "rnn cell: invalid size %lld x %lld", (long long)B, (long long)H);
	PZ_REQUIRE(mode == 0 || mode == 1, "rnn cell: mode must be 0 (relu) or 1 (tanh)");
	rnn_cell_fwd_kernel<<<(unsigned)pz_cdiv(B * H, kThreads), kThreads, 0, pz_stream(stream)>>>(h, bw, br, (int)B, (int)H, mode);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

int pz_rnn_cell_bwd(const float* dy, const float* dh_next, const float* h, float* dpre, int64_t n, int mode, void* stream)
{
	PZ_REQUIRE(n > 0 && n < (1ll << 31), "rnn cell: invalid size %lld", (long long)n);
	PZ_REQUIRE(mode == 0 || mode == 1, "rnn cell: mode must be 0 (relu) or 1 (tanh)");
	rnn_cell_bwd_kernel<<<(unsigned)pz_cdiv(n, kThreads), kThreads, 0, pz_stream(stream)>>>(dy, dh_next, h, dpre, (int)n, mode);
	pz_count_launch(1);
	PZ_LAUNCH_CHECK();
	return PZ_OK;
}

}  // extern "C"
