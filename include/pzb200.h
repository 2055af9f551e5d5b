/* pzb200.h -- C-ABI of libpzb200.so: a B200 (sm_100a) native replacement for the per-operator GPU
 * backend of PuzzleLib (the cuDNN/cuBLAS/NVRTC layer under Cuda/Backend.py).
 *
 * Every function returns 0 on success or a PZ_ERR_* code; pz_last_error() gives the message of the last
 * failure on the calling thread.  All pointers named x/y/dx/... are DEVICE pointers unless the name says
 * "host".  `stream` is a cudaStream_t passed as void* (NULL = the library's default stream, which is the
 * CUDA legacy default stream -- the reference enqueues everything there, SURVEY 8b "Threading / streams").
 * Tensors are dense row-major NCHW / (rows, cols), as in the reference (CuDnn.c:137-158,202-204).
 *
 * Each entry names the reference interface it replaces (paths relative to the PuzzleLib tree).
 */
#ifndef PZB200_H
#define PZB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- status / dtypes / enums */
enum { PZ_OK = 0, PZ_ERR_VALUE = 1, PZ_ERR_CUDA = 2, PZ_ERR_MEMORY = 3, PZ_ERR_NCCL = 4, PZ_ERR_UNSUPPORTED = 5 };

/* dtype codes (reference: Cuda_DataType, Cuda/Source/Core/Driver.h:164-194; bf16 is new) */
enum {
	PZ_F32 = 0, PZ_F16 = 1, PZ_BF16 = 2, PZ_F64 = 3,
	PZ_I8 = 4, PZ_U8 = 5, PZ_I16 = 6, PZ_U16 = 7, PZ_I32 = 8, PZ_U32 = 9, PZ_I64 = 10, PZ_U64 = 11
};

/* activation kinds (reference: Cuda/Kernels/ElementWise.py:9-492) */
enum {
	PZ_ACT_SIGMOID = 0, PZ_ACT_TANH = 1, PZ_ACT_RELU = 2, PZ_ACT_LEAKYRELU = 3, PZ_ACT_ELU = 4,
	PZ_ACT_SOFTPLUS = 5, PZ_ACT_CLIP = 6, PZ_ACT_GELU = 7
};

/* pooling modes: values of cudnnPoolingMode_t, exported by the reference as CuDnn.POOL_MODE_*
 * (Cuda/Backend.py:104-108) */
enum { PZ_POOL_MAX = 0, PZ_POOL_AVG_WITH_PAD = 1, PZ_POOL_AVG_NO_PAD = 2, PZ_POOL_MAX_DETERMINISM = 3 };

/* batch-norm modes: cudnnBatchNormMode_t (Cuda/Backend.py:122-125) */
enum { PZ_BN_PER_ACTIVATION = 0, PZ_BN_SPATIAL = 1, PZ_BN_SPATIAL_PERSISTENT = 2 };

/* softmax modes: cudnnSoftmaxMode_t (Cuda/Backend.py:111-113): per-activation = INSTANCE, spatial = CHANNEL */
enum { PZ_SOFTMAX_PER_ACTIVATION = 0, PZ_SOFTMAX_SPATIAL = 1 };

const char* pz_last_error(void);
int pz_version(void);

/* ---------------------------------------------------------------- device / memory / streams
 * replaces Cuda/Source/Core/{Device,Buffer,Allocator,Stream}.c */
int pz_device_count(int* count);                               /* Device.c  Device.count()        */
int pz_device_set(int index);                                  /* Device.c  Device.set()          */
int pz_device_get(int* index);
int pz_device_name(int index, char* buf, int buflen);          /* Device.c  Device.name()         */
int pz_device_sm_count(int* count);
int pz_device_cc(int* major, int* minor);
int pz_device_synchronize(void);                               /* Device.c  Device.synchronize()  */
int pz_mem_info(size_t* free_bytes, size_t* total_bytes);      /* Driver.c  getMemoryInfo()       */

int pz_malloc(void** ptr, size_t nbytes);                      /* Buffer.c  Cuda_Buffer_init      */
int pz_free(void* ptr);
int pz_host_alloc(void** ptr, size_t nbytes);                  /* pinned host staging             */
int pz_host_free(void* ptr);

/* caching allocator with the reference's bin function: sizes rounded up to (4..7)*2^e, i.e. two mantissa
 * bits (Allocator.c:29-67); blocks return to their bin and go back to the driver only on free_held */
int pz_pool_create(void** pool);                               /* Allocator.c MemoryPool()        */
int pz_pool_destroy(void* pool);
int pz_pool_alloc(void* pool, size_t nbytes, void** ptr, size_t* granted);
int pz_pool_release(void* pool, void* ptr, size_t granted);    /* Allocator.c MemoryPool_hold     */
int pz_pool_free_held(void* pool);                             /* Allocator.c freeHeld()          */
int pz_pool_stats(void* pool, size_t* held_blocks, size_t* held_bytes, size_t* active_blocks, size_t* active_bytes);
size_t pz_pool_alloc_size(size_t nbytes);                      /* bin -> block size, for tests    */

int pz_memcpy_h2d(void* dst, const void* host_src, size_t nbytes, void* stream, int async);
int pz_memcpy_d2h(void* host_dst, const void* src, size_t nbytes, void* stream, int async);
int pz_memcpy_d2d(void* dst, const void* src, size_t nbytes, void* stream);
/* kind: 0 = d2d, 1 = h2d, 2 = d2h (Driver.c memcpy2D, used by GPUArray slices and concatenate/split/tile) */
int pz_memcpy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, int kind,
				void* stream);
int pz_memset8(void* ptr, uint8_t value, size_t count, void* stream);    /* Buffer.c fillD8  */
int pz_memset16(void* ptr, uint16_t value, size_t count, void* stream);  /* Buffer.c fillD16 */
int pz_memset32(void* ptr, uint32_t value, size_t count, void* stream);  /* Buffer.c fillD32 */

/* NULL stream arguments resolve to the library's current stream: the legacy default stream (what the reference uses for every
   call, SURVEY 8b "Threading / streams") unless redirected here */
int pz_set_default_stream(void* stream);
/* whole-step CUDA graphs (launch-overhead removal behind Sequential / Handler.handle, Containers/Sequential.py:186-234):
   capture everything the operator API enqueues between begin and end, replay it with one launch */
int pz_graph_begin(void* stream);
int pz_graph_end(void* stream, void** exec);
int pz_graph_launch(void* exec, void* stream);
int pz_graph_destroy(void* exec);
/* ---- kernel modules next to the hot path (backend.prelumod / padmod / embedmod / upsamplemod) ----
 * PReLU over [N][C][S] float32 tensors, one slope per channel or (shared != 0) one for all -- Cuda/Kernels/PRelu.py:69-132 */
int pz_prelu_fwd(const void* x, const void* slopes, void* y, int64_t N, int64_t C, int64_t S, int shared, void* stream);
int pz_prelu_bwd_data(const void* dy, const void* slopes, const void* x, void* dx, int64_t N, int64_t C, int64_t S, int shared, void* stream);
/* dslopes[c] (or dslopes[0]) = sum dy * x * (x <= 0); deterministic block reduction */
int pz_prelu_bwd_params(const void* x, const void* dy, void* dslopes, int64_t N, int64_t C, int64_t S, int shared, void* stream);
/* reflection padding of `planes` H x W maps (1-d: H = 1) by (up, bottom, left, right); float32 / float16 -- Cuda/Kernels/Pad.py:155-229.
 * The backward pass gathers (no atomics) and overwrites dx */
int pz_reflectpad_fwd(int dtype, const void* x, void* y, int64_t planes, int H, int W, int up, int bp, int lp, int rp, void* stream);
int pz_reflectpad_bwd(int dtype, const void* dy, void* dx, int64_t planes, int H, int W, int up, int bp, int lp, int rp, void* stream);
/* embedding lookup out[i] = W[idx[i]] (index -1: zero row) and the vocabulary update W[idx[i]] += scale * grad[i] --
 * Cuda/Kernels/Embedder.py:56-87 */
int pz_embed_fwd(int dtype, const void* idx, const void* W, void* out, int64_t size, int64_t emb, void* stream);
int pz_embed_bwd(int dtype, const void* idx, const void* grad, void* W, float scale, int64_t size, int64_t emb, void* stream);
/* up-sampling of `planes` D x H x W float32 volumes (2-d: D = 1) by integer factors -- Cuda/Kernels/Upsample.py:313-454.
 * linear: r* = (in - 1) / (out - 1) as float32; the backward pass accumulates into a zeroed dx */
int pz_upsample_nearest_fwd(const void* x, void* y, int64_t planes, int D, int H, int W, int ds, int hs, int ws, void* stream);
int pz_upsample_nearest_bwd(const void* dy, void* dx, int64_t planes, int D, int H, int W, int ds, int hs, int ws, void* stream);
int pz_upsample_linear_fwd(const void* x, void* y, int64_t planes, int D, int H, int W, int oD, int oH, int oW, float rd, float rh, float rw,
						   int three_d, void* stream);
int pz_upsample_linear_bwd(const void* dy, void* dx, int64_t planes, int D, int H, int W, int oD, int oH, int oW, float rd, float rh, float rw,
						   int three_d, void* stream);
/* mapLRN with a means tensor = divisive normalisation over `planes` H x W maps (cudnnDivisiveNormalization, CuDnnNorm.c:329-510;
 * the LCN module): y = x * (K + alpha/n^2 * sum_win (x_j - m_i)^2)^-beta; the backward pass returns dx and dmeans and needs
 * planes*H*W floats of scratch */
int pz_divnorm_fwd(int dtype, const void* x, const void* means, void* y, int64_t planes, int64_t H, int64_t W, int n, float alpha, float beta, float K,
				   void* stream);
int pz_divnorm_bwd(int dtype, const void* x, const void* means, const void* grad, void* dx, void* dmeans, void* tmp, int64_t planes, int64_t H,
				   int64_t W, int n, float alpha, float beta, float K, void* stream);
/* spatial transformer (cudnnSpatialTfGridGenerator* + cudnnSpatialTfSampler*, CuDnnSpatialTf.c:20-222): theta [B][2][3] -> grid
 * [B][oH][oW][2] -> out [B][C][oH][oW] by bilinear sampling of data [B][C][H][W]; the backward pass returns dx (zeroed here),
 * dtheta [B][2][3] and dgrid; float32 / float16 */
int pz_spatialtf_fwd(int dtype, const void* data, const void* theta, void* grid, void* out, int64_t B, int64_t C, int H, int W, int oH, int oW,
					 void* stream);
int pz_spatialtf_bwd(int dtype, const void* grad, const void* data, const void* grid, void* dx, void* dtheta, void* dgrid, int64_t B, int64_t C,
					 int H, int W, int oH, int oW, void* stream);
/* CTC loss over softmax outputs y [T][B][V] (Cuda/Kernels/CTC.py:232-269): labels = the B label strings concatenated, offsets[B + 1]
 * their prefix sums, datalen[B] the valid time steps; writes alphas (T * (2 * offsets[B] + B) floats, the reference's layout), nll[B],
 * adds sum(nll) to *error and writes grad [T][B][V] (zeroed by the caller) = d(sum nll) / d y */
int pz_ctc_loss(const void* y, const void* datalen, const void* labels, const void* offsets, void* alphas, void* nll, void* error, void* grad, int T,
				int B, int V, int max_label_len, int blank, void* stream);
/* diagnosis: the 32 per-role phase clock sums of the tcgen05 engine since the last call (only in builds with -DPZ_TIMELINE,
 * PZ_ERR_UNSUPPORTED otherwise); no reference counterpart */
int pz_debug_timeline(unsigned long long* out);
int pz_stream_create(void** stream);
int pz_stream_destroy(void* stream);
int pz_stream_synchronize(void* stream);
int pz_event_create(void** event);
int pz_event_create_sync(void** event);                      /* stream-ordering only (no timing), usable under capture */
int pz_event_destroy(void* event);
int pz_event_record(void* event, void* stream);
int pz_event_synchronize(void* event);
int pz_event_elapsed_ms(void* start, void* stop, float* ms);
int pz_stream_wait_event(void* stream, void* event);

/* number of kernels this library has launched on the calling process since load (bench "gpu_launches") */
uint64_t pz_launch_count(void);

/* per-launch CUDA-event profiling of one kernel family (bench.py roofline numbers). family: 0 tcgen05 GEMM/conv engine,
 * 1 batch-norm forward, 2 batch-norm backward, 3 elementwise, 4 pooling, 5 other; -1 = all. flops / bytes are the
 * ALGORITHMIC ones of the recorded launches (operands read once, results written once). */
int pz_profile_enable(int on);
int pz_profile_collect(int family, double* total_ms, double* flops, double* bytes, uint64_t* launches);

/* ---------------------------------------------------------------- elementwise (bandwidth-bound)
 * replaces the NVRTC-JIT kernels of Cuda/Kernels/ElementWise.py and Cuda/GPUArray.py */
/* out = f(in);  a, b are the optional scalars (leakyRelu a, elu a, clip a,b). ElementWise.py:9-445 */
int pz_act_fwd(int kind, int dtype, void* out, const void* in, int64_t n, float a, float b, void* stream);
/* ingrad = f'(outgrad, ref) where ref = the activation's OUTPUT (its INPUT for gelu). ElementWise.py:36-492 */
int pz_act_bwd(int kind, int dtype, void* ingrad, const void* outgrad, const void* ref, int64_t n, float a, float b,
			   void* stream);
int pz_axpy(int dtype, void* y, const void* x, float alpha, int64_t n, void* stream);       /* toVectorAddVectorKer :582-606  */
int pz_axpby(int dtype, void* out, const void* x, float alpha, const void* y, float beta, int64_t n,
			 void* stream);                                                                 /* addKer :1017-1045 (also Grid.py:133) */
int pz_scale_shift(int dtype, void* out, const void* in, float a, float b, int64_t n, void* stream); /* linearKer :1073-1099 */
int pz_mul(int dtype, void* out, const void* a, const void* b, int64_t n, void* stream);    /* mulKer :1047-1071 */
int pz_add2(int dtype, void* out, const void* a, const void* b, int64_t n, void* stream);   /* Add.py:15-23 as one pass */
/* y = (0 + x1*a1) + x2*a2 (x2 may be NULL: y = 0 + x1*a1): the launch sequence fill(0) + toVectorAddVector (+ toVectorAddVector)
 * of Modules/Add.py:15-23 / Replicate.py:18-29 fused into one pass with identical bits (intermediate rounded to the storage type) */
int pz_axpy2(int dtype, void* y, const void* x1, float a1, const void* x2, float a2, int64_t n, void* stream);
/* the same fused sum followed by the ReLU the next module applies: y = (0 + a1*x1) + a2*x2, out = max(y, 0); both stored
 * (Modules/Add.py:15-23 then Modules/Activation.py:69-71 in one pass over the tensors) */
int pz_axpy2_relu(int dtype, void* y, void* out, const void* x1, float a1, const void* x2, float a2, int64_t n, void* stream);
/* ... or by the ReLU derivative: y = (0 + a1*x1) + a2*x2, ingrad = y * (ref > 0) (Modules/Replicate.py:24-29 then
 * Modules/Activation.py:74-76) */
int pz_axpy2_relu_bwd(int dtype, void* y, void* ingrad, const void* x1, float a1, const void* x2, float a2, const void* ref, int64_t n,
					  void* stream);
/* float32 math mode of pz_conv2d_* and pz_gemm (reference: enableTensorOps, CuDnn.c:61-74 / CuBlas.c:91-106): 0 = TF32
 * tensor-core products with fp32 accumulation (default; within 1e-3 of fp32), 1 = exact fp32 FMAs on the CUDA cores (what
 * cuDNN / cuBLAS give the reference's float32 tensors on this stack; a verification path, not tuned) */
int pz_set_exact_fp32(int on);
int pz_exact_fp32(void);
/* `slice=` launches (Cuda/SourceModule.py:162-200, the `<name>_strided` twin of every ElementwiseKernel; callers:
 * Modules/Activation.py:71-76, NoiseInjector.py:73-93, Dropout.py:58-74): elements start, start + step, ... < min(stop, n) */
int pz_act_fwd_slice(int kind, int dtype, void* out, const void* in, int64_t n, float a, float b, int64_t start, int64_t stop,
					 int64_t step, void* stream);
int pz_act_bwd_slice(int kind, int dtype, void* ingrad, const void* outgrad, const void* ref, int64_t n, float a, float b,
					 int64_t start, int64_t stop, int64_t step, void* stream);
int pz_axpby_slice(int dtype, void* out, const void* x, float alpha, const void* y, float beta, int64_t n, int64_t start,
				   int64_t stop, int64_t step, void* stream);
int pz_mul_slice(int dtype, void* out, const void* a, const void* b, int64_t n, int64_t start, int64_t stop, int64_t step,
				 void* stream);

/* The remaining elementwise kernels `Backend/Kernels/ElementWise.py:96-125` and `Costs.py:45-49` bind from the backend object.
 * One entry: `ptrs` are the kernel's pointer arguments in the reference's order, `scalars` its float arguments, `aux` its two
 * int arguments (pointwise costs); element i = start + k * step < min(stop, n).
 *   PZ_EW_ABS (out, in)                                  absKer           ElementWise.py:1117
 *   PZ_EW_WEIGHT_DECAY (grad, param; rate)               weightDecayKer   :1109    grad -= rate * param
 *   PZ_EW_L1_PENALTY (outgrad, ingrad, data; a)          l1penaltyKer     :1125
 *   PZ_EW_L1_GRAD (grad, pred, target; norm)             l1gradKer        :1133
 *   PZ_EW_RBM (out, in, uni)                             rbmKer           :1100    out = uni < sigmoid(in)
 *   PZ_EW_RMSPROP (param, grad, ms; lr, factor, eps)     rmspropKer       :860
 *   PZ_EW_RMSPROP_GRAVES (param, grad, mg, ms, delta; lr, alpha, momRate, eps)     :906
 *   PZ_EW_ADAGRAD (param, grad, h; lr, eps)              adagradKer       :664
 *   PZ_EW_ADADELTA (param, grad, msg, msdx; rho, eps)    adadeltaKer      :614
 *   PZ_EW_SMORMS3 (param, grad, mem f32, mg f32, ms f32; lr, eps)  smorms3Ker :957
 *   PZ_EW_BCE (scores, labels i32, totalError, grad; aux numsamples, spatialDim)   Costs.py:8
 *   PZ_EW_HINGE (scores, labels i32, totalError, grad; aux numsamples, numcases)   Costs.py:25
 *   PZ_EW_SMOOTH_L1 (pred, target, totalError, grad; norm, fullnorm)               Costs.py:42
 *   PZ_EW_L1_HINGE (x1, x2, labels i32, totalError, g1, g2; aux numsamples, numcases)  Costs.py:58 */
enum {
	PZ_EW_ABS = 0, PZ_EW_WEIGHT_DECAY = 1, PZ_EW_L1_PENALTY = 2, PZ_EW_L1_GRAD = 3, PZ_EW_RBM = 4, PZ_EW_RMSPROP = 5,
	PZ_EW_RMSPROP_GRAVES = 6, PZ_EW_ADAGRAD = 7, PZ_EW_ADADELTA = 8, PZ_EW_SMORMS3 = 9,
	PZ_EW_BCE = 10, PZ_EW_HINGE = 11, PZ_EW_SMOOTH_L1 = 12, PZ_EW_L1_HINGE = 13
};
int pz_eltwise(int op, int dtype, void* const* ptrs, int nptrs, const float* scalars, int nscalars, const int* aux, int64_t n,
			   int64_t start, int64_t stop, int64_t step, void* stream);
int pz_cast(int dst_dtype, void* dst, int src_dtype, const void* src, int64_t n, void* stream); /* GPUArray.py astype */
/* dst[r][i] += src[r][i], pitched rows (3-d transposed convolution: scatter-add of per-slice gradients) */
int pz_add2d(int dtype, void* dst, int64_t dpitch, const void* src, int64_t spitch, int64_t width, int64_t rows, void* stream);
int pz_fill64(void* ptr, uint64_t value, int64_t count, void* stream);                      /* GPUArray.py:167-180 8-byte fill */
/* mom = momRate*mom + learnRate*grad; param += mom  (ElementWise.py:771-800 classicMomSGD) */
int pz_sgd_momentum(int dtype, void* param, const void* grad, void* mom, float learn_rate, float mom_rate,
					int64_t n, void* stream);
/* min / max reduction of a float tensor into a 1-element device buffer (GPUArray.py min/max) */
int pz_reduce_minmax(int dtype, const void* in, int64_t n, int want_max, void* out, void* stream);

/* ---------------------------------------------------------------- matrix-vector helpers
 * replaces Cuda/Kernels/MatVec.py */
/* out[z][y][x] = mat[z][y][x] + vec: axis=1, vecdim==cols: vec[z][x]; axis=1, vecdim<cols: vec[x % vecdim];
 * axis=0: vec[z][y]   (MatVec.py:128-171,346-374) */
int pz_addvec2mat(int dtype, void* out, const void* mat, const void* vec, int64_t z, int64_t rows, int64_t cols,
				  int axis, int64_t vecdim, void* stream);
/* tensor viewed as [z][h][w]: reduce_rows=1 -> out[z][h] = beta*out + alpha*sum_w ; else out[z][w] = beta*out +
 * alpha*sum_h   (MatVec.py:60-91,273-308) */
int pz_matsum(int dtype, void* out, const void* tensor, int64_t z, int64_t h, int64_t w, int reduce_rows,
			  float alpha, float beta, void* stream);
/* argmax/argmin along the middle axis of [z][h][w] (w=1 for the last-axis case after the caller's view)
 * -> int32 idx[z][w]; first occurrence wins (MatVec.py:8-57,231-268) */
int pz_argminmax(int dtype, int32_t* idx, const void* tensor, int64_t z, int64_t h, int64_t w, int want_max,
				 void* stream);

/* ---------------------------------------------------------------- batch normalisation
 * replaces cudnnBatchNormalizationForwardTraining/Inference/Backward (CuDnnNorm.c:31-71,158-194).
 * x, y: [N][C][S] of `dtype`; scale/bias/mean/var/saves: fp32 [C] (CuDnnNorm.c:118-121). */
int pz_bn_fwd_train(int dtype, const void* x, void* y, int64_t N, int64_t C, int64_t S, const float* scale,
					const float* bias, float* running_mean, float* running_var, float* save_mean,
					float* save_invvar, double eps, double factor, void* stream);
/* pz_bn_fwd_train followed by the ReLU the next module applies to y: y AND z = y * (y > 0) are stored; in the cluster kernel by
 * the same pass (Modules/BatchNormND.py:66-78 then Modules/Activation.py:69-71), otherwise by the plain ReLU kernel */
int pz_bn_fwd_train_relu(int dtype, const void* x, void* y, void* z, int64_t N, int64_t C, int64_t S, const float* scale, const float* bias,
						 float* running_mean, float* running_var, float* save_mean, float* save_invvar, double eps, double factor,
						 void* stream);
int pz_bn_fwd_infer(int dtype, const void* x, void* y, int64_t N, int64_t C, int64_t S, const float* scale,
					const float* bias, const float* mean, const float* var, double eps, void* stream);
int pz_bn_bwd(int dtype, const void* x, const void* dy, void* dx, int64_t N, int64_t C, int64_t S,
			  const float* scale, const float* save_mean, const float* save_invvar, float* dscale, float* dbias,
			  void* stream);
/* pz_bn_bwd followed by the parameter-gradient accumulation of BatchNormND.accGradParams (Modules/BatchNormND.py:74-83: two
 * addVectorToVector launches, Backend/Blas.py:43-58 -> ElementWise.py:1030) in the same pass:
 *   scale_acc = scale_alpha * dscale + scale_beta * scale_acc,  bias_acc = bias_alpha * dbias + bias_beta * bias_acc
 * (same bits as the two launches; either accumulator may be NULL; they must not overlap another tensor of the call) */
int pz_bn_bwd_acc(int dtype, const void* x, const void* dy, void* dx, int64_t N, int64_t C, int64_t S,
				  const float* scale, const float* save_mean, const float* save_invvar, float* dscale, float* dbias,
				  float* scale_acc, float scale_alpha, float scale_beta, float* bias_acc, float bias_alpha, float bias_beta,
				  void* stream);

/* ---------------------------------------------------------------- pooling
 * replaces cudnnPoolingForward/Backward (CuDnnPool.c:81,171) -- 2-D; 1-D is H=1 */
int pz_pool2d_fwd(int dtype, int mode, const void* x, void* y, int64_t planes, int H, int W, int OH, int OW,
				  int fh, int fw, int sh, int sw, int ph, int pw, void* stream);
int pz_pool2d_bwd(int dtype, int mode, const void* x, const void* y, const void* dy, void* dx, int64_t planes,
				  int H, int W, int OH, int OW, int fh, int fw, int sh, int sw, int ph, int pw, void* stream);
/* fp32 max-pool with int32 argmax mask, bit-exact with Cuda/Kernels/Pool.py:10-112 */
int pz_maxpool2d_mask_fwd(const float* x, float* y, int32_t* mask, int64_t planes, int H, int W, int OH, int OW,
						  int fh, int fw, int sh, int sw, int ph, int pw, void* stream);
int pz_maxpool2d_mask_bwd(const float* dy, const int32_t* mask, float* dx, int64_t planes, int H, int W, int OH,
						  int OW, int fh, int fw, int sh, int sw, int ph, int pw, void* stream);
/* y must be zeroed by the caller (reference uses GPUArray.zeros, Pool.py:178) */
int pz_maxunpool2d_fwd(const float* x, const int32_t* mask, float* y, int64_t planes, int inHW, int outHW,
					   void* stream);
int pz_maxunpool2d_bwd(const float* dy, const int32_t* mask, float* dx, int64_t planes, int inHW, int outHW,
					   void* stream);

/* ---------------------------------------------------------------- softmax
 * replaces cudnnSoftmaxForward/Backward, SOFTMAX_ACCURATE (CuDnn.c:984,1063). Tensor [N][C][S]:
 * spatial mode normalises over C for each (n, s); per-activation mode over C*S for each n. */
int pz_softmax_fwd(int dtype, int mode, const void* x, void* y, int64_t N, int64_t C, int64_t S, void* stream);
int pz_softmax_bwd(int dtype, int mode, const void* y, const void* dy, void* dx, int64_t N, int64_t C, int64_t S,
				   void* stream);

/* ---------------------------------------------------------------- GEMM (tcgen05 / TMEM)
 * replaces cublasGemmEx as called by CuBlas_Context_gemm (CuBlas.c:327-403): row-major
 * C[M][N] = alpha * op(A) * op(B) + beta * C, fp32 storage computed as TF32 with fp32 accumulation in TMEM
 * (the reference sets CUBLAS_GEMM_DEFAULT_TENSOR_OP, CuBlas.c:91-106).  op(A) is M x K, op(B) is K x N;
 * lda/ldb/ldc are the row pitches (elements) of the STORED matrices.  bias (optional, length N) is added
 * in the epilogue (Linear.py:36-40 addVecToMat folded in). */
int pz_gemm(int dtype, const void* A, const void* B, void* C, int64_t M, int64_t N, int64_t K, int64_t lda,
			int64_t ldb, int64_t ldc, int transA, int transB, float alpha, float beta, const void* bias,
			void* stream);

/* ---------------------------------------------------------------- convolution (implicit GEMM on tcgen05)
 * replaces cudnnConvolutionForward / BackwardData / BackwardFilter / BackwardBias
 * (CuDnn.c:397-449,517-571,652-712,375-394).  Cross-correlation, NCHW, filter [K][C/G][R][S].
 * P,Q: output spatial size.  All three passes accept stride / pad / dilation / groups. */
typedef struct pz_conv2d_desc {
	int N, C, H, W;        /* input  (data)  tensor */
	int K, R, S;           /* filter: K output maps, R x S taps, C/groups input maps per filter */
	int P, Q;              /* output (grad)  spatial size */
	int stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, groups;
} pz_conv2d_desc;

/* y = conv(x, w) (+ bias[K]) */
int pz_conv2d_fprop(int dtype, const pz_conv2d_desc* d, const void* x, const void* w, const void* bias, void* y,
					void* stream);
/* dx = conv_transpose(dy, w) (+ bias[C], used when this is a Deconv forward, Dnn.py:211-215).
 * workspace: pz_conv2d_dgrad_workspace() bytes of device scratch (re-packed filter). */
size_t pz_conv2d_dgrad_workspace(int dtype, const pz_conv2d_desc* d);
int pz_conv2d_dgrad(int dtype, const pz_conv2d_desc* d, const void* dy, const void* w, const void* bias, void* dx,
					void* workspace, size_t workspace_bytes, void* stream);
/* dw = alpha * sum_{n,p,q} x (*) dy + beta * dw   (in place; CuDnn.c:682-685 scale/momentum) */
int pz_conv2d_wgrad(int dtype, const pz_conv2d_desc* d, const void* x, const void* dy, void* dw, float alpha,
					float beta, void* stream);
/* db[c] = alpha * sum_{n,s} t[n][c][s] + beta * db[c]   (cudnnConvolutionBackwardBias, CuDnn.c:375-394) */
int pz_bias_grad(int dtype, const void* t, void* db, int64_t N, int64_t C, int64_t S, float alpha, float beta,
				 void* stream);

/* ---- training closure either side of the path
 * pz_cross_entropy: Cuda/Kernels/Costs.py:77-106,133-157,213-247 (CostModule.crossEntropy after its softmax): probs
 *   (samples, cases, spatial) fp32 row-major, labels (samples, spatial) int32, weights (cases,) fp32 or NULL;
 *   grad = w_c * ((c == label) - p) / samples, *error += sum(-w_label * log(p_label)) / spatial  (error is NOT cleared here).
 * pz_count_mismatch: Costs.py:178-182 (calcAccuracy): *out += number of i with x[i] != y[i] (int32 inputs, fp32 count).
 * pz_sgd_nesterov: ElementWise.py:815-857; pz_adam: ElementWise.py:709-755 (mg / ms are fp32 whatever the parameter dtype). */
int pz_cross_entropy(const void* probs, const void* labels, const void* weights, int64_t samples, int64_t cases, int64_t spatial,
					 void* error, void* grad, void* stream);
int pz_count_mismatch(const void* x, const void* y, int64_t n, void* out, void* stream);
/* axis permutation of a contiguous tensor (Cuda/Source/Libs/CuDnnMemory.c transpose / moveaxis / swapaxes):
 * out (contiguous, extents out_shape[0..ndim)) [i0..] = in[sum_d i_d * in_stride[d]] (strides in elements), ndim <= 8 */
int pz_permute(int itemsize, void* out, const void* in, int ndim, const int64_t* out_shape, const int64_t* in_stride, void* stream);
/* random fills (Cuda/Source/Libs/CuRand.c:119-230): kind 0 = n raw uint32, 1 = n floats uniform in (a, b], 2 = n floats
 * normal(mean a, stddev b).  Counter-based Philox4x32-10: the values are a function of (seed, offset, index) only; the caller
 * advances offset by ceil(n / 4) per fill.  Not bit-compatible with cuRAND's XORWOW.
 * pz_dropout (ElementWise.py:495-580): out = in * (rands[i / mapsize] < partition) / p; rands are uint32 for float32 data,
 * uint16 for half / bfloat16; mapsize 1 = dropoutKer, H*W = dropout2dKer */
/* grouped matrix-vector product (Cuda/Kernels/MatVec.py:93-124,311-343; Backend/Blas.py:82-86 mulTensorOnVecGroup):
 * mat [z][h][w]; on_rows: out[z][h] = beta*out + alpha*sum_w mat*vec[z][w]; else out[z][w] = beta*out + alpha*sum_h mat*vec[z][h] */
int pz_matvec(int dtype, void* out, const void* mat, const void* vec, int64_t z, int64_t h, int64_t w, int on_rows, float alpha,
			  float beta, void* stream);
/* vector reductions behind blas.dot / l1norm / l2norm (Cuda/Source/Libs/CuBlas.c: cublasSdot, cublasSasum, cublasSnrm2):
 * kind 0 = sum x*y, 1 = sum |x|, 2 = sum x*x; fp32 accumulation; *out (float, device) += result */
int pz_vec_reduce(int dtype, int kind, const void* x, const void* y, int64_t n, void* out, void* stream);
/* local response normalisation (Cuda/Source/Libs/CuDnnNorm.c:329-690; formulas Cuda/Wrappers/CuDnnNorm.py:185-268):
 * mode 0 = across maps (crossMapLRN), 1 = within a map (mapLRN, no means tensor); window [i - (n-1)/2, i + n - (n-1)/2) clipped.
 * pz_lrn_bwd needs `tmp`: N*C*H*W floats of scratch. */
int pz_lrn_fwd(int dtype, int mode, const void* x, void* y, int64_t N, int64_t C, int64_t H, int64_t W, int n, float alpha, float beta,
			   float K, void* stream);
int pz_lrn_bwd(int dtype, int mode, const void* x, const void* grad, void* dx, void* tmp, int64_t N, int64_t C, int64_t H, int64_t W,
			   int n, float alpha, float beta, float K, void* stream);
int pz_rng_fill(int kind, void* out, int64_t n, uint64_t seed, uint64_t offset, float a, float b, void* stream);
/* the same fill with the generator offset (uint64, units of 4 words) read from and advanced in DEVICE memory: a captured graph
 * that replays the fill draws new numbers each time */
int pz_rng_fill_dev(int kind, void* out, int64_t n, uint64_t seed, void* offset_dev, float a, float b, void* stream);
int pz_dropout(int dtype, void* out, const void* in, const void* rands, uint32_t partition, float p, int64_t n, int64_t mapsize,
			   void* stream);
/* costmod.svm (Cuda/Kernels/Costs.py:109-130,249-279): l1 / squared hinge over (samples, cases, spatial...) scores */
int pz_svm(int l2, const void* scores, const void* labels, int64_t samples, int64_t cases, int64_t spatial, void* error, void* grad,
		   void* stream);
/* getAccuracyKernel reductions (Costs.py:172-205): kind 0 calcBCEAccuracy, 1 klDivergence (also writes grad), 2 l1HingeAccuracy;
 * *out += the sum */
int pz_cost_reduce(int kind, const void* x, const void* y, void* grad, float gradnorm, int64_t n, void* out, void* stream);
int pz_reduce_minmax_i32(const void* in, int64_t n, int want_max, void* out, void* stream);
int pz_dropout_slice(int dtype, void* out, const void* in, const void* rands, uint32_t partition, float p, int64_t n,
					 int64_t mapsize, int64_t start, int64_t stop, int64_t step, void* stream);
int pz_sgd_nesterov(int dtype, void* param, const void* grad, void* mom, float learn_rate, float mom_rate, int64_t n, void* stream);
int pz_adam(int dtype, void* param, const void* grad, void* mg, void* ms, float learn_rate, float fix1, float fix2, float epsilon,
			int64_t n, void* stream);

/* ---- recurrent cells (Cuda/Source/Libs/CuDnnRnn.c:565-1000 cudnnRNNForwardTraining / BackwardData cell math; the matrix
   products are pz_gemm calls).  fp32; gate order in a 4H row: i, f, c, o (Cuda/Backend.py:264-306 linear layers 0..3) */
int pz_lstm_cell_fwd(float* gates, const float* bw, const float* br, const float* c_prev, float* c_out, float* h_out, int64_t B,
                     int64_t H, void* stream);
int pz_lstm_cell_bwd(const float* dy, const float* dh_next, float* dc_io, const float* acts, const float* c, const float* c_prev,
                     float* dgates, int64_t B, int64_t H, int first, void* stream);
/* GRU, gate order r, i, h (Cuda/Backend.py:309-350 linear layers 0..2; formulas pinned by CuDnnRnn.py:303-352) */
int pz_gru_cell_fwd(float* gx, float* gh, const float* bw, const float* br, const float* h_prev, float* h_out, int64_t B, int64_t H,
                    void* stream);
int pz_gru_cell_bwd(const float* dy, const float* dh_next, const float* acts, const float* gh, const float* h_prev, float* dgx,
                    float* dgh, float* dh_carry, int64_t B, int64_t H, void* stream);
int pz_rnn_cell_fwd(float* h, const float* bw, const float* br, int64_t B, int64_t H, int mode, void* stream);
int pz_rnn_cell_bwd(const float* dy, const float* dh_next, const float* h, float* dpre, int64_t n, int mode, void* stream);

/* ---------------------------------------------------------------- data-parallel gradient sync (NCCL over NVLink)
 * replaces Grid.py's CUDA-IPC parent/child star (Grid.py:66-157; Buffer.c:411-424) */
int pz_nccl_version(int* version);                                   /* ncclGetVersion of the loaded libnccl */
int pz_nccl_unique_id(void* id128);                                  /* 128-byte ncclUniqueId, host memory */
int pz_nccl_comm_init(void** comm, int nranks, int rank, const void* id128);
int pz_nccl_comm_destroy(void* comm);
/* in-place sum all-reduce then scale by `scale` (1/P for Grid.sumTensor's mean, Grid.py:126-133) */
int pz_nccl_allreduce_mean(void* comm, int dtype, void* buf, int64_t count, float scale, void* stream);
/* the mean over the ranks of `nseg` disjoint segments of a buffer as ONE grouped launch on `stream` (ncclAvg): a bucket of the
 * overlapped gradient synchronisation (puzzlelib_b200/grid.py GradientSync; replaces the star reduce of Grid.py:123-157) */
int pz_nccl_allreduce_avg_segments(void* comm, int dtype, void* const* ptrs, const int64_t* counts, int nseg, void* stream);
int pz_nccl_broadcast(void* comm, int dtype, void* buf, int64_t count, int root, void* stream); /* Grid.py:114-121 */
/* fused: all-reduce(sum) the flat gradient, then mom = mr*mom + lr*(grad/P); param += mom in one pass
 * (Optimizer.py:166-170 + MomentumSGD.py:24-27) */
/* the local half of the fused call: grad *= scale; mom = mr*mom + lr*grad; param += mom (single pass) */
int pz_mean_sgd_momentum(int dtype, void* param, void* grad, void* mom, int64_t count, float scale, float learn_rate,
						 float mom_rate, void* stream);
int pz_nccl_allreduce_sgd_momentum(void* comm, int dtype, void* param, void* grad, void* mom, int64_t count,
								   float scale, float learn_rate, float mom_rate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PZB200_H */
