/*
 * vitk.h -- C-ABI of libvitk.so: hand-written sm_100a CUDA kernels for the ViT-B/16 PAD hot path.
 *
 * The reference (ArchitRastogi20/vit-spoof-detection-pda) has no FFI layer: its hot path is a
 * torch.nn.Module that bottoms out in timm -> torch -> cuDNN/cuBLAS/SDPA/ATen library kernels.
 * Every entry point below replaces one of those library-call sites (K1..K12 in SURVEY.md 2.1);
 * the reference interface each one stands in for is cited as /root/reference file:line.
 *
 * Conventions
 *   - extern "C", POD arguments only: device pointers, integer sizes, float scalars, the CUDA
 *     stream as void* (cudaStream_t).  No torch types.  The library allocates nothing persistent:
 *     the caller owns every buffer (parameters, gradients, activations, workspace).
 *   - every entry returns int: 0 = ok, non-zero = cudaError_t or VITK_ERR_*;
 *     vitk_last_error_string() describes the last failure on the calling thread.
 *   - all work is enqueued on the caller's stream; no hidden synchronisation.
 *   - "rows" M = batch * 197 tokens; DIM 768, 12 heads x 64, MLP 3072 are compile-time constants
 *     of the kernels (the reference fixes model_name = vit_base_patch16_224, train_advanced.py:33).
 *   - dtype arguments: VITK_F32 / VITK_BF16 describe the storage type of activation buffers.
 *     precision VITK_PREC_FP32_VALIDATE = fp32 storage + fp32 FFMA arithmetic everywhere (the 1e-4
 *     parity mode); VITK_PREC_BF16 = bf16 GEMM/attention operands on tcgen05 / mma tensor cores,
 *     fp32 accumulation, fp32 residual stream, LayerNorm/softmax/loss/Adam in fp32
 *     (mirrors the reference's autocast placement).
 */
#ifndef VITK_H_
#define VITK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VITK_VERSION 101

/* model constants (timm vit_base_patch16_224, train_advanced.py:33,190) */
#define VITK_IMG 224
#define VITK_PATCH 16
#define VITK_NPATCH 196
#define VITK_NTOK 197
#define VITK_DIM 768
#define VITK_HEADS 12
#define VITK_HEAD_DIM 64
#define VITK_MLP 3072
#define VITK_HEAD_HIDDEN 512

enum { VITK_OK = 0, VITK_ERR_ARG = 10001, VITK_ERR_UNSUPPORTED = 10002, VITK_ERR_DRIVER = 10003 };
enum { VITK_F32 = 0, VITK_BF16 = 1 };
enum { VITK_PREC_FP32_VALIDATE = 0, VITK_PREC_BF16 = 1 };

/* epilogues of the Linear forward (replaces cuBLASLt + ATen add/gelu call sites K4,K6,K7,K8) */
enum {
  VITK_EPI_BIAS = 0,          /* y = x W^T + b                                 (out: act dtype)          */
  VITK_EPI_BIAS_GELU = 1,     /* u = x W^T + b ; y = gelu_erf(u) ; aux = gelu_erf'(u)   (both act dtype)   */
  VITK_EPI_BIAS_RESIDUAL = 2, /* y = res + x W^T + b                           (res, out: fp32)          */
  VITK_EPI_QKV_SCATTER = 3    /* y = x W^T + b written head-major [36][M][64]  (out: act dtype)          */
};
/* matrix storage of an activation operand */
enum {
  VITK_LAYOUT_ROWMAJOR = 0,   /* [M][C] */
  VITK_LAYOUT_HEADMAJOR = 1   /* [C/64][M][64]  (q,k,v and their gradients) */
};
/* GEMM engine: the tcgen05/TMEM kernel, or the SIMT FFMA kernel (always used by FP32_VALIDATE) */
enum { VITK_ENGINE_AUTO = 0, VITK_ENGINE_SIMT = 1, VITK_ENGINE_TCGEN05 = 2 };

int vitk_version(void);
const char* vitk_last_error_string(void);
/* number of SMs / compute capability of the current device (host query; 0 on failure) */
int vitk_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* 1 when this library is the development build (libvitk_dev.so: device-side tracer + timing experiments), else 0 */
int vitk_is_dev_build(void);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm over the last dim (768).  Replaces ATen native_layer_norm reached from timm
 * Block.norm1/.norm2, vit.norm (eps 1e-6) and classifier[0] (eps 1e-5, train_advanced.py:194).
 *   x: fp32 [rows] with row stride x_stride (elements); y: y_dtype, dense [rows][768];
 *   mean/rstd: fp32 [rows] saved for backward (may be NULL in eval).
 * bwd: dx = (dres ? dres : 0) + LN'(dy) ; also written as bf16 to dx16 when non-NULL (dres may alias dx);
 *   dgamma/dbeta are ACCUMULATED (+=) into fp32 [768];
 *   dx_colsum (optional) += column sums of the dx written: dx is the gradient of the residual stream, i.e. the
 *   dY of the Linear whose output was added into it, so this IS that Linear's bias gradient (proj / fc2) --
 *   fused here instead of re-reading dY in a separate reduction.
 * ------------------------------------------------------------------------------------------- */
int vitk_layernorm_fwd(const float* x, int64_t x_stride, const float* gamma, const float* beta,
                       void* y, int y_dtype, float* mean, float* rstd, int rows, float eps, void* stream);
int vitk_layernorm_bwd(const void* dy, int dy_dtype, const float* x, int64_t x_stride,
                       const float* gamma, const float* mean, const float* rstd, const float* dres,
                       float* dx, void* dx16, float* dgamma, float* dbeta, float* dx_colsum,
                       int rows, void* stream);

/* ---------------------------------------------------------------------------------------------
 * nn.Linear forward / backward (timm Attention.qkv, Attention.proj, Mlp.fc1, Mlp.fc2).
 *   fwd   : Y[M,N]  = X[M,K] W[N,K]^T + b, epilogue as above.   `aux` = gelu'(u) output for BIAS_GELU (saved
 *           for backward instead of the pre-activation; may be NULL in eval), residual (fp32) for BIAS_RESIDUAL.
 *   dgrad : dX[M,K] = dY[M,N] W[N,K]   ; if gelu_grad != NULL: dX *= gelu_grad  (the aux of the forward; [M,K])
 *   wgrad : dW[N,K] += dY[M,N]^T X[M,K]; db[N] += colsum(dY)   (dW, db fp32; += so autograd-style
 *           accumulation and split-K share one code path; caller zeroes the gradient buffer)
 * X/dY/W are `dtype` (VITK_F32 for FP32_VALIDATE, VITK_BF16 otherwise; W then is the bf16 shadow).
 * ------------------------------------------------------------------------------------------- */
int vitk_linear_fwd(const void* x, int x_layout, const void* w, const float* bias, void* y, void* aux,
                    int M, int N, int K, int epilogue, int dtype, int engine, void* stream);
/* dx_colsum (optional): fp32 [K] += column sums of dX, fused into this GEMM's epilogue instead of a separate pass over
 * dX.  With gelu_grad it is the bias gradient of the Linear in front of the GELU (fc1).  Without, the model driver uses
 * it on the proj dgrad: with attention dropout 0 the column sums of the attention-output gradient ARE the v section of
 * the qkv bias gradient (softmax rows sum to one), see csrc/attention.cu. */
int vitk_linear_dgrad(const void* dy, int dy_layout, const void* w, void* dx, const void* gelu_grad,
                      float* dx_colsum, int M, int N, int K, int dtype, int engine, void* stream);
int vitk_linear_wgrad(const void* dy, int dy_layout, const void* x, float* dw, float* db,
                      int M, int N, int K, int dtype, int engine, void* stream);
/* The same two calls with a scratch of `scratch_floats` 4-byte elements (all zero on entry; handed back all zero;
 * vitk_gemm_tail_scratch_floats(N of the output) always suffices) that permits the SPLIT TAIL of the tcgen05 engine: when
 * the output tiles leave the last wave of the persistent grid at most a quarter full and the reduction is >= 24 k-blocks deep
 * (bs 64: fc2 forward, fc1 dgrad, qkv dgrad -- 150 tiles on 74 CTA pairs = two waves + a wave of two tiles), the tiles of
 * the full waves stay whole-K items with the fused epilogue and the k-blocks of the remaining tiles are dealt out to all
 * CTA pairs of the SAME launch; partial accumulators meet in the scratch (red.global.add.v4.f32) and the last CTA to arrive
 * at a region's ticket applies the epilogue (csrc/gemm_tc.cu: tc_tail_plan, "partial item"; plan query: vitk_gemm_tail_plan).
 * The tail tiles' sums are rounded like the others but their fp32 summation order is not reproducible run to run --
 * vitk_model_fwd / _bwd_stage pass a scratch in training mode only.  scratch == NULL is the plain call. */
int vitk_linear_fwd_ws(const void* x, int x_layout, const void* w, const float* bias, void* y, void* aux,
                       int M, int N, int K, int epilogue, int dtype, int engine, float* scratch, size_t scratch_floats,
                       void* stream);
int vitk_linear_dgrad_ws(const void* dy, int dy_layout, const void* w, void* dx, const void* gelu_grad,
                         float* dx_colsum, int M, int N, int K, int dtype, int engine, float* scratch,
                         size_t scratch_floats, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Patch embedding (timm PatchEmbed.proj = Conv2d(3,768,k16,s16) + _pos_embed; K1,K2) as an im2col-free GEMM: the patch
 * matrix is never materialised in global memory.  VITK_PREC_BF16 (tcgen05 engine): TMA boxes of a 5-D tensor map over the
 * NCHW fp32 image -- or over the uint8 HWC image -- land in shared memory, converter warps write the bf16 operand tile
 * (ToTensor + Normalize first on the uint8 edge), tcgen05.mma consumes it against the bf16 weight shadow.
 * VITK_PREC_FP32_VALIDATE / VITK_ENGINE_SIMT: the FFMA GEMM reads the image in place.
 *   fwd  : x0[b,1+p,:] = patch(b,p) . Wpe^T + bpe + pos[1+p] ; x0[b,0,:] = cls + pos[0]   (fp32 out, [B*197][768])
 *          wpe = fp32 [768][3*16*16] (vit.patch_embed.proj.weight); wpe16 = its bf16 shadow (tcgen05 path; NULL otherwise)
 *   wgrad: dWpe += dx0^T patches ; dbpe += colsum(dx0[:,1:]) ; dpos += sum_b dx0 ; dcls += sum_b dx0[b,0]
 *          dx0 = fp32 gradient of x0 ([B*197][768]); dx0_bf16 = the same in bf16 (tcgen05 path; NULL otherwise);
 *          images = the fp32 NCHW input; images need no gradient.
 * ------------------------------------------------------------------------------------------- */
int vitk_patch_embed_fwd(const float* images, const float* wpe, const void* wpe16, const float* bpe, const float* cls,
                         const float* pos, float* x0, int batch, int precision, int engine, void* stream);
/* same forward from uint8 HWC pixels [B][224][224][3]: ToTensor (x / 255) + Normalize ((x - mean[c]) / std[c]) of the
 * reference's transforms (train_advanced.py:174-175, 180-181; test.py get_test_transforms) fused into the patch loader:
 * 150 KB instead of 602 KB read per image, no fp32 NCHW tensor materialised on the bf16 path.  nchw_scratch (fp32
 * [B][3][224][224], may be NULL on the bf16 / tcgen05 path) is only written by the fp32-validate / SIMT path. */
int vitk_patch_embed_fwd_u8(const uint8_t* images_hwc, const float* mean3, const float* std3, const float* wpe,
                            const void* wpe16, const float* bpe, const float* cls, const float* pos, float* x0,
                            float* nchw_scratch, int batch, int precision, int engine, void* stream);
/* ToTensor + Normalize alone: uint8 HWC -> fp32 NCHW (the weight gradient of a uint8-fed training step reads this) */
int vitk_u8_to_nchw(const uint8_t* images_hwc, const float* mean3, const float* std3, float* out, int batch, void* stream);
int vitk_patch_embed_wgrad(const float* dx0, const void* dx0_bf16, const float* images, float* dwpe, float* dbpe,
                           float* dcls, float* dpos, int batch, int precision, int engine, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-head self-attention core, N=197, d=64, 12 heads, scale 1/8, no mask, no dropout
 * (timm Attention -> F.scaled_dot_product_attention; K5).
 *   qkv : head-major [36][M][64]  (q heads 0..11, k heads 12..23, v heads 24..35)
 *   out : row-major [M][768] (heads concatenated), lse: fp32 [12][M] (log-sum-exp of scaled scores)
 *   bwd : dqkv head-major [36][M][64] from dout [M][768]
 * ------------------------------------------------------------------------------------------- */
int vitk_attn_fwd(const void* qkv, void* out, float* lse, int batch, int dtype, void* stream);
/* dqkv_colsum (optional): fp32 [2304] += column sums of dqkv (= the qkv bias gradient), fused into the kernel */
int vitk_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                  float* dqkv_colsum, int batch, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Classifier head on the CLS feature (train_advanced.py:193-200, 204) -- fp32 always.
 *   feat [B][768] (CLS rows of vit.norm output, dense) -> logits [B][C]
 *   LN(1e-5) -> *mask1 -> Linear(768,512) -> GELU -> *mask2 -> Linear(512,C)
 *   mask1 [B][768] / mask2 [B][512] are pre-scaled dropout masks (0 or 1/(1-p)); NULL = no dropout.
 *   save: caller buffer of vitk_head_save_floats(B) floats (activations for backward).
 *   bwd: dfeat [B][768]; parameter grads ACCUMULATED (+=).
 * ------------------------------------------------------------------------------------------- */
size_t vitk_head_save_floats(int batch);
int vitk_head_fwd(const float* feat, const float* ln_w, const float* ln_b, const float* w1,
                  const float* b1, const float* w2, const float* b2, const float* mask1,
                  const float* mask2, float* logits, float* save, int batch, int num_classes, void* stream);
int vitk_head_bwd(const float* dlogits, float* save, const float* ln_w, const float* w1,
                  const float* w2, const float* mask1, const float* mask2, float* dfeat, float* dln_w,
                  float* dln_b, float* dw1, float* db1, float* dw2, float* db2, int batch,
                  int num_classes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Focal loss + gradient (FocalLoss.forward, train_advanced.py:98-107) fused with the eval
 * post-processing (softmax P(live)=probs[:,1], argmax: train_advanced.py:342,387-394; test.py:212-217).
 *   alpha: fp32 [C] per-class weights (scalar alpha = all entries equal; class weights:
 *   train_advanced.py:521-529).  reduction: 0 = mean, 1 = sum, 2 = none.
 *   loss_per_sample [B] (always written); loss_out [1] (mean or sum); dlogits [B][C] = d loss / d logits
 *   scaled by grad_scale (use 1/world for data parallel); probs1 [B] / preds int64 [B] / ncorrect int32[1]
 *   optional (NULL to skip).  C <= 8.
 * ------------------------------------------------------------------------------------------- */
int vitk_focal_fwd_bwd(const float* logits, const int64_t* targets, const float* alpha, float gamma,
                       int reduction, float grad_scale, float* loss_per_sample, float* loss_out,
                       float* dlogits, float* probs1, int64_t* preds, int* ncorrect, int batch,
                       int num_classes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Evaluation post-processing on the device: replaces the per-batch device->host copy of scores and the
 * numpy/sklearn loop of find_optimal_threshold (train_advanced.py:239-275: np.linspace(0.3, 0.7, 41), preds =
 * probs >= thresh, accuracy / precision / recall / F1 per threshold) and the confusion counts of calculate_metrics
 * (test.py:241-243).
 *   vitk_threshold_hist  : hist[(label == 1)][k] += 1 with k = |{s : thresholds[s] <= probs[i]}| (float64 comparison, as
 *                          numpy does for a float32 array against a float64 scalar); thresholds ascending, steps <= 255;
 *                          hist = uint64 [2][steps + 1], ACCUMULATED across calls (caller zeroes it once per epoch)
 *   vitk_threshold_counts: counts[s] = (tp, fp, tn, fn) at thresholds[s] (live = 1 is the positive class), int64 [steps][4]
 * ------------------------------------------------------------------------------------------- */
int vitk_threshold_hist(const float* probs, const int64_t* labels, const double* thresholds, int n, int steps,
                        unsigned long long* hist, void* stream);
int vitk_threshold_counts(const unsigned long long* hist, int steps, long long* counts, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused multi-tensor Adam / AdamW over one flat fp32 parameter buffer
 * (torch.optim.AdamW at train_advanced.py:592-597, Adam-L2 per README.md:140-147,
 *  clip_grad_norm_ at :334, GradScaler.unscale_ at :333).
 *   pass 1: vitk_grad_sumsq  -> sumsq[0] = sum(g^2) over the flat grads (deterministic 2-stage)
 *           (`partial` = caller scratch of vitk_grad_sumsq_scratch_floats() floats)
 *   pass 2: vitk_adam_step   -> g' = g * grad_mult * clipcoef, clipcoef = min(1, max_norm /
 *           (sqrt(sumsq)*grad_mult + 1e-6)) when sumsq != NULL and max_norm > 0 (torch semantics);
 *           mode 0 (Adam-L2): g' += wd*p before the moments; mode 1 (AdamW): p *= 1 - lr*wd first.
 *           Writes p, m, v and (when p16 != NULL) the bf16 shadow of p in the same pass.
 *           step is the 1-based step count (bias corrections 1-beta^step).
 * ------------------------------------------------------------------------------------------- */
size_t vitk_grad_sumsq_scratch_floats(void);
int vitk_grad_sumsq(const float* g, size_t n, float* partial, float* sumsq, void* stream);
int vitk_adam_step(float* p, const float* g, float* m, float* v, void* p16, size_t n, double lr,
                   double beta1, double beta2, double eps, double weight_decay, int mode, int step,
                   float grad_mult, const float* sumsq, float max_norm, void* stream);
/* The same step for a CUDA-GRAPH-CAPTURED training loop: the step count and the learning rate are read from device memory
 * (step_dev is incremented by the launch itself; lr_dev is written by the host between replays), the bias corrections are
 * computed on the device into hyper (3 floats of caller scratch). */
int vitk_adam_step_graph(float* p, const float* g, float* m, float* v, void* p16, size_t n, const float* lr_dev,
                         int* step_dev, float* hyper, double beta1, double beta2, double eps, double weight_decay, int mode,
                         float grad_mult, const float* sumsq, float max_norm, void* stream);
/* g *= grad_mult * clipcoef in place, clipcoef as in vitk_adam_step (1 when sumsq == NULL or max_norm <= 0): the eager
 * form of GradScaler.unscale_ (train_advanced.py:333) and clip_grad_norm_ (:334) for callers that pair them with stock
 * torch pieces instead of the fused Adam pass */
int vitk_grad_scale(float* g, size_t n, float grad_mult, const float* sumsq, float max_norm, void* stream);
/* ---------------------------------------------------------------------------------------------
 * Data-parallel optimizer step fused with its collectives over NVLink / NVSwitch multicast (csrc/dp_nvls.cu; new
 * functionality for BASELINE configs[4], the reference is single-GPU).  Rank r owns `n` elements (its slice) of the flat
 * buffers; `*_local` point at the slice in this rank's memory, `*_mc` at the same offset behind the MULTICAST address of the
 * symmetric allocation (torch.distributed._symmetric_memory provides both).  The caller orders the ranks with three
 * stream-ordered barriers: gradients final -> reduce_sumsq -> norm table complete -> adam_bcast -> parameters complete.
 *   vitk_nvls_reduce_sumsq: g_local[i] = scale * sum_ranks g[i] (multimem.ld_reduce through the switch), and the slice's sum
 *       of squares to entry `slot` of every rank's norm table (table_mc: multicast address of a float table with one entry per
 *       (domain, rank)); partial = vitk_nvls_scratch_floats() floats of local scratch; max_ctas > 0 bounds the grid (a
 *       reduce that runs on a side stream under the backward pass).
 *   vitk_nvls_adam_bcast:   Adam / AdamW on the slice exactly as vitk_adam_step (clip coefficient from the sum of the `world`
 *       table entries, max_norm <= 0 or sq_table == NULL: no clip), updated fp32 masters to p_mc and their bf16 shadow to
 *       p16_mc (may be NULL): every rank's copy of the slice is written by the switch.  m, v: the slice's moments (local).
 *       local_only_ranges (host array of n_ranges <= 64 [lo, hi) element pairs relative to the slice, multiples of 4): fp32
 *       masters in these ranges are stored to p_local only -- the tensor-core GEMM weights, which every rank reads through
 *       the bf16 shadow; vitk_nvls_bcast_f32 (a plain multicast copy with at most max_ctas CTAs, meant for a side stream under
 *       the next forward pass) brings the other ranks' fp32 copies up to date.
 * ------------------------------------------------------------------------------------------- */
size_t vitk_nvls_scratch_floats(void);
int vitk_nvls_reduce_sumsq(float* g_local, const float* g_mc, size_t n, float scale, float* partial, float* table_mc, int slot,
                           int max_ctas, void* stream);
int vitk_nvls_adam_bcast(float* p_local, float* p_mc, void* p16_mc, const float* g_local, float* m, float* v, size_t n,
                         double lr, double beta1, double beta2, double eps, double weight_decay, int mode, int step,
                         float grad_mult, const float* sq_table, int world, float max_norm,
                         const int64_t* local_only_ranges, int n_ranges, void* stream);
int vitk_nvls_bcast_f32(const float* src, float* dst_mc, size_t n, int max_ctas, void* stream);
/* p16 = bf16(p) over a flat buffer (after load_state_dict / external optimizers) */
int vitk_cast_f32_to_bf16(const float* src, void* dst, size_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Whole-model drivers: one call enqueues the full encoder+head forward (a1,a2,a3 of SURVEY.md 8a)
 * or one backward stage, so Python makes O(depth) ctypes calls and the sequence is CUDA-graph
 * capturable.  Flat parameter layout = state_dict order, each tensor padded to a
 * multiple of 64 elements: query it with vitk_param_layout().
 * ------------------------------------------------------------------------------------------- */
typedef struct vitk_model {
  int32_t batch, depth, num_classes, precision; /* VITK_PREC_*                                        */
  int32_t training;                             /* 1: save activations for backward                   */
  int32_t engine;                               /* VITK_ENGINE_*                                      */
  const float* params;                          /* flat fp32 master parameters                        */
  const void* params16;                         /* flat bf16 shadow (same offsets); NULL in FP32 mode */
  float* grads;                                 /* flat fp32 gradients (same offsets), accumulated    */
  void* workspace;                              /* vitk_workspace_bytes() bytes, 256-B aligned        */
  const float* images;                          /* [B][3][224][224] fp32 NCHW                         */
  float* logits;                                /* [B][C] fp32                                        */
  const float* mask1;                           /* optional dropout masks of the head (see head_fwd)  */
  const float* mask2;
  const float* dlogits;                         /* [B][C] fp32, input of backward                     */
  int32_t frozen_backbone;                      /* 1: backward stops after the head (config 4)        */
  int32_t sm_budget;                            /* > 0: SMs the persistent kernels of THIS call may occupy (the rest is
                                                   left to a collective in flight, see dp.py); 0 = all                 */
  /* input edge: when images_u8 != NULL the patch loader reads uint8 HWC pixels [B][224][224][3]
   * and applies ToTensor + Normalize (train_advanced.py:174-175 / 180-181: x/255, (x - mean[c]) / std[c], fp32, in that
   * order) itself; `images` is then ignored */
  const uint8_t* images_u8;
  float norm_mean[3], norm_std[3];
  int32_t flags;                                /* VITK_FLAG_* */
  int32_t reserved;
} vitk_model;
/* keep the weight-gradient GEMMs on the caller's stream instead of the library's side stream (bench.py's per-launch
 * timing: kernels must not overlap while they are bracketed by events) */
#define VITK_FLAG_WGRAD_INLINE 1

/* n_tensors = 4 + 12*depth + 8; offsets/sizes in elements, state_dict order. returns total elements */
int64_t vitk_param_layout(int depth, int num_classes, int64_t* offsets, int64_t* sizes, int max_tensors);
size_t vitk_workspace_bytes(int batch, int depth, int precision, int training);
int vitk_model_fwd(const vitk_model* m, void* stream);
/* backward stages in the order gradients become final (DP buckets are all-reduced in between):
 *   stage 0            : head + final norm              -> grads of classifier.*, vit.norm.*
 *   stage 1 + (depth-1-i): block i (i = depth-1 .. 0)   -> grads of vit.blocks.i.*
 *   stage depth+1      : patch embedding, cls, pos      -> grads of vit.patch_embed.*, cls, pos    */
int vitk_model_bwd_stage(const vitk_model* m, int stage, void* stream);
int vitk_model_num_bwd_stages(int depth);

/* Tuning overrides for tests and A/B timing -- process-wide, 0 = default everywhere, every one of them yields valid
 * results: key 1 = whole-K tiles for accumulate (wgrad) GEMMs, 2 = force BLOCK_N (128/192/256), 4 = CTA group (1 single
 * CTAs, 2 pairs), 5 = per-thread epilogue IO, 6 = no programmatic dependent launch, 13 = 1 stream-K instead of sliced
 * split-K (> 1: fill threshold in percent).  The development build (libvitk_dev.so) adds: 0 = swap LBO/SBO of MN-major
 * operands, 7 = timing-only bit mask (skip epilogue body / operand loads / ...: RESULTS INVALID), 12 = whole qkv bias
 * gradient from the attention kernel.  The release library rejects those keys: it contains no result-invalidating path. */
int vitk_debug_set(int key, int value);
/* SM budget of the CALLING THREAD for stand-alone kernel calls and the plan queries below (thread-local; 0 = all SMs;
 * returns the previous value).  The whole-model drivers take theirs per call from vitk_model.sm_budget. */
int vitk_set_sm_budget(int n);
/* Device-side tracer (development build only; VITK_ERR_UNSUPPORTED otherwise): thread 0 of every CTA of every libvitk
 * kernel appends (globaltimer ns, kernel id << 48 | phase << 40 | SM id << 24 | block index, aux) -- three 64-bit words
 * -- to `buf` (device memory, `bytes` bytes: word 0 = number of marks, word 1 = capacity, marks from word 2).  Phase 0 = CTA
 * started, 1 = stream dependencies satisfied, 2 = CTA finished.  tools/step_timeline.py turns that into a per-kernel
 * timeline of a real training step (side stream and programmatic dependent launch left on). */
int vitk_trace_start(void* buf, size_t bytes);
int vitk_trace_stop(void);
/* Host-only view of the tcgen05 GEMM's work decomposition for C[I][J] += over R (no launch; only the SM count / budget is
 * consulted): tile width BLOCK_N, CTA group (1 single CTAs with 128-row tiles, 2 CTA pairs with 256-row tiles), mode
 * (0 whole-K tiles strided over the persistent clusters, 1 contiguous stream-K ranges, 2 sliced split-K: one k-slice of
 * one tile per cluster, 3 split tail), number of clusters launched, tile grid and number of 64-deep k-blocks.  accumulate == 1
 * is the weight-gradient epilogue (fp32 +=), 2 / 3 a non-accumulating GEMM with a tail scratch and bf16 / fp32 output; b_mn_major != 0 says the B operand is stored with its row index contiguous
 * (dgrad's W, wgrad's X).  vitk_gemm_plan_items writes the (tile, first k-block, end k-block) triples cluster `cluster`
 * walks -- the same iterator code the kernel's warps run -- and returns their count (negative VITK_ERR_* on bad
 * arguments); tests/test_host_logic.py checks that all clusters together cover every (tile, k-block) exactly once. */
int vitk_gemm_plan(int I, int J, int R, int accumulate, int b_mn_major, int* block_n, int* cta_group, int* mode,
                   int* n_clusters, int* n_tiles_m, int* n_tiles_n, int* kb_total);
int vitk_gemm_plan_items(int I, int J, int R, int accumulate, int b_mn_major, int cluster, int* items, int max_items);
/* Split tail of a non-accumulating GEMM (vitk_linear_fwd_ws / _dgrad_ws) given vitk_gemm_tail_scratch_floats(J) of scratch:
 * *n_whole tiles keep whole-K items with the fused epilogue, the k-blocks of the last *n_tail tiles (0: plain launch) are
 * dealt out to all clusters, *tail_row0 = first matrix row of the first tail tile; f32_out != 0: the epilogue writes fp32
 * (bias + residual), else bf16.  vitk_gemm_plan / _plan_items with accumulate == 2 (bf16 out) / 3 (fp32 out) describe the
 * same launch (mode 3) item by item.  Host-only. */
int vitk_gemm_tail_plan(int I, int J, int R, int b_mn_major, int f32_out, int* n_whole, int* n_tail, int* tail_row0);
size_t vitk_gemm_tail_scratch_floats(int J);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long vitk_launch_count(void);
/* per-launch GEMM timing with CUDA events on the launching stream (bench.py's live roofline):
 * enable(1) clears and starts recording, read() synchronises and returns up to `max` records:
 * ms[i] and info[5*i..5*i+4] = I, J, R (C[I][J] += over R), epilogue mode, engine. */
int vitk_prof_enable(int on);
int vitk_prof_read(float* ms, int* info, int max);

#ifdef __cplusplus
}
#endif
#endif /* VITK_H_ */
