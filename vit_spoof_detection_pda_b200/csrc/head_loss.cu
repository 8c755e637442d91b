// Classifier head (fwd/bwd) and the fused focal loss -- latency-bound, fp32 throughout.
//   head : /root/reference/train_advanced.py:193-200, 204  (LN(1e-5) -> Dropout -> Linear(768,512) -> GELU
//          -> Dropout -> Linear(512,C)); dropout enters as caller-provided pre-scaled masks.
//   focal: /root/reference/train_advanced.py:98-107 (alpha * (1-pt)^gamma * ce, mean/sum/none) with the
//          per-class alpha generalisation (class weights: train_advanced.py:521-529), fused with its
//          gradient and with softmax P(live) / argmax / accuracy (train_advanced.py:342-343, 387-394;
//          test.py:212-217).
#include "common.cuh"

namespace vitk {

constexpr int HD_IN = VITK_DIM;           // 768
constexpr int HD_HID = VITK_HEAD_HIDDEN;  // 512
constexpr int HD_THREADS = 512;
// per-sample save area (floats): xhat | y | z1 | dz1 | dym | rstd, ticket (+pad) | dy accumulator of the split backward
constexpr int HS_XHAT = 0, HS_Y = 768, HS_Z1 = 1536, HS_DZ1 = 2048, HS_DYM = 2560, HS_RSTD = 3328, HS_ACC = 3392, HS_TOTAL = 4160;
constexpr float HD_EPS = 1e-5f;

__device__ __forceinline__ float block_sum_512(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < HD_THREADS / 32; ++w) s += red[w];
  return s;
}

__global__ void __launch_bounds__(HD_THREADS)
head_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2,
                const float* __restrict__ b2, const float* __restrict__ mask1, const float* __restrict__ mask2,
                float* __restrict__ logits, float* __restrict__ save, int num_classes) {
  pdl_sync_traced(TK_HEAD);
  __shared__ __align__(16) float ys[HD_IN];
  __shared__ float as[HD_HID];
  __shared__ float red[HD_THREADS / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = feat + (int64_t)b * HD_IN;
  float* sv = save ? save + (int64_t)b * HS_TOTAL : nullptr;
  const float x0 = x[tid], x1 = (tid < HD_IN - HD_THREADS) ? x[tid + HD_THREADS] : 0.f;
  const float mean = block_sum_512(x0 + x1, red) * (1.0f / HD_IN);
  const float d0 = x0 - mean, d1 = (tid < HD_IN - HD_THREADS) ? x1 - mean : 0.f;
  const float var = block_sum_512(d0 * d0 + d1 * d1, red) * (1.0f / HD_IN);
  const float rstd = 1.0f / sqrtf(var + HD_EPS);
  for (int k = tid; k < HD_IN; k += HD_THREADS) {
    const float xh = (k == tid ? d0 : d1) * rstd;
    float y = xh * ln_w[k] + ln_b[k];
    if (mask1) y *= mask1[(int64_t)b * HD_IN + k];
    ys[k] = y;
    if (sv) { sv[HS_XHAT + k] = xh; sv[HS_Y + k] = y; }
  }
  if (sv && tid == 0) { sv[HS_RSTD] = rstd; sv[HS_RSTD + 1] = 0.f; }   // +1: arrival ticket of the split backward
  if (sv)
    for (int k = tid; k < HD_IN; k += HD_THREADS) sv[HS_ACC + k] = 0.f;   // accumulated by the backward's 4 CTAs per sample
  __syncthreads();
  // 768 -> 512 mat-vec: every warp owns 32 consecutive hidden units; the LN output lives in registers (6 float4 per
  // lane) and two weight rows (12 independent 16-byte loads per lane) are in flight at a time -- the kernel is
  // latency-bound on the 1.5 MB weight read, not on arithmetic.
  {
    float4 yv[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) yv[i] = *reinterpret_cast<const float4*>(ys + (i * 32 + lane) * 4);
    for (int t = warp * 32; t < warp * 32 + 32; t += 2) {
      const float4* w0 = reinterpret_cast<const float4*>(w1 + (int64_t)t * HD_IN);
      const float4* w1r = reinterpret_cast<const float4*>(w1 + (int64_t)(t + 1) * HD_IN);
      float4 wa[6], wb[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) { wa[i] = __ldg(w0 + i * 32 + lane); wb[i] = __ldg(w1r + i * 32 + lane); }
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        a0 = fmaf(yv[i].x, wa[i].x, fmaf(yv[i].y, wa[i].y, fmaf(yv[i].z, wa[i].z, fmaf(yv[i].w, wa[i].w, a0))));
        a1 = fmaf(yv[i].x, wb[i].x, fmaf(yv[i].y, wb[i].y, fmaf(yv[i].z, wb[i].z, fmaf(yv[i].w, wb[i].w, a1))));
      }
      a0 = warp_sum(a0);
      a1 = warp_sum(a1);
      if (lane < 2) {
        const int tt = t + lane;
        const float z = (lane == 0 ? a0 : a1) + b1[tt];
        if (sv) sv[HS_Z1 + tt] = z;
        float g = gelu_erf(z);
        if (mask2) g *= mask2[(int64_t)b * HD_HID + tt];
        as[tt] = g;
      }
    }
  }
  __syncthreads();
  for (int c = warp; c < num_classes; c += HD_THREADS / 32) {
    const float* wr = w2 + (int64_t)c * HD_HID;
    float a = 0.f;
    for (int t = lane; t < HD_HID; t += 32) a = fmaf(as[t], wr[t], a);
    a = warp_sum(a);
    if (lane == 0) logits[(int64_t)b * num_classes + c] = a + b2[c];
  }
  trace_end(TK_HEAD);
}

// per-sample backward to the feature; leaves dz1 and dym in the save area for the parameter-grad kernel.
// The 512 x 768 weight read per sample is the whole cost (latency-bound): HB_SPLIT CTAs per sample each take a quarter
// of the hidden units, add their partial dy into the save area, and the last one to arrive (ticket) finishes the
// LayerNorm backward.  (head_fwd zeroes the accumulator and the ticket.)
constexpr int HB_SPLIT = 4;
constexpr int HB_T = HD_HID / HB_SPLIT;   // 128 hidden units per CTA

__global__ void __launch_bounds__(HD_THREADS)
head_bwd_data_kernel(const float* __restrict__ dlogits, float* save, const float* __restrict__ ln_w,
                     const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ mask1,
                     const float* __restrict__ mask2, float* __restrict__ dfeat, int num_classes) {
  pdl_sync_traced(TK_HEAD);
  __shared__ float dz[HB_T];
  __shared__ float red[HD_THREADS / 32];
  __shared__ int is_last;
  const int b = blockIdx.x, q = blockIdx.y, tid = threadIdx.x;
  float* sv = save + (int64_t)b * HS_TOTAL;
  if (tid < HB_T) {
    const int t = q * HB_T + tid;
    float da = 0.f;
    for (int c = 0; c < num_classes; ++c) da = fmaf(dlogits[(int64_t)b * num_classes + c], w2[(int64_t)c * HD_HID + t], da);
    if (mask2) da *= mask2[(int64_t)b * HD_HID + t];
    const float v = da * gelu_erf_grad(sv[HS_Z1 + t]);
    dz[tid] = v;
    sv[HS_DZ1 + t] = v;
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int k = tid + e * HD_THREADS;
    if (k < HD_IN) {
      float dy = 0.f;
      const float* wc = w1 + (int64_t)(q * HB_T) * HD_IN + k;
#pragma unroll 16
      for (int t = 0; t < HB_T; ++t) dy = fmaf(dz[t], __ldg(wc + (int64_t)t * HD_IN), dy);   // 16 loads in flight
      atomicAdd(sv + HS_ACC + k, dy);
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(reinterpret_cast<int*>(sv + HS_RSTD + 1), 1) == HB_SPLIT - 1);
  __syncthreads();
  if (!is_last) { trace_mark(TK_HEAD, 2); return; }
  __threadfence();
  if (tid == 0) sv[HS_RSTD + 1] = 0.f;        // ticket back to zero
  float dxh[2], xh[2];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int k = tid + e * HD_THREADS;
    dxh[e] = 0.f; xh[e] = 0.f;
    if (k < HD_IN) {
      float dy = __ldcg(sv + HS_ACC + k);      // complete sum (L2: the other CTAs' atomics)
      sv[HS_ACC + k] = 0.f;                    // self-cleaning: a second backward of the same forward starts from zero
      if (mask1) dy *= mask1[(int64_t)b * HD_IN + k];
      sv[HS_DYM + k] = dy;
      xh[e] = sv[HS_XHAT + k];
      dxh[e] = dy * ln_w[k];
      s1 += dxh[e];
      s2 += dxh[e] * xh[e];
    }
  }
  const float c1 = block_sum_512(s1, red) * (1.0f / HD_IN);
  const float c2 = block_sum_512(s2, red) * (1.0f / HD_IN);
  const float rstd = sv[HS_RSTD];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int k = tid + e * HD_THREADS;
    if (k < HD_IN) dfeat[(int64_t)b * HD_IN + k] = rstd * (dxh[e] - c1 - xh[e] * c2);
  }
  trace_end(TK_HEAD);
}

// parameter gradients: blockIdx.x in [0,64) -> 8 rows t of dw1 (+ db1) per CTA, so the saved activations of the batch
// are read 64 times instead of 512; 64 -> dln; 65.. -> dw2/db2 class rows
constexpr int HP_ROWS = 8;
constexpr int HP_BLOCKS = HD_HID / HP_ROWS;   // 64

__global__ void __launch_bounds__(256)
head_bwd_param_kernel(const float* __restrict__ dlogits, const float* __restrict__ save, const float* __restrict__ mask2,
                      float* __restrict__ dln_w, float* __restrict__ dln_b, float* __restrict__ dw1,
                      float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2, int batch,
                      int num_classes) {
  pdl_sync_traced(TK_HEAD);
  const int blk = blockIdx.x, tid = threadIdx.x;
  if (blk < HP_BLOCKS) {
    const int t0 = blk * HP_ROWS;
    float acc[HP_ROWS][3], sb[HP_ROWS];
#pragma unroll
    for (int r = 0; r < HP_ROWS; ++r) { acc[r][0] = acc[r][1] = acc[r][2] = 0.f; sb[r] = 0.f; }
#pragma unroll 4
    for (int b = 0; b < batch; ++b) {
      const float* sv = save + (int64_t)b * HS_TOTAL;
      const float4 d0 = *reinterpret_cast<const float4*>(sv + HS_DZ1 + t0);
      const float4 d1 = *reinterpret_cast<const float4*>(sv + HS_DZ1 + t0 + 4);
      const float d[HP_ROWS] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
      const float y0 = sv[HS_Y + tid], y1 = sv[HS_Y + tid + 256], y2 = sv[HS_Y + tid + 512];
#pragma unroll
      for (int r = 0; r < HP_ROWS; ++r) {
        sb[r] += d[r];
        acc[r][0] = fmaf(d[r], y0, acc[r][0]);
        acc[r][1] = fmaf(d[r], y1, acc[r][1]);
        acc[r][2] = fmaf(d[r], y2, acc[r][2]);
      }
    }
#pragma unroll
    for (int r = 0; r < HP_ROWS; ++r) {
#pragma unroll
      for (int e = 0; e < 3; ++e) dw1[(int64_t)(t0 + r) * HD_IN + tid + e * 256] += acc[r][e];
      if (tid == 0) db1[t0 + r] += sb[r];
    }
  } else if (blk == HP_BLOCKS) {
    for (int k = tid; k < HD_IN; k += 256) {
      float g = 0.f, bb = 0.f;
#pragma unroll 8
      for (int b = 0; b < batch; ++b) {
        const float* sv = save + (int64_t)b * HS_TOTAL;
        const float d = sv[HS_DYM + k];
        g = fmaf(d, sv[HS_XHAT + k], g);
        bb += d;
      }
      dln_w[k] += g;
      dln_b[k] += bb;
    }
  } else {
    const int c = blk - HP_BLOCKS - 1;
    float sb = 0.f;
    for (int t = tid; t < HD_HID; t += 256) {
      float a = 0.f;
#pragma unroll 8
      for (int b = 0; b < batch; ++b) {
        const float* sv = save + (int64_t)b * HS_TOTAL;
        float g = gelu_erf(sv[HS_Z1 + t]);
        if (mask2) g *= mask2[(int64_t)b * HD_HID + t];
        a = fmaf(dlogits[(int64_t)b * num_classes + c], g, a);
      }
      dw2[(int64_t)c * HD_HID + t] += a;
    }
    if (tid == 0) {
      for (int b = 0; b < batch; ++b) sb += dlogits[(int64_t)b * num_classes + c];
      db2[c] += sb;
    }
  }
  trace_end(TK_HEAD);
}

constexpr int FOCAL_MAX_C = 8;

__global__ void __launch_bounds__(256)
focal_kernel(const float* __restrict__ logits, const int64_t* __restrict__ targets, const float* __restrict__ alpha,
             float gamma, int reduction, float grad_scale, float* __restrict__ loss_per_sample,
             float* __restrict__ loss_out, float* __restrict__ dlogits, float* __restrict__ probs1,
             int64_t* __restrict__ preds, int* __restrict__ ncorrect, int batch, int C) {
  pdl_sync_traced(TK_FOCAL);
  __shared__ float red[8];
  __shared__ int redi[8];
  const float gscale = grad_scale * (reduction == 0 ? 1.0f / (float)batch : 1.0f);
  float lsum = 0.f;
  int correct = 0;
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    float z[FOCAL_MAX_C];
    float mx = -INFINITY;
    int am = 0;
#pragma unroll
    for (int c = 0; c < FOCAL_MAX_C; ++c)
      if (c < C) {
        z[c] = logits[(int64_t)b * C + c];
        if (z[c] > mx) { mx = z[c]; am = c; }
      }
    float se = 0.f;
#pragma unroll
    for (int c = 0; c < FOCAL_MAX_C; ++c)
      if (c < C) se += expf(z[c] - mx);
    const float lse = mx + logf(se);
    const int64_t yt = targets[b];
    const bool bad_target = yt < 0 || yt >= (int64_t)C;   // F.cross_entropy device-asserts; here the loss and the gradient
    const int y = bad_target ? 0 : (int)yt;               // of that sample become NaN (ignore_index is not supported)
    float zy = 0.f;
#pragma unroll
    for (int c = 0; c < FOCAL_MAX_C; ++c)
      if (c == y) zy = z[c];
    const float ce = bad_target ? __int_as_float(0x7fc00000) : lse - zy;
    const float pt = expf(-ce);
    const float om = 1.0f - pt;
    const float w = alpha[y];
    const float fl = w * powf(om, gamma) * ce;
    loss_per_sample[b] = fl;
    lsum += fl;
    if (dlogits) {
      // d fl / d ce = w * [ (1-pt)^g + ce * g * (1-pt)^(g-1) * pt ]
      const float dfl = w * (powf(om, gamma) + ce * gamma * powf(om, gamma - 1.0f) * pt);
#pragma unroll
      for (int c = 0; c < FOCAL_MAX_C; ++c)
        if (c < C) {
          const float sm = expf(z[c] - lse);
          dlogits[(int64_t)b * C + c] = gscale * dfl * (sm - (c == y ? 1.0f : 0.0f));
        }
    }
    if (probs1) probs1[b] = (C > 1) ? expf(z[1] - lse) : 1.0f;
    if (preds) preds[b] = am;
    correct += (am == y);
  }
  // deterministic block reduction (fixed order)
  lsum = warp_sum(lsum);
  for (int o = 16; o > 0; o >>= 1) correct += __shfl_xor_sync(0xffffffffu, correct, o);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = lsum; redi[threadIdx.x >> 5] = correct; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    int n = 0;
    for (int w = 0; w < 8; ++w) { s += red[w]; n += redi[w]; }
    if (loss_out) loss_out[0] = (reduction == 0) ? s / (float)batch : s;
    if (ncorrect) ncorrect[0] = n;
  }
  trace_end(TK_FOCAL);
}

}  // namespace vitk

using namespace vitk;

extern "C" size_t vitk_head_save_floats(int batch) { return (size_t)batch * HS_TOTAL; }

extern "C" int vitk_head_fwd(const float* feat, const float* ln_w, const float* ln_b, const float* w1, const float* b1,
                             const float* w2, const float* b2, const float* mask1, const float* mask2, float* logits,
                             float* save, int batch, int num_classes, void* stream) {
  VITK_CHECK_ARG(feat && ln_w && ln_b && w1 && b1 && w2 && b2 && logits && batch > 0 && num_classes > 0);
  VITK_LAUNCH((head_fwd_kernel), batch, HD_THREADS, 0, (cudaStream_t)stream, feat, ln_w, ln_b, w1, b1, w2, b2, mask1, mask2, logits, save, num_classes);
  return VITK_OK;
}

extern "C" int vitk_head_bwd(const float* dlogits, float* save, const float* ln_w, const float* w1, const float* w2,
                             const float* mask1, const float* mask2, float* dfeat, float* dln_w, float* dln_b,
                             float* dw1, float* db1, float* dw2, float* db2, int batch, int num_classes, void* stream) {
  VITK_CHECK_ARG(dlogits && save && ln_w && w1 && w2 && dfeat && batch > 0 && num_classes > 0);
  cudaStream_t st = (cudaStream_t)stream;
  VITK_LAUNCH((head_bwd_data_kernel), dim3(batch, HB_SPLIT), HD_THREADS, 0, st, dlogits, save, ln_w, w1, w2, mask1, mask2, dfeat, num_classes);
  if (dw1) {
    VITK_CHECK_ARG(dln_w && dln_b && db1 && dw2 && db2);
    VITK_LAUNCH((head_bwd_param_kernel), HP_BLOCKS + 1 + num_classes, 256, 0, st, dlogits, save, mask2, dln_w, dln_b, dw1, db1, dw2, db2, batch, num_classes);
  }
  return VITK_OK;
}

extern "C" int vitk_focal_fwd_bwd(const float* logits, const int64_t* targets, const float* alpha, float gamma,
                                  int reduction, float grad_scale, float* loss_per_sample, float* loss_out,
                                  float* dlogits, float* probs1, int64_t* preds, int* ncorrect, int batch,
                                  int num_classes, void* stream) {
  VITK_CHECK_ARG(logits && targets && alpha && loss_per_sample && batch > 0);
  VITK_CHECK_ARG(num_classes >= 1 && num_classes <= FOCAL_MAX_C && reduction >= 0 && reduction <= 2);
  VITK_LAUNCH((focal_kernel), 1, 256, 0, (cudaStream_t)stream, logits, targets, alpha, gamma, reduction, grad_scale, loss_per_sample, loss_out, dlogits, probs1, preds, ncorrect, batch, num_classes);
  return VITK_OK;
}
