// Fused multi-tensor Adam / AdamW over one flat fp32 buffer -- HBM-bound (28 B/param, +2 B/param for
// the bf16 shadow).  Replaces the torch foreach path behind
//   scaler.unscale_ / clip_grad_norm_(model.parameters(), 1.0) / scaler.step(optimizer)
// at /root/reference/train_advanced.py:333-335 with optimizer = AdamW(lr 3e-4, wd 0.05, betas (0.9,0.999))
// (train_advanced.py:592-597) or Adam + L2 weight decay 1e-4 (README.md:140-147).
// Arithmetic order follows torch/optim/adam.py (_single_tensor_adam / _multi_tensor_adam):
//   AdamW: p *= 1 - lr*wd        | Adam: g += wd*p
//   m = m + (g - m)*(1-b1)  (lerp) ; v = v*b2 + (1-b2)*g*g
//   denom = sqrt(v)/sqrt(1-b2^t) + eps ; p -= (lr/(1-b1^t)) * m/denom
#include <math.h>

#include "common.cuh"

namespace vitk {

constexpr int SUMSQ_MAX_CTAS = 1024;

__global__ void __launch_bounds__(256)
sumsq_partial_kernel(const float* __restrict__ g, size_t n, float* __restrict__ partial) {
  pdl_sync_traced(TK_SUMSQ);
  __shared__ float red[8];
  float s = 0.f;
  const size_t n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0)
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
  trace_end(TK_SUMSQ);
}

__global__ void __launch_bounds__(256)
sumsq_final_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ sumsq) {
  pdl_sync_traced(TK_SUMSQ);
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partial[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    sumsq[0] = t;
  }
  trace_end(TK_SUMSQ);
}

struct AdamArgs {
  float lr, beta1, beta2, omb1, omb2, eps, wd, step_size, bc2_sqrt, decay, grad_mult, max_norm;
  int mode;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a, float gm) {
  g *= gm;
  if (a.mode == 1) p *= a.decay;
  else if (a.wd != 0.f) g = fmaf(a.wd, p, g);
  m = m + (g - m) * a.omb1;
  v = fmaf(a.omb2, g * g, v * a.beta2);  // addcmul_(g, g, value = 1-b2): value is rounded from double
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p = p - a.step_size * (m / denom);
}

// Step-dependent scalars computed ON THE DEVICE (CUDA-graph capture of the training step: a replayed launch cannot carry a new
// step count or learning rate in its arguments): step = ++*step_dev; hyper = {step_size, sqrt(bias correction 2), decay}
__global__ void adam_prepare_kernel(const float* __restrict__ lr_dev, int* __restrict__ step_dev, double beta1, double beta2,
                                    double wd, float* __restrict__ hyper) {
  pdl_sync_traced(TK_ADAM);
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int step = *step_dev + 1;
    *step_dev = step;
    const double lr = (double)lr_dev[0];
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    hyper[0] = (float)(lr / bc1);
    hyper[1] = (float)sqrt(bc2);
    hyper[2] = (float)(1.0 - lr * wd);
  }
  trace_end(TK_ADAM);
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            bf16* __restrict__ p16, size_t n, AdamArgs a, const float* __restrict__ sumsq, const float* __restrict__ hyper) {
  pdl_sync_traced(TK_ADAM);
  if (hyper) { a.step_size = hyper[0]; a.bc2_sqrt = hyper[1]; a.decay = hyper[2]; }
  float gm = a.grad_mult;
  if (sumsq && a.max_norm > 0.f) {
    const float total = sqrtf(sumsq[0]) * a.grad_mult;
    gm *= fminf(1.0f, a.max_norm / (total + 1e-6f));
  }
  const size_t n4 = n / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adam_one(pv.x, gv.x, mv.x, vv.x, a, gm);
    adam_one(pv.y, gv.y, mv.y, vv.y, a, gm);
    adam_one(pv.z, gv.z, mv.z, vv.z, a, gm);
    adam_one(pv.w, gv.w, mv.w, vv.w, a, gm);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (p16) {
      uint2 u;
      u.x = pack_bf16x2(pv.x, pv.y);
      u.y = pack_bf16x2(pv.z, pv.w);
      reinterpret_cast<uint2*>(p16)[i] = u;
    }
  }
  if (blockIdx.x == 0)
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      float pv = p[i], mv = m[i], vv = v[i];
      adam_one(pv, g[i], mv, vv, a, gm);
      p[i] = pv; m[i] = mv; v[i] = vv;
      if (p16) p16[i] = __float2bfloat16_rn(pv);
    }
  trace_end(TK_ADAM);
}

// g *= grad_mult * clipcoef in place (eager unscale / clip for callers that mix the fused pieces with stock torch ones:
// optim.py FusedGradScaler.unscale_, clip_grad_norm_ without a FusedAdam)
__global__ void __launch_bounds__(256)
grad_scale_kernel(float* __restrict__ g, size_t n, float grad_mult, const float* __restrict__ sumsq, float max_norm) {
  pdl_sync_traced(TK_SUMSQ);
  float gm = grad_mult;
  if (sumsq && max_norm > 0.f) {
    const float total = sqrtf(sumsq[0]) * grad_mult;
    gm *= fminf(1.0f, max_norm / (total + 1e-6f));
  }
  const size_t n4 = n / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<float4*>(g)[i];
    v.x *= gm; v.y *= gm; v.z *= gm; v.w *= gm;
    reinterpret_cast<float4*>(g)[i] = v;
  }
  if (blockIdx.x == 0)
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) g[i] *= gm;
  trace_end(TK_SUMSQ);
}

__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n) {
  pdl_sync_traced(TK_CAST);
  const size_t n4 = n / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(dst)[i] = u;
  }
  if (blockIdx.x == 0)
    for (size_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) dst[i] = __float2bfloat16_rn(src[i]);
  trace_end(TK_CAST);
}

static int stream_grid(size_t n4) {
  const size_t want = (n4 + 255) / 256;
  const size_t cap = (size_t)sm_count() * 8;
  return (int)(want < cap ? (want ? want : 1) : cap);
}

}  // namespace vitk

using namespace vitk;

extern "C" size_t vitk_grad_sumsq_scratch_floats(void) { return SUMSQ_MAX_CTAS; }

extern "C" int vitk_grad_sumsq(const float* g, size_t n, float* partial, float* sumsq, void* stream) {
  VITK_CHECK_ARG(g && partial && sumsq && ((uintptr_t)g % 16) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = stream_grid(n / 4);
  if (grid > SUMSQ_MAX_CTAS) grid = SUMSQ_MAX_CTAS;
  VITK_LAUNCH((sumsq_partial_kernel), grid, 256, 0, st, g, n, partial);
  VITK_LAUNCH((sumsq_final_kernel), 1, 256, 0, st, partial, grid, sumsq);
  return VITK_OK;
}

extern "C" int vitk_grad_scale(float* g, size_t n, float grad_mult, const float* sumsq, float max_norm, void* stream) {
  VITK_CHECK_ARG(g && ((uintptr_t)g % 16) == 0);
  VITK_LAUNCH((grad_scale_kernel), stream_grid(n / 4), 256, 0, (cudaStream_t)stream, g, n, grad_mult, sumsq, max_norm);
  return VITK_OK;
}

extern "C" int vitk_adam_step(float* p, const float* g, float* m, float* v, void* p16, size_t n, double lr,
                              double beta1, double beta2, double eps, double weight_decay, int mode, int step,
                              float grad_mult, const float* sumsq, float max_norm, void* stream) {
  VITK_CHECK_ARG(p && g && m && v && step >= 1 && (mode == 0 || mode == 1));
  VITK_CHECK_ARG(((uintptr_t)p % 16) == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)m % 16) == 0 && ((uintptr_t)v % 16) == 0);
  VITK_CHECK_ARG(p16 == nullptr || ((uintptr_t)p16 % 8) == 0);
  AdamArgs a;
  a.lr = (float)lr; a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps; a.wd = (float)weight_decay;
  a.mode = mode;
  a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
  a.grad_mult = grad_mult; a.max_norm = max_norm;
  // python-double scalar math, as torch does for the non-capturable path
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  a.step_size = (float)(lr / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  a.decay = (float)(1.0 - lr * weight_decay);
  VITK_LAUNCH((adam_kernel), stream_grid(n / 4), 256, 0, (cudaStream_t)stream, p, g, m, v, (bf16*)p16, n, a, sumsq, (const float*)nullptr);
  return VITK_OK;
}

extern "C" int vitk_adam_step_graph(float* p, const float* g, float* m, float* v, void* p16, size_t n, const float* lr_dev,
                                    int* step_dev, float* hyper, double beta1, double beta2, double eps, double weight_decay,
                                    int mode, float grad_mult, const float* sumsq, float max_norm, void* stream) {
  VITK_CHECK_ARG(p && g && m && v && lr_dev && step_dev && hyper && (mode == 0 || mode == 1));
  VITK_CHECK_ARG(((uintptr_t)p % 16) == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)m % 16) == 0 && ((uintptr_t)v % 16) == 0);
  VITK_CHECK_ARG(p16 == nullptr || ((uintptr_t)p16 % 8) == 0);
  AdamArgs a;
  a.lr = 0.f; a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps; a.wd = (float)weight_decay;
  a.mode = mode;
  a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
  a.grad_mult = grad_mult; a.max_norm = max_norm;
  a.step_size = 0.f; a.bc2_sqrt = 1.f; a.decay = 1.f;      // replaced by `hyper` on the device
  cudaStream_t st = (cudaStream_t)stream;
  VITK_LAUNCH((adam_prepare_kernel), 1, 32, 0, st, lr_dev, step_dev, beta1, beta2, weight_decay, hyper);
  VITK_LAUNCH((adam_kernel), stream_grid(n / 4), 256, 0, st, p, g, m, v, (bf16*)p16, n, a, sumsq, (const float*)hyper);
  return VITK_OK;
}

extern "C" int vitk_cast_f32_to_bf16(const float* src, void* dst, size_t n, void* stream) {
  VITK_CHECK_ARG(src && dst && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst % 8) == 0);
  VITK_LAUNCH((cast_bf16_kernel), stream_grid(n / 4), 256, 0, (cudaStream_t)stream, src, (bf16*)dst, n);
  return VITK_OK;
}
