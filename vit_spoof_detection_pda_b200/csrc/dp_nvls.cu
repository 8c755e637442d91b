// Data-parallel optimizer step fused with its collectives over NVLink 5 / NVSwitch multicast (NVLS):
//
//      gradient all-reduce  +  global-norm clip  +  Adam / AdamW  +  parameter all-gather
//
// in TWO kernels per rank that touch every gradient and parameter byte exactly once.  The reference has no multi-GPU code
// (SURVEY.md 2.1); a stock data-parallel loop would run  ncclAllReduce(grads) -> clip_grad_norm_ -> Adam  with the full
// 345 MB fp32 all-reduce taking SMs from the backward GEMMs while it is in flight and every rank repeating the full
// 2.6 GB Adam pass.  Here the flat gradient / parameter / bf16-shadow buffers of all ranks are symmetric memory mapped
// behind ONE multicast address each (torch.distributed._symmetric_memory is the plumbing: allocation, handle exchange,
// stream-ordered barriers), rank r owns slice r of the flat index space, and
//
//   kernel 1  vitk_nvls_reduce_sumsq   g[slice r] = (1 / world) * multimem.ld_reduce.add.v4.f32 [g_mc + slice r]
//             -- the NVSwitch adds the eight ranks' values in flight and returns ONE reduced vector: the reduce-scatter of a
//             two-shot all-reduce, written back in place into the local slice -- together with the slice's sum of squares
//             (the global-norm clip needs the norm of the REDUCED gradient), broadcast to slot r of every rank's norm table
//             with multimem.st;
//   kernel 2  vitk_nvls_adam_bcast     clip coefficient from the eight partial sums, Adam / AdamW on slice r (moments exist
//             only for the slice: 1/8 of the optimizer traffic per rank), and the updated fp32 masters AND the bf16 weight shadow
//             of the slice stored through the multicast addresses: every rank's copy is written by the switch (the
//             all-gather).
//
// The flat index space is cut into a few DOMAINS along the order in which backward finalises gradients (dp.py); rank r owns
// slice r of every domain.  Kernel 1 of all but the last domain runs on a side stream UNDER the backward pass with a bounded
// number of CTAs -- shared-memory-free, so they sit next to the persistent GEMM CTAs instead of taking their SMs (the NCCL
// path had to give up 32 of 148 SMs at 8 GPUs) -- and only the last domain's reduce-scatter plus kernel 2 remain after it.
// Ordering between ranks is by stream-ordered symmetric-memory barriers (host side, dp.py): domain final on every rank ->
// kernel 1 ... -> norm table complete -> kernel 2 -> parameters complete.
#include <math.h>

#include "common.cuh"

namespace vitk {

constexpr int NVLS_MAX_CTAS = 1024;

__device__ __forceinline__ float4 mc_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st_f32x4(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mc_st_b32x2(void* mc, uint32_t a, uint32_t b) {     // two 32-bit words (4 bf16) as a bit pattern
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(mc), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)) : "memory");
}
__device__ __forceinline__ void mc_st_f32(float* mc, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(mc), "f"(v) : "memory");
}

// kernel 1: reduce-scatter of this rank's slice through the switch + partial sums of squares of the reduced slice
__global__ void __launch_bounds__(256)
nvls_reduce_sumsq_kernel(float* __restrict__ g_local, const float* __restrict__ g_mc, size_t n4, float scale,
                         float* __restrict__ partial) {
  pdl_sync_traced(TK_SUMSQ);
  __shared__ float red[8];
  float s = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = mc_ld_reduce_add(g_mc + 4 * i);
    v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    reinterpret_cast<float4*>(g_local)[i] = v;
    s += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
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
// fp32 masters of the slice -> every rank, with a bounded number of CTAs (runs on a side stream next to the forward GEMMs)
__global__ void __launch_bounds__(256)
nvls_bcast_f32_kernel(const float* __restrict__ src, float* __restrict__ dst_mc, size_t n4) {
  pdl_sync_traced(TK_CAST);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    mc_st_f32x4(dst_mc + 4 * i, reinterpret_cast<const float4*>(src)[i]);
  trace_end(TK_CAST);
}
// deterministic second stage; the slice's sum goes to slot `rank` of EVERY rank's table through the multicast address
__global__ void __launch_bounds__(256)
nvls_sumsq_bcast_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ table_mc, int rank) {
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
    mc_st_f32(table_mc + rank, t);
  }
  trace_end(TK_SUMSQ);
}

struct NvlsAdamArgs {
  float beta1, beta2, omb1, omb2, eps, wd, step_size, bc2_sqrt, decay, grad_mult, max_norm;
  int mode, world;
  // [lo4, hi4) ranges (float4 units, relative to the slice) whose fp32 masters are NOT multicast by this kernel: the weights
  // of the tensor-core GEMMs, which forward / backward only ever read through the bf16 shadow.  Their masters are stored
  // locally here and reach the other ranks by vitk_nvls_bcast_f32 on a side stream, under the next forward pass.
  int n_ranges;
  uint32_t lo4[64], hi4[64];
};

__device__ __forceinline__ void nvls_adam_one(float& p, float g, float& m, float& v, const NvlsAdamArgs& a, float gm) {
  g *= gm;
  if (a.mode == 1) p *= a.decay;
  else if (a.wd != 0.f) g = fmaf(a.wd, p, g);
  m = m + (g - m) * a.omb1;
  v = fmaf(a.omb2, g * g, v * a.beta2);
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p = p - a.step_size * (m / denom);
}

// kernel 2: Adam on the slice (same arithmetic order as adam.cu / torch), results stored through the multicast addresses
__global__ void __launch_bounds__(256)
nvls_adam_bcast_kernel(const float* p_local, float* p_local_w, float* __restrict__ p_mc, void* __restrict__ p16_mc,
                       const float* __restrict__ g_local, float* __restrict__ m, float* __restrict__ v, size_t n4, NvlsAdamArgs a,
                       const float* __restrict__ sq_table) {
  pdl_sync_traced(TK_ADAM);
  float gm = a.grad_mult;
  if (sq_table && a.max_norm > 0.f) {
    float tot = 0.f;
    for (int r = 0; r < a.world; ++r) tot += sq_table[r];      // same order on every rank: identical coefficient everywhere
    const float total = sqrtf(tot) * a.grad_mult;
    gm *= fminf(1.0f, a.max_norm / (total + 1e-6f));
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 pv = reinterpret_cast<const float4*>(p_local)[i];
    const float4 gv = reinterpret_cast<const float4*>(g_local)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    nvls_adam_one(pv.x, gv.x, mv.x, vv.x, a, gm);
    nvls_adam_one(pv.y, gv.y, mv.y, vv.y, a, gm);
    nvls_adam_one(pv.z, gv.z, mv.z, vv.z, a, gm);
    nvls_adam_one(pv.w, gv.w, mv.w, vv.w, a, gm);
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    bool local_only = false;
    for (int r = 0; r < a.n_ranges; ++r) local_only |= (i >= a.lo4[r]) & (i < a.hi4[r]);
    if (local_only) reinterpret_cast<float4*>(p_local_w)[i] = pv;
    else mc_st_f32x4(p_mc + 4 * i, pv);
    if (p16_mc) mc_st_b32x2(reinterpret_cast<char*>(p16_mc) + 8 * i, pack_bf16x2(pv.x, pv.y), pack_bf16x2(pv.z, pv.w));
  }
  trace_end(TK_ADAM);
}

static int nvls_grid(size_t n4) {
  const size_t want = (n4 + 255) / 256;
  const size_t cap = (size_t)sm_count() * 8;
  size_t g = want < cap ? (want ? want : 1) : cap;
  if (g > NVLS_MAX_CTAS) g = NVLS_MAX_CTAS;
  return (int)g;
}

}  // namespace vitk

using namespace vitk;

extern "C" size_t vitk_nvls_scratch_floats(void) { return NVLS_MAX_CTAS; }

extern "C" int vitk_nvls_reduce_sumsq(float* g_local, const float* g_mc, size_t n, float scale, float* partial, float* table_mc,
                                      int slot, int max_ctas, void* stream) {
  VITK_CHECK_ARG(g_local && g_mc && partial && table_mc && slot >= 0 && n % 4 == 0 && max_ctas >= 0);
  VITK_CHECK_ARG(((uintptr_t)g_local % 16) == 0 && ((uintptr_t)g_mc % 16) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = nvls_grid(n / 4);
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;      // a reduce that runs under the backward pass leaves the SMs to it
  VITK_LAUNCH((nvls_reduce_sumsq_kernel), grid, 256, 0, st, g_local, g_mc, n / 4, scale, partial);
  VITK_LAUNCH((nvls_sumsq_bcast_kernel), 1, 256, 0, st, partial, grid, table_mc, slot);
  return VITK_OK;
}

extern "C" int vitk_nvls_bcast_f32(const float* src, float* dst_mc, size_t n, int max_ctas, void* stream) {
  VITK_CHECK_ARG(src && dst_mc && n % 4 == 0 && max_ctas >= 1 && ((uintptr_t)src % 16) == 0 && ((uintptr_t)dst_mc % 16) == 0);
  int grid = nvls_grid(n / 4);
  if (grid > max_ctas) grid = max_ctas;
  VITK_LAUNCH((nvls_bcast_f32_kernel), grid, 256, 0, (cudaStream_t)stream, src, dst_mc, n / 4);
  return VITK_OK;
}

extern "C" int vitk_nvls_adam_bcast(float* p_local, float* p_mc, void* p16_mc, const float* g_local, float* m, float* v,
                                    size_t n, double lr, double beta1, double beta2, double eps, double weight_decay, int mode,
                                    int step, float grad_mult, const float* sq_table, int world, float max_norm,
                                    const int64_t* local_only_ranges, int n_ranges, void* stream) {
  VITK_CHECK_ARG(p_local && p_mc && g_local && m && v && step >= 1 && (mode == 0 || mode == 1) && n % 4 == 0 && world >= 1);
  VITK_CHECK_ARG(((uintptr_t)p_local % 16) == 0 && ((uintptr_t)p_mc % 16) == 0 && ((uintptr_t)g_local % 16) == 0 &&
                 ((uintptr_t)m % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)p16_mc % 8) == 0);
  VITK_CHECK_ARG(n_ranges >= 0 && n_ranges <= 64 && (n_ranges == 0 || local_only_ranges));
  NvlsAdamArgs a;
  a.n_ranges = n_ranges;
  for (int r = 0; r < n_ranges; ++r) {      // element ranges relative to the slice, multiples of 4
    VITK_CHECK_ARG(local_only_ranges[2 * r] % 4 == 0 && local_only_ranges[2 * r + 1] % 4 == 0);
    a.lo4[r] = (uint32_t)(local_only_ranges[2 * r] / 4);
    a.hi4[r] = (uint32_t)(local_only_ranges[2 * r + 1] / 4);
  }
  a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps; a.wd = (float)weight_decay;
  a.mode = mode; a.world = world;
  a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
  a.grad_mult = grad_mult; a.max_norm = max_norm;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  a.step_size = (float)(lr / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  a.decay = (float)(1.0 - lr * weight_decay);
  VITK_LAUNCH((nvls_adam_bcast_kernel), nvls_grid(n / 4), 256, 0, (cudaStream_t)stream, (const float*)p_local, p_local, p_mc, p16_mc,
              g_local, m, v, n / 4, a, sq_table);
  return VITK_OK;
}
