// Shared helpers for libvitk (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vitk.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libvitk is written for sm_100a only"
#endif

namespace vitk {

typedef __nv_bfloat16 bf16;

void set_error(const char* fmt, ...);
void count_launch();
int sm_count();
int set_sm_budget(int n);   // thread-local; returns the previous value
int set_max_dyn_smem_once(const void* fn, int bytes);

#define VITK_CHECK_ARG(cond)                                                         \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      vitk::set_error("%s:%d: argument check failed: %s", __FILE__, __LINE__, #cond); \
      return VITK_ERR_ARG;                                                           \
    }                                                                                \
  } while (0)

#define VITK_CUDA(expr)                                                                        \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      vitk::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
      return (int)_e;                                                                          \
    }                                                                                          \
  } while (0)

// every kernel launch in the library goes through this macro: it also counts launches (bench.py "gpu_launches")
#define VITK_LAUNCH_CHECK()                 \
  do {                                      \
    vitk::count_launch();                   \
    VITK_CUDA(cudaPeekAtLastError());       \
  } while (0)

// Host-side cache of encoded TMA tensor maps.  Workspaces, parameters and gradient buffers are persistent, so the same
// (base, dims, strides, box, type, swizzle) keys recur every step; cuTensorMapEncodeTiled costs microseconds and a
// training step needs ~700 maps -- a lookup keeps the launch path off the host's critical path.  A descriptor depends on
// nothing but its key, so entries never go stale.  `key` is hashed as raw bytes: zero-initialise it before filling it in.
struct TmapKey {
  const void* base;
  uint64_t dims[3];
  uint64_t strides[2];
  uint32_t box[3];
  uint32_t rank, dtype, swizzle, l2promo;
};
bool tmap_cache_get(const TmapKey& key, void* map128);   // map128: CUtensorMap (128 bytes)
void tmap_cache_put(const TmapKey& key, const void* map128);

// Programmatic dependent launch (PDL).  Every kernel of the library is launched with the programmatic-stream-
// serialization attribute and begins with pdl_sync(): griddepcontrol.wait (returns once the preceding kernel of the
// stream has completed and flushed) followed by griddepcontrol.launch_dependents (the next kernel may start being
// scheduled as SMs drain).  Launch latency, CTA ramp-up and the prologue in front of pdl_sync() -- barrier init,
// TMEM allocation, tensor-map prefetch -- overlap the tail of the previous kernel; no kernel touches global memory
// before its pdl_sync().  vitk_debug_set(6, 1) turns the attribute off (plain stream order) for A/B timing.
bool pdl_enabled();

// Device-side tracer -- exists only in the development build (libvitk_dev.so, -DVITK_DEV; see build.py).  Thread 0 of
// every CTA appends (globaltimer, kernel id | phase | SM id | block index, aux) records to a caller-provided device
// buffer: phase 0 = CTA started, 1 = its stream dependencies are satisfied (griddepcontrol.wait returned), 2 = CTA
// finished.  That is a per-SM timeline of a real training step with programmatic dependent launch and the
// weight-gradient side stream left ON (tools/step_timeline.py) -- what CUDA-event bracketing cannot give without
// serialising the launches.  In the release build trace_mark() compiles to nothing.
enum TraceKernel {
  TK_GEMM_TC = 1, TK_ATTN_FWD = 2, TK_ATTN_BWD = 3, TK_LN_FWD = 4, TK_LN_BWD = 5, TK_ADAM = 6, TK_SUMSQ = 7, TK_CAST = 8,
  TK_HEAD = 9, TK_FOCAL = 10, TK_COLSUM = 11, TK_PATCH_EMBED = 12, TK_EMBED_GRADS = 13, TK_SCATTER_CLS = 14,
  TK_GEMM_SIMT = 15, TK_ATTN_SIMT = 16, TK_EVAL = 17, TK_PATCH_WGRAD = 18,
  TK_GEMM_TAIL = 19
};
#ifdef VITK_DEV
void trace_register(void (*setter)(unsigned long long*));   // api.cu: one setter per translation unit
#endif
#ifdef __CUDACC__
#ifdef VITK_DEV
static __device__ unsigned long long* g_trace_buf = nullptr;   // [0] record count, [1] capacity, [2] detail offset, [3] pad, records of 3 words from [4]
namespace {
struct TraceTU {
  static void set(unsigned long long* p) { cudaMemcpyToSymbol(g_trace_buf, &p, sizeof(p)); }
  TraceTU() { trace_register(&set); }
};
static TraceTU s_trace_tu;
}  // namespace
__device__ __forceinline__ void trace_mark(uint32_t kid, uint32_t phase, unsigned long long aux = 0) {
  if (threadIdx.x != 0) return;
  unsigned long long* b = g_trace_buf;
  if (!b) return;
  const unsigned long long idx = atomicAdd(b, 1ull);
  if (idx >= b[1]) return;
  unsigned long long t;
  uint32_t smid;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  unsigned long long* r = b + 4 + 3 * idx;
  r[0] = t;
  r[1] = ((unsigned long long)kid << 48) | ((unsigned long long)phase << 40) | ((unsigned long long)(smid & 0xffffu) << 24) |
         (unsigned long long)(blockIdx.x & 0xffffffu);
  r[2] = aux;
}
// Kernel-internal marks (CTA 0 only, any single thread the caller elects): timestamps go to STATIC slots of a detail region
// behind the launch marks -- slot = (phase - 8) * 64 + (aux & 63), overwritten by every launch -- so that a mark is one
// special-register read and one fire-and-forget store (an atomically allocated record would stall the marking thread for
// an L2 round trip and distort exactly the fine-grained timing it is meant to show).  trace_detail_base() is called once
// per warp role (one dependent load), trace_detail() per event.
constexpr int TRACE_DETAIL_PHASES = 24, TRACE_DETAIL_SLOTS = 64;
__device__ __forceinline__ unsigned long long* trace_detail_base(uint32_t kid) {
  if (blockIdx.x != 0) return nullptr;
  unsigned long long* b = g_trace_buf;
  if (!b) return nullptr;
  const unsigned long long off = b[2];          // word offset of the detail area (0 = off)
  return off ? b + off + (unsigned long long)kid * (TRACE_DETAIL_PHASES * TRACE_DETAIL_SLOTS) : nullptr;
}
__device__ __forceinline__ void trace_detail(unsigned long long* dbase, uint32_t phase, uint32_t aux = 0) {
  if (!dbase) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  dbase[(phase - 8) * TRACE_DETAIL_SLOTS + (aux & (TRACE_DETAIL_SLOTS - 1))] = t;
}
// end-of-CTA mark for kernels whose threads all reach the end of the kernel body
__device__ __forceinline__ void trace_end(uint32_t kid, unsigned long long aux = 0) {
  __syncthreads();
  trace_mark(kid, 2, aux);
}
#else
__device__ __forceinline__ void trace_mark(uint32_t, uint32_t, unsigned long long = 0) {}
__device__ __forceinline__ unsigned long long* trace_detail_base(uint32_t) { return nullptr; }
__device__ __forceinline__ void trace_detail(unsigned long long*, uint32_t, uint32_t = 0) {}
__device__ __forceinline__ void trace_end(uint32_t, unsigned long long = 0) {}
#endif
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
// pdl_sync() for kernels that start with it: CTA-start and dependencies-satisfied marks around the wait
__device__ __forceinline__ void pdl_sync_traced(uint32_t kid, unsigned long long aux = 0) {
  trace_mark(kid, 0, aux);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  trace_mark(kid, 1, aux);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KA, typename... A>
static inline cudaError_t launch_kernel(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
}
#endif
// launch + count + error check
#define VITK_LAUNCH(kern, grid, block, smem, st, ...)                                  \
  do {                                                                                 \
    vitk::count_launch();                                                              \
    VITK_CUDA(vitk::launch_kernel(kern, dim3(grid), dim3(block), smem, st, __VA_ARGS__)); \
  } while (0)

#define VITK_TRY(expr)          \
  do {                          \
    int _r = (expr);            \
    if (_r != 0) return _r;     \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// exact-erf GELU (nn.GELU default, timm Mlp / train_advanced.py:197) and its derivative
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// Fast GELU and GELU' together for the bf16 tensor-core epilogue (outputs are rounded to bf16, rel. step 2^-8):
// erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7) with one MUFU.EX2 + one MUFU.RCP per element; the same
// exponential e^{-x^2/2} serves erf(x/sqrt2) and the normal pdf.  ~17 instructions for both outputs.
// The fp32-validate path keeps erff / expf.
__device__ __forceinline__ void gelu_fast_both(float x, float& g, float& dg) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170368f));  // exp(-x^2/2)
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, e, 1.0f);
  const float cdf = fmaf(0.5f, copysignf(erf_abs, x), 0.5f);
  g = x * cdf;
  dg = fmaf(x * 0.39894228040143267794f, e, cdf);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int dtype_size(int dtype) { return dtype == VITK_BF16 ? 2 : 4; }

}  // namespace vitk
