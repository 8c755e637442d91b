// LayerNorm(768) forward / backward -- HBM-bound, one warp per row, row held in registers.
// Replaces ATen native_layer_norm reached from timm Block.norm1/.norm2, vit.norm (eps 1e-6) and
// classifier[0] (eps 1e-5; /root/reference/train_advanced.py:194).
//
// Algorithmic bytes per row: fwd 768*(4 in + sizeof(T) out) (+8 stats); bwd 768*(sizeof(T) dy + 4 x
// + 4 dres + 4 dx [+2 dx16]).
#include "common.cuh"

namespace vitk {

constexpr int LN_COLS = VITK_DIM;        // 768
constexpr int LN_VEC = LN_COLS / 128;    // 6 float4 per lane
constexpr int LN_WARPS = 8;

template <typename T> struct Vec4IO;
template <> struct Vec4IO<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4IO<bf16> {
  static __device__ __forceinline__ float4 load(const bf16* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void store(bf16* p, float4 v) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// Forward: one warp normalises one row.  The kernel is a pure stream (58 MB at batch 64) whose first version -- row in
// registers, loads issued by the threads -- ran at 2.8 TB/s inside the step (tools/step_timeline.py: 21 us per launch): too
// few bytes in flight per SM and four waves of short-lived CTAs.  Now the rows are STAGED THROUGH SHARED MEMORY BY THE
// BULK-COPY ENGINE: every warp owns a ring of LN_RING 3 KB row buffers, lane 0 keeps cp.async.bulk copies (one per row,
// completion on an mbarrier) in flight for the next LN_RING rows of the warp, and the lanes only ever read shared memory.
// One persistent wave of 2 CTAs per SM: 16 warps x 4 rows x 3 KB = 192 KB in flight per SM, no registers tied up by
// outstanding loads.
constexpr int LN_RING = 4;
constexpr int LN_ROW_BYTES = LN_COLS * 4;                         // 3072
constexpr int LN_FWD_SMEM = LN_WARPS * LN_RING * LN_ROW_BYTES + LN_WARPS * LN_RING * 8;   // 98,560 B

__device__ __forceinline__ uint32_t ln_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ln_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
// lane 0: arm the barrier and start the bulk copy of one row (3072 B, 16-byte aligned source)
__device__ __forceinline__ void ln_fetch_row(uint32_t dst, const float* src, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(LN_ROW_BYTES) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(LN_ROW_BYTES), "r"(bar) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(LN_WARPS * 32, 2)
ln_fwd_kernel(const float* __restrict__ x, int64_t x_stride, const float* __restrict__ gamma,
              const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ mean_out,
              float* __restrict__ rstd_out, int rows, float eps) {
  trace_mark(TK_LN_FWD, 0);
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t ring = ln_smem_u32(ln_smem) + (uint32_t)warp * (LN_RING * LN_ROW_BYTES);
  const uint32_t bars = ln_smem_u32(ln_smem) + LN_WARPS * LN_RING * LN_ROW_BYTES + (uint32_t)warp * (LN_RING * 8);
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < LN_RING; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8u * s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  pdl_sync();
  trace_mark(TK_LN_FWD, 1);
  const int nw = gridDim.x * LN_WARPS;
  const int row0 = blockIdx.x * LN_WARPS + warp;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < LN_RING; ++s) {
      const int r = row0 + s * nw;
      if (r < rows) ln_fetch_row(ring + s * LN_ROW_BYTES, x + (int64_t)r * x_stride, bars + 8u * s);
    }
  }
  // gamma / beta of this lane's columns stay in registers for all rows of the warp (fetched under the first row copies)
  float4 g[LN_VEC], b[LN_VEC];
#pragma unroll
  for (int i = 0; i < LN_VEC; ++i) {
    g[i] = __ldg(reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4));
    b[i] = __ldg(reinterpret_cast<const float4*>(beta + (i * 32 + lane) * 4));
  }
  int slot = 0;
  uint32_t phase = 0;
  for (int row = row0; row < rows; row += nw) {
    ln_mbar_wait(bars + 8u * slot, phase);
    float4 v[LN_VEC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < LN_VEC; ++i) {
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w)
                   : "r"(ring + slot * LN_ROW_BYTES + (i * 32 + lane) * 16));
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
    __syncwarp();                                   // every lane has read the slot: it may be refilled
    if (lane == 0) {
      const int rn = row + LN_RING * nw;
      if (rn < rows) ln_fetch_row(ring + slot * LN_ROW_BYTES, x + (int64_t)rn * x_stride, bars + 8u * slot);
    }
    const float mean = warp_sum(s) * (1.0f / LN_COLS);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < LN_VEC; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
    const float var = warp_sum(q) * (1.0f / LN_COLS);
    const float rstd = 1.0f / sqrtf(var + eps);
    T* yr = y + (int64_t)row * LN_COLS;
#pragma unroll
    for (int i = 0; i < LN_VEC; ++i) {
      float4 o;
      o.x = v[i].x * rstd * g[i].x + b[i].x;
      o.y = v[i].y * rstd * g[i].y + b[i].y;
      o.z = v[i].z * rstd * g[i].z + b[i].z;
      o.w = v[i].w * rstd * g[i].w + b[i].w;
      Vec4IO<T>::store(yr + (i * 32 + lane) * 4, o);
    }
    if (lane == 0 && mean_out) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
    if (++slot == LN_RING) { slot = 0; phase ^= 1; }
  }
  trace_end(TK_LN_FWD);
}

// Persistent over rows.  dgamma/dbeta partials live in per-warp shared-memory accumulators (only the owning
// warp touches its slice, so no synchronisation until the end) instead of 48 registers per lane: the kernel
// then fits 2 co-resident CTAs (16 warps, ~9 KB of loads in flight per warp) per SM in one persistent wave.
// Final cross-warp sum -> one fp32 atomicAdd per column per CTA into dgamma/dbeta.
constexpr int LN_BWD_SMEM = LN_WARPS * 3 * LN_COLS * (int)sizeof(float);  // 72 KB: dgamma, dbeta, colsum(dx)

template <typename T>
__global__ void __launch_bounds__(LN_WARPS * 32, 2)
ln_bwd_kernel(const T* __restrict__ dy, const float* __restrict__ x, int64_t x_stride,
              const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
              const float* dres, float* dx, bf16* __restrict__ dx16,  // dres may alias dx (in-place residual-grad update)
              float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dx_colsum, int rows) {
  pdl_sync_traced(TK_LN_BWD);
  extern __shared__ float ln_acc[];  // [warp][3][768]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* acc_g = ln_acc + (size_t)warp * 3 * LN_COLS;
  float* acc_b = acc_g + LN_COLS;
  float* acc_c = acc_b + LN_COLS;  // column sums of the dx this kernel writes (= bias gradient of the Linear that produced x)
#pragma unroll
  for (int i = 0; i < LN_VEC; ++i) {
    *reinterpret_cast<float4*>(acc_g + (i * 32 + lane) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(acc_b + (i * 32 + lane) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(acc_c + (i * 32 + lane) * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = blockIdx.x * LN_WARPS + warp; row < rows; row += gridDim.x * LN_WARPS) {
    const float* xr = x + (int64_t)row * x_stride;
    const T* dyr = dy + (int64_t)row * LN_COLS;
    const float mu = mean[row], rs = rstd[row];
    float4 xh[LN_VEC], gy[LN_VEC], rr[LN_VEC];
    float s1 = 0.f, s2 = 0.f;
    // the residual-gradient row is fetched together with x / dy (one memory round trip per row instead of two)
#pragma unroll
    for (int i = 0; i < LN_VEC; ++i)
      rr[i] = dres ? *reinterpret_cast<const float4*>(dres + (int64_t)row * LN_COLS + (i * 32 + lane) * 4)
                   : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < LN_VEC; ++i) {
      const int c = (i * 32 + lane) * 4;
      const float4 xv = *reinterpret_cast<const float4*>(xr + c);
      const float4 d = Vec4IO<T>::load(dyr + c);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + c));
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      gy[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
      s1 += (gy[i].x + gy[i].y) + (gy[i].z + gy[i].w);
      s2 += (gy[i].x * xh[i].x + gy[i].y * xh[i].y) + (gy[i].z * xh[i].z + gy[i].w * xh[i].w);
      float4 ag = *reinterpret_cast<float4*>(acc_g + c), ab = *reinterpret_cast<float4*>(acc_b + c);
      ag.x += d.x * xh[i].x; ag.y += d.y * xh[i].y; ag.z += d.z * xh[i].z; ag.w += d.w * xh[i].w;
      ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w;
      *reinterpret_cast<float4*>(acc_g + c) = ag;
      *reinterpret_cast<float4*>(acc_b + c) = ab;
    }
    const float c1 = warp_sum(s1) * (1.0f / LN_COLS);
    const float c2 = warp_sum(s2) * (1.0f / LN_COLS);
#pragma unroll
    for (int i = 0; i < LN_VEC; ++i) {
      const int c = (i * 32 + lane) * 4;
      float4 o;
      o.x = rs * (gy[i].x - c1 - xh[i].x * c2);
      o.y = rs * (gy[i].y - c1 - xh[i].y * c2);
      o.z = rs * (gy[i].z - c1 - xh[i].z * c2);
      o.w = rs * (gy[i].w - c1 - xh[i].w * c2);
      o.x += rr[i].x; o.y += rr[i].y; o.z += rr[i].z; o.w += rr[i].w;
      *reinterpret_cast<float4*>(dx + (int64_t)row * LN_COLS + c) = o;
      if (dx16) Vec4IO<bf16>::store(dx16 + (int64_t)row * LN_COLS + c, o);
      if (dx_colsum) {
        float4 ac = *reinterpret_cast<float4*>(acc_c + c);
        ac.x += o.x; ac.y += o.y; ac.z += o.z; ac.w += o.w;
        *reinterpret_cast<float4*>(acc_c + c) = ac;
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * LN_COLS; c += LN_WARPS * 32) {
    float* dst = c < LN_COLS ? dgamma : (c < 2 * LN_COLS ? dbeta : dx_colsum);
    if (!dst) continue;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < LN_WARPS; ++w) s += ln_acc[(size_t)w * 3 * LN_COLS + c];
    atomicAdd(dst + (c % LN_COLS), s);
  }
  trace_end(TK_LN_BWD);
}

}  // namespace vitk

using namespace vitk;


extern "C" int vitk_layernorm_fwd(const float* x, int64_t x_stride, const float* gamma, const float* beta,
                                  void* y, int y_dtype, float* mean, float* rstd, int rows, float eps,
                                  void* stream) {
  VITK_CHECK_ARG(x && gamma && beta && y && rows >= 0);
  VITK_CHECK_ARG((mean == nullptr) == (rstd == nullptr));
  VITK_CHECK_ARG(x_stride % 4 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)y % 16) == 0);
  if (rows == 0) return VITK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  VITK_CHECK_ARG(((x_stride * 4) % 16) == 0);
  VITK_TRY(set_max_dyn_smem_once((const void*)ln_fwd_kernel<float>, LN_FWD_SMEM));
  VITK_TRY(set_max_dyn_smem_once((const void*)ln_fwd_kernel<bf16>, LN_FWD_SMEM));
  int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  const int cap = sm_count() * 2;   // one persistent wave: 2 CTAs per SM (96 KB of row ring each)
  if (grid > cap) grid = cap;
  if (y_dtype == VITK_F32)
    VITK_LAUNCH((ln_fwd_kernel<float>), grid, LN_WARPS * 32, LN_FWD_SMEM, st, x, x_stride, gamma, beta, (float*)y, mean, rstd, rows, eps);
  else if (y_dtype == VITK_BF16)
    VITK_LAUNCH((ln_fwd_kernel<bf16>), grid, LN_WARPS * 32, LN_FWD_SMEM, st, x, x_stride, gamma, beta, (bf16*)y, mean, rstd, rows, eps);
  else
    VITK_CHECK_ARG(!"bad dtype");
  return VITK_OK;
}

extern "C" int vitk_layernorm_bwd(const void* dy, int dy_dtype, const float* x, int64_t x_stride,
                                  const float* gamma, const float* mean, const float* rstd, const float* dres,
                                  float* dx, void* dx16, float* dgamma, float* dbeta, float* dx_colsum, int rows,
                                  void* stream) {
  VITK_CHECK_ARG(dy && x && gamma && mean && rstd && dx && rows >= 0);
  VITK_CHECK_ARG(x_stride % 4 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)dy % 16) == 0 && ((uintptr_t)dx % 16) == 0);
  if (rows == 0) return VITK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  VITK_TRY(set_max_dyn_smem_once((const void*)ln_bwd_kernel<float>, LN_BWD_SMEM));
  VITK_TRY(set_max_dyn_smem_once((const void*)ln_bwd_kernel<bf16>, LN_BWD_SMEM));
  int grid = (rows + LN_WARPS - 1) / LN_WARPS;
  const int cap = sm_count() * 2;  // 2 CTAs per SM are co-resident (launch bounds): one persistent wave
  if (grid > cap) grid = cap;
  if (dy_dtype == VITK_F32)
    VITK_LAUNCH((ln_bwd_kernel<float>), grid, LN_WARPS * 32, LN_BWD_SMEM, st, (const float*)dy, x, x_stride, gamma, mean, rstd, dres, dx, (bf16*)dx16, dgamma, dbeta, dx_colsum, rows);
  else if (dy_dtype == VITK_BF16)
    VITK_LAUNCH((ln_bwd_kernel<bf16>), grid, LN_WARPS * 32, LN_BWD_SMEM, st, (const bf16*)dy, x, x_stride, gamma, mean, rstd, dres, dx, (bf16*)dx16, dgamma, dbeta, dx_colsum, rows);
  else
    VITK_CHECK_ARG(!"bad dtype");
  return VITK_OK;
}
