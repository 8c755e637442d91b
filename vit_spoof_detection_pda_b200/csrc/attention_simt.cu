// Multi-head self-attention core for N=197, d=64 (timm Attention -> F.scaled_dot_product_attention,
// reached from /root/reference/train_advanced.py:203) -- fp32 FFMA version.
// This is the arithmetic of VITK_PREC_FP32_VALIDATE; templated on the storage type so it also
// cross-checks the tensor-core kernel in attention_mma.cu.
// One CTA per (batch, head); K/V (and Q/dO in backward) staged in shared memory as fp32 with a
// 65-float row pitch (conflict-free for both row-wise and column-wise walks); warp-shuffle softmax.
#include "common.cuh"

namespace vitk {

constexpr int AT_N = VITK_NTOK;      // 197
constexpr int AT_D = VITK_HEAD_DIM;  // 64
constexpr int AT_P = AT_D + 1;       // smem row pitch
constexpr int AT_WARPS = 8;
constexpr int AT_JP = 7;             // ceil(197/32) keys per lane
constexpr float AT_SCALE = 0.125f;   // 64^-0.5

template <typename T>
__device__ __forceinline__ void at_load_matrix(const T* __restrict__ src, float* dst) {
  // src: dense [197][64]
  for (int idx = threadIdx.x; idx < AT_N * AT_D; idx += blockDim.x) dst[(idx >> 6) * AT_P + (idx & 63)] = to_f32(src[idx]);
}

template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32)
attn_fwd_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse, int batch) {
  pdl_sync_traced(TK_ATTN_SIMT);
  extern __shared__ float smem[];
  float* Ks = smem;
  float* Vs = Ks + AT_N * AT_P;
  float* Ws = Vs + AT_N * AT_P;  // per warp: q[64] + p[224]
  const int b = blockIdx.x / VITK_HEADS, h = blockIdx.x % VITK_HEADS;
  const int64_t M = (int64_t)batch * AT_N;
  const T* qg = qkv + ((int64_t)(0 * VITK_HEADS + h) * M + (int64_t)b * AT_N) * AT_D;
  const T* kg = qkv + ((int64_t)(1 * VITK_HEADS + h) * M + (int64_t)b * AT_N) * AT_D;
  const T* vg = qkv + ((int64_t)(2 * VITK_HEADS + h) * M + (int64_t)b * AT_N) * AT_D;
  at_load_matrix(kg, Ks);
  at_load_matrix(vg, Vs);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* qs = Ws + warp * (64 + 224);
  float* ps = qs + 64;
  for (int i = warp; i < AT_N; i += AT_WARPS) {
    qs[lane] = to_f32(qg[i * AT_D + lane]);
    qs[lane + 32] = to_f32(qg[i * AT_D + lane + 32]);
    __syncwarp();
    float s[AT_JP];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < AT_JP; ++jj) {
      const int j = jj * 32 + lane;
      float a = 0.f;
      if (j < AT_N) {
        const float* kr = Ks + j * AT_P;
#pragma unroll 16
        for (int d = 0; d < AT_D; ++d) a = fmaf(qs[d], kr[d], a);
        a *= AT_SCALE;
        mx = fmaxf(mx, a);
      }
      s[jj] = a;
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < AT_JP; ++jj) {
      const int j = jj * 32 + lane;
      s[jj] = (j < AT_N) ? expf(s[jj] - mx) : 0.f;
      sum += s[jj];
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int jj = 0; jj < AT_JP; ++jj) ps[jj * 32 + lane] = s[jj] * inv;
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < AT_N; ++j) {
      const float p = ps[j];
      o0 = fmaf(p, Vs[j * AT_P + lane], o0);
      o1 = fmaf(p, Vs[j * AT_P + lane + 32], o1);
    }
    T* orow = out + ((int64_t)b * AT_N + i) * VITK_DIM + h * AT_D;
    orow[lane] = from_f32<T>(o0);
    orow[lane + 32] = from_f32<T>(o1);
    if (lane == 0 && lse) lse[(int64_t)h * M + (int64_t)b * AT_N + i] = mx + logf(sum);
    __syncwarp();
  }
  trace_end(TK_ATTN_SIMT);
}

template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32)
attn_bwd_simt_kernel(const T* __restrict__ qkv, const T* __restrict__ out, const T* __restrict__ dout,
                     const float* __restrict__ lse, T* __restrict__ dqkv, int batch) {
  pdl_sync_traced(TK_ATTN_SIMT);
  extern __shared__ float smem[];
  float* Qs = smem;
  float* Ks = Qs + AT_N * AT_P;
  float* Vs = Ks + AT_N * AT_P;
  float* dOs = Vs + AT_N * AT_P;
  float* Ls = dOs + AT_N * AT_P;   // lse[197]
  float* Ds = Ls + 200;            // delta[197]
  float* Ws = Ds + 200;            // per warp: p[224] + ds[224]
  const int b = blockIdx.x / VITK_HEADS, h = blockIdx.x % VITK_HEADS;
  const int64_t M = (int64_t)batch * AT_N;
  const int64_t hm = ((int64_t)h * M + (int64_t)b * AT_N) * AT_D;
  const int64_t hstride = (int64_t)VITK_HEADS * M * AT_D;
  at_load_matrix(qkv + hm, Qs);
  at_load_matrix(qkv + hm + hstride, Ks);
  at_load_matrix(qkv + hm + 2 * hstride, Vs);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // dO tile + delta_i = dO_i . O_i
  for (int i = warp; i < AT_N; i += AT_WARPS) {
    const int64_t g = ((int64_t)b * AT_N + i) * VITK_DIM + h * AT_D;
    const float d0 = to_f32(dout[g + lane]), d1 = to_f32(dout[g + lane + 32]);
    dOs[i * AT_P + lane] = d0;
    dOs[i * AT_P + lane + 32] = d1;
    const float dl = warp_sum(d0 * to_f32(out[g + lane]) + d1 * to_f32(out[g + lane + 32]));
    if (lane == 0) {
      Ds[i] = dl;
      Ls[i] = lse[(int64_t)h * M + (int64_t)b * AT_N + i];
    }
  }
  __syncthreads();
  float* ps = Ws + warp * 448;
  float* dss = ps + 224;
  // phase A: dQ_i = scale * sum_j dS_ij K_j
  for (int i = warp; i < AT_N; i += AT_WARPS) {
    const float* qr = Qs + i * AT_P;
    const float* dor = dOs + i * AT_P;
    const float li = Ls[i], di = Ds[i];
#pragma unroll
    for (int jj = 0; jj < AT_JP; ++jj) {
      const int j = jj * 32 + lane;
      float dsv = 0.f;
      if (j < AT_N) {
        const float* kr = Ks + j * AT_P;
        const float* vr = Vs + j * AT_P;
        float a = 0.f, dp = 0.f;
#pragma unroll 16
        for (int d = 0; d < AT_D; ++d) {
          a = fmaf(qr[d], kr[d], a);
          dp = fmaf(dor[d], vr[d], dp);
        }
        const float p = expf(a * AT_SCALE - li);
        dsv = p * (dp - di);
      }
      dss[j] = dsv;
    }
    __syncwarp();
    float a0 = 0.f, a1 = 0.f;
    for (int j = 0; j < AT_N; ++j) {
      const float dsv = dss[j];
      a0 = fmaf(dsv, Ks[j * AT_P + lane], a0);
      a1 = fmaf(dsv, Ks[j * AT_P + lane + 32], a1);
    }
    T* dq = dqkv + hm + (int64_t)i * AT_D;
    dq[lane] = from_f32<T>(a0 * AT_SCALE);
    dq[lane + 32] = from_f32<T>(a1 * AT_SCALE);
    __syncwarp();
  }
  // phase B: dV_j = sum_i P_ij dO_i ; dK_j = scale * sum_i dS_ij Q_i
  for (int j = warp; j < AT_N; j += AT_WARPS) {
    const float* kr = Ks + j * AT_P;
    const float* vr = Vs + j * AT_P;
#pragma unroll
    for (int ii = 0; ii < AT_JP; ++ii) {
      const int i = ii * 32 + lane;
      float p = 0.f, dsv = 0.f;
      if (i < AT_N) {
        const float* qr = Qs + i * AT_P;
        const float* dor = dOs + i * AT_P;
        float a = 0.f, dp = 0.f;
#pragma unroll 16
        for (int d = 0; d < AT_D; ++d) {
          a = fmaf(qr[d], kr[d], a);
          dp = fmaf(dor[d], vr[d], dp);
        }
        p = expf(a * AT_SCALE - Ls[i]);
        dsv = p * (dp - Ds[i]);
      }
      ps[i] = p;
      dss[i] = dsv;
    }
    __syncwarp();
    float v0 = 0.f, v1 = 0.f, k0 = 0.f, k1 = 0.f;
    for (int i = 0; i < AT_N; ++i) {
      const float p = ps[i], dsv = dss[i];
      v0 = fmaf(p, dOs[i * AT_P + lane], v0);
      v1 = fmaf(p, dOs[i * AT_P + lane + 32], v1);
      k0 = fmaf(dsv, Qs[i * AT_P + lane], k0);
      k1 = fmaf(dsv, Qs[i * AT_P + lane + 32], k1);
    }
    T* dk = dqkv + hm + hstride + (int64_t)j * AT_D;
    T* dv = dqkv + hm + 2 * hstride + (int64_t)j * AT_D;
    dk[lane] = from_f32<T>(k0 * AT_SCALE);
    dk[lane + 32] = from_f32<T>(k1 * AT_SCALE);
    dv[lane] = from_f32<T>(v0);
    dv[lane + 32] = from_f32<T>(v1);
    __syncwarp();
  }
  trace_end(TK_ATTN_SIMT);
}

constexpr size_t AT_FWD_SMEM = (size_t)(2 * AT_N * AT_P + AT_WARPS * (64 + 224)) * sizeof(float);
constexpr size_t AT_BWD_SMEM = (size_t)(4 * AT_N * AT_P + 400 + AT_WARPS * 448) * sizeof(float);

template <typename T>
static int attn_fwd_simt_t(const void* qkv, void* out, float* lse, int batch, cudaStream_t st) {
  VITK_TRY(set_max_dyn_smem_once((const void*)attn_fwd_simt_kernel<T>, (int)AT_FWD_SMEM));
  VITK_LAUNCH((attn_fwd_simt_kernel<T>), batch * VITK_HEADS, AT_WARPS * 32, AT_FWD_SMEM, st, (const T*)qkv, (T*)out, lse, batch);
  return VITK_OK;
}
template <typename T>
static int attn_bwd_simt_t(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int batch,
                           cudaStream_t st) {
  VITK_TRY(set_max_dyn_smem_once((const void*)attn_bwd_simt_kernel<T>, (int)AT_BWD_SMEM));
  VITK_LAUNCH((attn_bwd_simt_kernel<T>), batch * VITK_HEADS, AT_WARPS * 32, AT_BWD_SMEM, st, (const T*)qkv, (const T*)out, (const T*)dout, lse, (T*)dqkv, batch);
  return VITK_OK;
}

int attn_fwd_simt(const void* qkv, void* out, float* lse, int batch, int dtype, cudaStream_t st) {
  return dtype == VITK_BF16 ? attn_fwd_simt_t<bf16>(qkv, out, lse, batch, st) : attn_fwd_simt_t<float>(qkv, out, lse, batch, st);
}
int attn_bwd_simt(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int batch,
                  int dtype, cudaStream_t st) {
  return dtype == VITK_BF16 ? attn_bwd_simt_t<bf16>(qkv, out, dout, lse, dqkv, batch, st)
                            : attn_bwd_simt_t<float>(qkv, out, dout, lse, dqkv, batch, st);
}

}  // namespace vitk
