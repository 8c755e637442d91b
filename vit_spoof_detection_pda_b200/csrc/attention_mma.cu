// Fused multi-head self-attention forward / backward for sequence length 197, d = 64, bf16 operands on the
// tensor cores (mma.sync m16n8k16, fp32 accumulate), fp32 softmax with quad warp-shuffle reductions.
// Replaces F.scaled_dot_product_attention reached from timm Attention.forward
// (/root/reference/train_advanced.py:203 -> self.vit(x)).
//
// One CTA per (batch, head).  Q/K/V (and dO in backward) of that head are 197x64 contiguous tiles
// (head-major storage), staged once in shared memory with cp.async in a 128-byte-row XOR swizzle
// (16-byte chunk c of row r lives at chunk c ^ (r & 7)) so every ldmatrix is bank-conflict free.
// Rows 197..207 are zero padding (13 tiles of 16).
//   forward : each warp owns 16 query rows: S = Q K^T (16 x 208, registers), masked softmax, O = P V.
//   backward: phase A, per 16 query rows:  dQ = scale * [P o (dO V^T - D)] K
//             phase B, per 16 key rows:    dV = P^T dO ;  dK = scale * [P o (dO V^T - D)]^T Q
//             with P recomputed from the saved log-sum-exp; no atomics, no fp32 scratch in HBM.
// Algorithmic FLOPs per (b,h): forward 4*197*197*64; backward 14*197*197*64 (S and dP are recomputed
// in both phases).
#include "common.cuh"

namespace vitk {

constexpr int AM_N = VITK_NTOK;          // 197
constexpr int AM_NP = 208;               // padded rows
constexpr int AM_TILES = AM_NP / 16;     // 13
constexpr int AM_D = VITK_HEAD_DIM;      // 64
constexpr int AM_WARPS = 7;
constexpr int AM_THREADS = AM_WARPS * 32;
constexpr int AM_MAT_BYTES = AM_NP * 128;  // 26,624
constexpr float AM_SCALE = 0.125f;
constexpr float AM_LOG2E = 1.4426950408889634f;

int attn_debug_variant();  // gemm_tc.cu: vitk_debug_set(3, v)

__device__ __forceinline__ uint32_t am_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t am_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// stage a [197][64] bf16 matrix (row stride `ld` elements) into swizzled smem, zero the 11 pad rows
__device__ __forceinline__ void am_stage(uint32_t sbase, const bf16* __restrict__ g, int64_t ld) {
  for (int idx = threadIdx.x; idx < AM_N * 8; idx += AM_THREADS) {
    const int r = idx >> 3, c = idx & 7;
    cp_async16(sbase + am_off(r, c), g + (int64_t)r * ld + c * 8);
  }
  for (int idx = threadIdx.x; idx < (AM_NP - AM_N) * 8; idx += AM_THREADS) {
    const int r = AM_N + (idx >> 3), c = idx & 7;
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sbase + am_off(r, c)), "r"(0) : "memory");
  }
}

// A-operand fragments of a 16-row tile over the 64-wide d axis (4 k-steps)
__device__ __forceinline__ void am_load_a_tile(uint32_t sbase, int row0, int lane, uint32_t (&f)[4][4]) {
  const int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(sbase + am_off(r, ks * 2 + (lane >> 4)), f[ks]);
}

// acc[16 x 8] += A_tile[16 x 64] * Mat[n0..n0+7][0..63]^T   (B operand read non-transposed: rows of Mat are n)
__device__ __forceinline__ void am_mma_nt(float (&acc)[4], const uint32_t (&af)[4][4], uint32_t sbase, int n0, int lane) {
  const int r = n0 + (lane & 7);
#pragma unroll
  for (int kp = 0; kp < 2; ++kp) {
    uint32_t b[4];
    ldsm_x4(sbase + am_off(r, kp * 4 + (lane >> 3)), b);
    mma_bf16(acc, af[2 * kp], b[0], b[1]);
    mma_bf16(acc, af[2 * kp + 1], b[2], b[3]);
  }
}

// out[16 x 64] += A_frag[16 x 16] * Mat[k0..k0+15][0..63]        (B operand read transposed: rows of Mat are k)
__device__ __forceinline__ void am_mma_tn(float (&out)[8][4], const uint32_t (&a)[4], uint32_t sbase, int k0, int lane) {
  const int r = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
  for (int dp = 0; dp < 4; ++dp) {
    uint32_t b[4];
    ldsm_x4_t(sbase + am_off(r, dp * 2 + (lane >> 4)), b);
    mma_bf16(out[2 * dp], a, b[0], b[1]);
    mma_bf16(out[2 * dp + 1], a, b[2], b[3]);
  }
}

__global__ void __launch_bounds__(AM_THREADS)
attn_fwd_mma_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int batch) {
  pdl_sync();
  extern __shared__ __align__(1024) uint8_t am_smem[];
  const uint32_t sQ = am_smem_u32(am_smem), sK = sQ + AM_MAT_BYTES, sV = sK + AM_MAT_BYTES;
  const int b = blockIdx.x / VITK_HEADS, h = blockIdx.x % VITK_HEADS;
  const int64_t M = (int64_t)batch * AM_N;
  const int64_t hm = ((int64_t)h * M + (int64_t)b * AM_N) * AM_D;
  const int64_t hstride = (int64_t)VITK_HEADS * M * AM_D;
  am_stage(sQ, qkv + hm, AM_D);
  am_stage(sK, qkv + hm + hstride, AM_D);
  am_stage(sV, qkv + hm + 2 * hstride, AM_D);
  cp_async_wait_all();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float sl2 = AM_SCALE * AM_LOG2E;
  for (int qt = warp; qt < AM_TILES; qt += AM_WARPS) {
    uint32_t qf[4][4];
    am_load_a_tile(sQ, qt * 16, lane, qf);
    float s[26][4];
#pragma unroll
    for (int j = 0; j < 26; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      am_mma_nt(s[j], qf, sK, j * 8, lane);
    }
    // mask padded keys (>= 197) and take the row max (rows g and g+8)
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 26; ++j) {
      const int key = j * 8 + 2 * t;
      if (key >= AM_N) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
      if (key + 1 >= AM_N) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      m0 = fmaxf(m0, fmaxf(s[j][0], s[j][1]));
      m1 = fmaxf(m1, fmaxf(s[j][2], s[j][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
    const float mb0 = m0 * sl2, mb1 = m1 * sl2;
#pragma unroll
    for (int j = 0; j < 26; ++j) {
      s[j][0] = exp2f(fmaf(s[j][0], sl2, -mb0)); s[j][1] = exp2f(fmaf(s[j][1], sl2, -mb0));
      s[j][2] = exp2f(fmaf(s[j][2], sl2, -mb1)); s[j][3] = exp2f(fmaf(s[j][3], sl2, -mb1));
      l0 += s[j][0] + s[j][1];
      l1 += s[j][2] + s[j][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // O = P V
    float o[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 13; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      am_mma_tn(o, a, sV, kk * 16, lane);
    }
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    const int i0 = qt * 16 + g, i1 = i0 + 8;
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      const int d = dn * 8 + 2 * t;
      if (i0 < AM_N)
        *reinterpret_cast<uint32_t*>(out + ((int64_t)b * AM_N + i0) * VITK_DIM + h * AM_D + d) = pack_bf16x2(o[dn][0] * inv0, o[dn][1] * inv0);
      if (i1 < AM_N)
        *reinterpret_cast<uint32_t*>(out + ((int64_t)b * AM_N + i1) * VITK_DIM + h * AM_D + d) = pack_bf16x2(o[dn][2] * inv1, o[dn][3] * inv1);
    }
    if (lse && t == 0) {
      if (i0 < AM_N) lse[(int64_t)h * M + (int64_t)b * AM_N + i0] = m0 * AM_SCALE + logf(l0);
      if (i1 < AM_N) lse[(int64_t)h * M + (int64_t)b * AM_N + i1] = m1 * AM_SCALE + logf(l1);
    }
  }
}

// Column sums of one 16 x 64 output tile held in mma accumulator layout (rows g / g+8, cols 8*dn + 2t, +1), of the
// bf16-rounded scaled values actually stored, rows masked by validity: quad-stride shuffles over g, then lanes 0..3
// add into the CTA's shared accumulator.  This is the qkv bias gradient, fused instead of re-reading dqkv.
__device__ __forceinline__ void am_colsum_tile(float* cs, const float (&v)[8][4], float scale, bool valid0, bool valid1, int lane) {
  const int t = lane & 3;
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) {
    const float2 lo = unpack_bf16x2(pack_bf16x2(v[dn][0] * scale, v[dn][1] * scale));
    const float2 hi = unpack_bf16x2(pack_bf16x2(v[dn][2] * scale, v[dn][3] * scale));
    float a = (valid0 ? lo.x : 0.f) + (valid1 ? hi.x : 0.f);
    float b = (valid0 ? lo.y : 0.f) + (valid1 ? hi.y : 0.f);
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane < 4) {
      atomicAdd(cs + dn * 8 + 2 * t, a);
      atomicAdd(cs + dn * 8 + 2 * t + 1, b);
    }
  }
}

template <int MIN_CTAS>
__global__ void __launch_bounds__(AM_THREADS, MIN_CTAS)
attn_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ out, const bf16* __restrict__ dout,
                    const float* __restrict__ lse, bf16* __restrict__ dqkv, float* __restrict__ dqkv_colsum, int batch) {
  pdl_sync();
  extern __shared__ __align__(1024) uint8_t am_smem[];
  const uint32_t sQ = am_smem_u32(am_smem), sK = sQ + AM_MAT_BYTES, sV = sK + AM_MAT_BYTES, sdO = sV + AM_MAT_BYTES;
  float* Ls = reinterpret_cast<float*>(am_smem + 4 * AM_MAT_BYTES);  // lse * log2(e), padded rows 0
  float* Ds = Ls + AM_NP;                                            // delta_i = dO_i . O_i
  float* Cs = Ds + AM_NP;                                            // [3][64] column sums of dq, dk, dv (qkv bias grad)
  if (threadIdx.x < 3 * AM_D) Cs[threadIdx.x] = 0.f;
  const int b = blockIdx.x / VITK_HEADS, h = blockIdx.x % VITK_HEADS;
  const int64_t M = (int64_t)batch * AM_N;
  const int64_t hm = ((int64_t)h * M + (int64_t)b * AM_N) * AM_D;
  const int64_t hstride = (int64_t)VITK_HEADS * M * AM_D;
  const int64_t tok = ((int64_t)b * AM_N) * VITK_DIM + h * AM_D;
  am_stage(sQ, qkv + hm, AM_D);
  am_stage(sK, qkv + hm + hstride, AM_D);
  am_stage(sV, qkv + hm + 2 * hstride, AM_D);
  am_stage(sdO, dout + tok, VITK_DIM);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // delta_i = dO_i . O_i : one thread per row, 16 independent 16-byte loads in flight per thread (a
  // warp-per-row loop would serialise ~30 dependent global round trips per warp)
  if (threadIdx.x < AM_NP) {
    const int i = threadIdx.x;
    float dl = 0.f, ls = 0.f;
    if (i < AM_N) {
      const uint4* ap = reinterpret_cast<const uint4*>(dout + tok + (int64_t)i * VITK_DIM);
      const uint4* op = reinterpret_cast<const uint4*>(out + tok + (int64_t)i * VITK_DIM);
      uint4 av[8], ov[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) { av[c] = ap[c]; ov[c] = op[c]; }
      ls = lse[(int64_t)h * M + (int64_t)b * AM_N + i] * AM_LOG2E;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float2 a0 = unpack_bf16x2(av[c].x), a1 = unpack_bf16x2(av[c].y), a2 = unpack_bf16x2(av[c].z), a3 = unpack_bf16x2(av[c].w);
        const float2 o0 = unpack_bf16x2(ov[c].x), o1 = unpack_bf16x2(ov[c].y), o2 = unpack_bf16x2(ov[c].z), o3 = unpack_bf16x2(ov[c].w);
        dl += a0.x * o0.x + a0.y * o0.y + a1.x * o1.x + a1.y * o1.y + a2.x * o2.x + a2.y * o2.y + a3.x * o3.x + a3.y * o3.y;
      }
    }
    Ds[i] = dl;
    Ls[i] = ls;
  }
  cp_async_wait_all();
  __syncthreads();
  const float sl2 = AM_SCALE * AM_LOG2E;

  // ---------------- phase A: dQ for 16 query rows per warp iteration ----------------
  for (int qt = warp; qt < AM_TILES; qt += AM_WARPS) {
    uint32_t qf[4][4], dof[4][4];
    am_load_a_tile(sQ, qt * 16, lane, qf);
    am_load_a_tile(sdO, qt * 16, lane, dof);
    const float L0 = Ls[qt * 16 + g], L1 = Ls[qt * 16 + g + 8];
    const float D0 = Ds[qt * 16 + g], D1 = Ds[qt * 16 + g + 8];
    float dq[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) dq[dn][0] = dq[dn][1] = dq[dn][2] = dq[dn][3] = 0.f;
#pragma unroll 1
    for (int kb = 0; kb < AM_TILES; ++kb) {
      float s[2][4], dp[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
        am_mma_nt(s[nt], qf, sK, kb * 16 + nt * 8, lane);
        am_mma_nt(dp[nt], dof, sV, kb * 16 + nt * 8, lane);
      }
      uint32_t a[4];
      float ds[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        ds[nt][0] = exp2f(fmaf(s[nt][0], sl2, -L0)) * (dp[nt][0] - D0);
        ds[nt][1] = exp2f(fmaf(s[nt][1], sl2, -L0)) * (dp[nt][1] - D0);
        ds[nt][2] = exp2f(fmaf(s[nt][2], sl2, -L1)) * (dp[nt][2] - D1);
        ds[nt][3] = exp2f(fmaf(s[nt][3], sl2, -L1)) * (dp[nt][3] - D1);
      }
      a[0] = pack_bf16x2(ds[0][0], ds[0][1]); a[1] = pack_bf16x2(ds[0][2], ds[0][3]);
      a[2] = pack_bf16x2(ds[1][0], ds[1][1]); a[3] = pack_bf16x2(ds[1][2], ds[1][3]);
      am_mma_tn(dq, a, sK, kb * 16, lane);   // padded key rows of K are zero -> no masking needed
    }
    const int i0 = qt * 16 + g, i1 = i0 + 8;
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      const int d = dn * 8 + 2 * t;
      if (i0 < AM_N) *reinterpret_cast<uint32_t*>(dqkv + hm + (int64_t)i0 * AM_D + d) = pack_bf16x2(dq[dn][0] * AM_SCALE, dq[dn][1] * AM_SCALE);
      if (i1 < AM_N) *reinterpret_cast<uint32_t*>(dqkv + hm + (int64_t)i1 * AM_D + d) = pack_bf16x2(dq[dn][2] * AM_SCALE, dq[dn][3] * AM_SCALE);
    }
    if (dqkv_colsum) am_colsum_tile(Cs, dq, AM_SCALE, i0 < AM_N, i1 < AM_N, lane);
  }

  // ---------------- phase B: dK, dV for 16 key rows per warp iteration ----------------
  for (int kt = warp; kt < AM_TILES; kt += AM_WARPS) {
    uint32_t kf[4][4], vf[4][4];
    am_load_a_tile(sK, kt * 16, lane, kf);
    am_load_a_tile(sV, kt * 16, lane, vf);
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      dk[dn][0] = dk[dn][1] = dk[dn][2] = dk[dn][3] = 0.f;
      dv[dn][0] = dv[dn][1] = dv[dn][2] = dv[dn][3] = 0.f;
    }
#pragma unroll 1
    for (int qb = 0; qb < AM_TILES; ++qb) {
      float st[2][4], dpt[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f;
        dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f;
        am_mma_nt(st[nt], kf, sQ, qb * 16 + nt * 8, lane);    // S^T tile: rows = keys, cols = queries
        am_mma_nt(dpt[nt], vf, sdO, qb * 16 + nt * 8, lane);  // dP^T tile
      }
      uint32_t ap[4], ads[4];
      float p[2][4], ds[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int q0 = qb * 16 + nt * 8 + 2 * t;
        const float La = Ls[q0], Lb = Ls[q0 + 1], Da = Ds[q0], Db = Ds[q0 + 1];
        p[nt][0] = exp2f(fmaf(st[nt][0], sl2, -La)); p[nt][1] = exp2f(fmaf(st[nt][1], sl2, -Lb));
        p[nt][2] = exp2f(fmaf(st[nt][2], sl2, -La)); p[nt][3] = exp2f(fmaf(st[nt][3], sl2, -Lb));
        ds[nt][0] = p[nt][0] * (dpt[nt][0] - Da); ds[nt][1] = p[nt][1] * (dpt[nt][1] - Db);
        ds[nt][2] = p[nt][2] * (dpt[nt][2] - Da); ds[nt][3] = p[nt][3] * (dpt[nt][3] - Db);
      }
      ap[0] = pack_bf16x2(p[0][0], p[0][1]); ap[1] = pack_bf16x2(p[0][2], p[0][3]);
      ap[2] = pack_bf16x2(p[1][0], p[1][1]); ap[3] = pack_bf16x2(p[1][2], p[1][3]);
      ads[0] = pack_bf16x2(ds[0][0], ds[0][1]); ads[1] = pack_bf16x2(ds[0][2], ds[0][3]);
      ads[2] = pack_bf16x2(ds[1][0], ds[1][1]); ads[3] = pack_bf16x2(ds[1][2], ds[1][3]);
      am_mma_tn(dv, ap, sdO, qb * 16, lane);   // padded query rows of dO / Q are zero -> no masking needed
      am_mma_tn(dk, ads, sQ, qb * 16, lane);
    }
    const int j0 = kt * 16 + g, j1 = j0 + 8;
    bf16* dkg = dqkv + hm + hstride;
    bf16* dvg = dqkv + hm + 2 * hstride;
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      const int d = dn * 8 + 2 * t;
      if (j0 < AM_N) {
        *reinterpret_cast<uint32_t*>(dkg + (int64_t)j0 * AM_D + d) = pack_bf16x2(dk[dn][0] * AM_SCALE, dk[dn][1] * AM_SCALE);
        *reinterpret_cast<uint32_t*>(dvg + (int64_t)j0 * AM_D + d) = pack_bf16x2(dv[dn][0], dv[dn][1]);
      }
      if (j1 < AM_N) {
        *reinterpret_cast<uint32_t*>(dkg + (int64_t)j1 * AM_D + d) = pack_bf16x2(dk[dn][2] * AM_SCALE, dk[dn][3] * AM_SCALE);
        *reinterpret_cast<uint32_t*>(dvg + (int64_t)j1 * AM_D + d) = pack_bf16x2(dv[dn][2], dv[dn][3]);
      }
    }
    if (dqkv_colsum) {
      am_colsum_tile(Cs + AM_D, dk, AM_SCALE, j0 < AM_N, j1 < AM_N, lane);
      am_colsum_tile(Cs + 2 * AM_D, dv, 1.0f, j0 < AM_N, j1 < AM_N, lane);
    }
  }
  if (dqkv_colsum) {
    __syncthreads();
    if (threadIdx.x < 3 * AM_D) {
      const int sec = threadIdx.x / AM_D, d = threadIdx.x % AM_D;
      atomicAdd(dqkv_colsum + (sec * VITK_HEADS + h) * AM_D + d, Cs[threadIdx.x]);
    }
  }
}


// ================================================================================================
// v2 kernels: persistent CTAs (one per SM), 13 warps = one 16-row tile per warp, operands of the NEXT
// (batch, head) item prefetched with cp.async while the current one computes.
//
//   forward : Q/K/V double-buffered (2 x 78 KB).  Each warp: S = Q K^T in two key blocks (96 + 112 keys)
//             with an online-softmax rescale between them, so the score tile needs 56 instead of 104
//             registers and 13 warps fit one SM.
//   backward: NO recomputation.  Phase 1, warp = 16 keys:  S^T = K Q^T, dP^T = V dO^T (K, V fragments held in
//             registers), P^T / dS^T formed in registers, dV += P^T dO, dK += dS^T Q, and dS^T written
//             (bf16, [key][query], 432-byte rows: conflict-free for both the 4-byte writes and the
//             ldmatrix.trans reads) to shared memory.  Phase 2, warp = 16 queries: dQ = dS K with the A
//             operand ldmatrix.trans'ed out of the dS^T buffer.  10 N^2 d FLOPs instead of 14, half the exps.
//             delta_i = dO_i . O_i uses O rows fetched one item ahead into registers.
// ================================================================================================
constexpr int A2_WARPS = AM_TILES;              // 13
constexpr int A2_THREADS = A2_WARPS * 32;       // 416
constexpr int A2_DS_STRIDE = 432;               // bytes per key row of dS^T (208 queries x 2 B + 16 B skew)
constexpr int A2_DS_BYTES = AM_NP * A2_DS_STRIDE;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

// rows 0..196 of a [197][64] bf16 matrix -> swizzled smem, all threads of the CTA
__device__ __forceinline__ void a2_stage(uint32_t sbase, const bf16* __restrict__ g, int64_t ld) {
  for (int idx = threadIdx.x; idx < AM_N * 8; idx += A2_THREADS) {
    const int r = idx >> 3, c = idx & 7;
    cp_async16(sbase + am_off(r, c), g + (int64_t)r * ld + c * 8);
  }
}
__device__ __forceinline__ void a2_zero_pad_rows(uint32_t sbase) {
  for (int idx = threadIdx.x; idx < (AM_NP - AM_N) * 8; idx += A2_THREADS) {
    const int r = AM_N + (idx >> 3), c = idx & 7;
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sbase + am_off(r, c)), "r"(0) : "memory");
  }
}

// Store a 16 x 64 accumulator tile (rows g / g+8, cols 8*dn + 2t, +1) as bf16 rows of `ld` elements.  Lane pairs
// (t, t^1) swap halves so every lane writes 8 contiguous bytes: even t -> row g, odd t -> row g+8.
__device__ __forceinline__ void a2_store_tile(bf16* __restrict__ base, int64_t ld, int row0, int lane, const float (&v)[8][4],
                                              float s0, float s1) {
  const int g = lane >> 2, t = lane & 3;
  const bool odd = t & 1;
  const int row = row0 + g + (odd ? 8 : 0);
  const int col = 2 * (t & ~1);
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) {
    const uint32_t p0 = pack_bf16x2(v[dn][0] * s0, v[dn][1] * s0);   // row g
    const uint32_t p1 = pack_bf16x2(v[dn][2] * s1, v[dn][3] * s1);   // row g+8
    const uint32_t recv = __shfl_xor_sync(0xffffffffu, odd ? p0 : p1, 1);
    const uint2 w = odd ? make_uint2(recv, p1) : make_uint2(p0, recv);
    if (row < AM_N) *reinterpret_cast<uint2*>(base + (int64_t)row * ld + dn * 8 + col) = w;
  }
}

constexpr size_t A2_FWD_SMEM = 6 * (size_t)AM_MAT_BYTES;

__global__ void __launch_bounds__(A2_THREADS, 1)
attn_fwd_mma2_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, float* __restrict__ lse, int batch, int n_items) {
  extern __shared__ __align__(1024) uint8_t am_smem[];
  const uint32_t s0 = am_smem_u32(am_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t M = (int64_t)batch * AM_N;
  const int64_t hstride = (int64_t)VITK_HEADS * M * AM_D;
  const float sl2 = AM_SCALE * AM_LOG2E;
  for (int i = 0; i < 6; ++i) a2_zero_pad_rows(s0 + i * AM_MAT_BYTES);
  pdl_sync();
  int item = blockIdx.x;
  if (item >= n_items) return;
  auto issue = [&](int it_, int buf) {
    const int b = it_ / VITK_HEADS, h = it_ % VITK_HEADS;
    const int64_t hm = ((int64_t)h * M + (int64_t)b * AM_N) * AM_D;
    const uint32_t sb = s0 + buf * 3 * AM_MAT_BYTES;
    a2_stage(sb, qkv + hm, AM_D);
    a2_stage(sb + AM_MAT_BYTES, qkv + hm + hstride, AM_D);
    a2_stage(sb + 2 * AM_MAT_BYTES, qkv + hm + 2 * hstride, AM_D);
    cp_async_commit();
  };
  issue(item, 0);
  for (int it = 0;; ++it) {
    const int buf = it & 1;
    const int next = item + gridDim.x;
    cp_async_wait_all();
    __syncthreads();   // item's operands landed; every warp is done with the other buffer set
    if (next < n_items) issue(next, buf ^ 1);
    const uint32_t sQ = s0 + buf * 3 * AM_MAT_BYTES, sK = sQ + AM_MAT_BYTES, sV = sK + AM_MAT_BYTES;
    const int b = item / VITK_HEADS, h = item % VITK_HEADS;
    const int qt = warp;
    uint32_t qf[4][4];
    am_load_a_tile(sQ, qt * 16, lane, qf);
    float o[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    // ---- key block 0: keys 0..95 (12 n-tiles), block 1: keys 96..207 (14 n-tiles, keys >= 197 masked)
#pragma unroll
    for (int blk = 0; blk < 2; ++blk) {
      constexpr int NT0 = 12, NT1 = 14;
      const int nt_cnt = blk == 0 ? NT0 : NT1;
      const int key_base = blk == 0 ? 0 : NT0 * 8;
      float s[NT1][4];
#pragma unroll
      for (int j = 0; j < NT1; ++j) {
        if (j < nt_cnt) {
          s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
          am_mma_nt(s[j], qf, sK, key_base + j * 8, lane);
        }
      }
      float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < NT1; ++j) {
        if (j < nt_cnt) {
          if (blk == 1) {
            const int key = key_base + j * 8 + 2 * t;
            if (key >= AM_N) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
            if (key + 1 >= AM_N) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
          }
          bm0 = fmaxf(bm0, fmaxf(s[j][0], s[j][1]));
          bm1 = fmaxf(bm1, fmaxf(s[j][2], s[j][3]));
        }
      }
      bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
      bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
      const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);
      if (blk == 1) {
        const float a0 = ex2_approx((m0 - nm0) * sl2), a1 = ex2_approx((m1 - nm1) * sl2);
        l0 *= a0; l1 *= a1;
#pragma unroll
        for (int dn = 0; dn < 8; ++dn) { o[dn][0] *= a0; o[dn][1] *= a0; o[dn][2] *= a1; o[dn][3] *= a1; }
      }
      m0 = nm0; m1 = nm1;
      const float mb0 = m0 * sl2, mb1 = m1 * sl2;
      float bl0 = 0.f, bl1 = 0.f;
#pragma unroll
      for (int j = 0; j < NT1; ++j) {
        if (j < nt_cnt) {
          s[j][0] = ex2_approx(fmaf(s[j][0], sl2, -mb0)); s[j][1] = ex2_approx(fmaf(s[j][1], sl2, -mb0));
          s[j][2] = ex2_approx(fmaf(s[j][2], sl2, -mb1)); s[j][3] = ex2_approx(fmaf(s[j][3], sl2, -mb1));
          bl0 += s[j][0] + s[j][1];
          bl1 += s[j][2] + s[j][3];
        }
      }
      l0 += bl0; l1 += bl1;
#pragma unroll
      for (int kk = 0; kk < NT1 / 2; ++kk) {
        if (2 * kk < nt_cnt) {
          uint32_t a[4];
          a[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
          a[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
          a[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          a[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
          am_mma_tn(o, a, sV, key_base + kk * 16, lane);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    a2_store_tile(out + ((int64_t)b * AM_N) * VITK_DIM + h * AM_D, VITK_DIM, qt * 16, lane, o, inv0, inv1);
    if (lse && t == 0) {
      const int i0 = qt * 16 + g, i1 = i0 + 8;
      if (i0 < AM_N) lse[(int64_t)h * M + (int64_t)b * AM_N + i0] = m0 * AM_SCALE + logf(l0);
      if (i1 < AM_N) lse[(int64_t)h * M + (int64_t)b * AM_N + i1] = m1 * AM_SCALE + logf(l1);
    }
    if (next >= n_items) break;
    item = next;
  }
}

// smem: Q | dO | V | K[0] | K[1] | dS^T | Ls | Ds
constexpr size_t A2_BWD_SMEM = 5 * (size_t)AM_MAT_BYTES + A2_DS_BYTES + 2 * AM_NP * sizeof(float);

__global__ void __launch_bounds__(A2_THREADS, 1)
attn_bwd_mma2_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ out, const bf16* __restrict__ dout,
                     const float* __restrict__ lse, bf16* __restrict__ dqkv, int batch, int n_items) {
  extern __shared__ __align__(1024) uint8_t am_smem[];
  const uint32_t sQ = am_smem_u32(am_smem), sdO = sQ + AM_MAT_BYTES, sV = sdO + AM_MAT_BYTES, sK0 = sV + AM_MAT_BYTES;
  const uint32_t sDS = sK0 + 2 * AM_MAT_BYTES;
  float* Ls = reinterpret_cast<float*>(am_smem + 5 * AM_MAT_BYTES + A2_DS_BYTES);  // lse * log2(e); padded rows 0
  float* Ds = Ls + AM_NP;                                                        // delta_i = dO_i . O_i
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int64_t M = (int64_t)batch * AM_N;
  const int64_t hstride = (int64_t)VITK_HEADS * M * AM_D;
  const float sl2 = AM_SCALE * AM_LOG2E;
  for (int i = 0; i < 5; ++i) a2_zero_pad_rows(sQ + i * AM_MAT_BYTES);
  pdl_sync();
  int item = blockIdx.x;
  if (item >= n_items) return;

  uint4 ov[8];        // O row `threadIdx.x` of the item whose delta is computed next (threads < 197)
  float lsv = 0.f;
  auto fetch_o = [&](int it_) {
    if (threadIdx.x < AM_N) {
      const int b = it_ / VITK_HEADS, h = it_ % VITK_HEADS;
      const uint4* op = reinterpret_cast<const uint4*>(out + ((int64_t)b * AM_N + threadIdx.x) * VITK_DIM + h * AM_D);
#pragma unroll
      for (int c = 0; c < 8; ++c) ov[c] = __ldg(op + c);
      lsv = __ldg(lse + (int64_t)h * M + (int64_t)b * AM_N + threadIdx.x);
    }
  };
  {
    const int b = item / VITK_HEADS, h = item % VITK_HEADS;
    const int64_t hm = ((int64_t)h * M + (int64_t)b * AM_N) * AM_D;
    a2_stage(sQ, qkv + hm, AM_D);
    a2_stage(sdO, dout + ((int64_t)b * AM_N) * VITK_DIM + h * AM_D, VITK_DIM);
    a2_stage(sV, qkv + hm + 2 * hstride, AM_D);
    a2_stage(sK0, qkv + hm + hstride, AM_D);
    cp_async_commit();
    fetch_o(item);
  }
  for (int it = 0;; ++it) {
    const uint32_t sK = sK0 + (it & 1) * AM_MAT_BYTES, sKn = sK0 + ((it & 1) ^ 1) * AM_MAT_BYTES;
    const int next = item + gridDim.x;
    const bool has_next = next < n_items;
    const int b = item / VITK_HEADS, h = item % VITK_HEADS;
    const int64_t hm = ((int64_t)h * M + (int64_t)b * AM_N) * AM_D;
    int64_t hm_n = 0, tok_n = 0;
    if (has_next) {
      const int bn = next / VITK_HEADS, hn = next % VITK_HEADS;
      hm_n = ((int64_t)hn * M + (int64_t)bn * AM_N) * AM_D;
      tok_n = ((int64_t)bn * AM_N) * VITK_DIM + hn * AM_D;
    }
    cp_async_wait_all();
    __syncthreads();   // Q, dO, V, K of `item` landed; every warp finished phase 2 of the previous item

    // ---- delta / lse rows, K and V fragments of this warp's 16 keys
    if (threadIdx.x < AM_NP) {
      const int i = threadIdx.x;
      float dl = 0.f, ls = 0.f;
      if (i < AM_N) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 av;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(av.x), "=r"(av.y), "=r"(av.z), "=r"(av.w) : "r"(sdO + am_off(i, c)));
          const float2 a0 = unpack_bf16x2(av.x), a1 = unpack_bf16x2(av.y), a2 = unpack_bf16x2(av.z), a3 = unpack_bf16x2(av.w);
          const float2 o0 = unpack_bf16x2(ov[c].x), o1 = unpack_bf16x2(ov[c].y), o2 = unpack_bf16x2(ov[c].z), o3 = unpack_bf16x2(ov[c].w);
          dl += a0.x * o0.x + a0.y * o0.y + a1.x * o1.x + a1.y * o1.y + a2.x * o2.x + a2.y * o2.y + a3.x * o3.x + a3.y * o3.y;
        }
        ls = lsv * AM_LOG2E;
      }
      Ds[i] = dl;
      Ls[i] = ls;
    }
    uint32_t kf[4][4], vf[4][4];
    am_load_a_tile(sK, warp * 16, lane, kf);
    am_load_a_tile(sV, warp * 16, lane, vf);
    __syncthreads();   // Ls / Ds visible; V buffer free
    if (has_next) {
      a2_stage(sV, qkv + hm_n + 2 * hstride, AM_D);
      a2_stage(sKn, qkv + hm_n + hstride, AM_D);
    }
    cp_async_commit();

    // ---------------- phase 1: this warp's 16 keys against all queries ----------------
    {
      float dk[8][4], dv[8][4];
#pragma unroll
      for (int dn = 0; dn < 8; ++dn) {
        dk[dn][0] = dk[dn][1] = dk[dn][2] = dk[dn][3] = 0.f;
        dv[dn][0] = dv[dn][1] = dv[dn][2] = dv[dn][3] = 0.f;
      }
      const uint32_t ds_w = sDS + (uint32_t)(warp * 16 + g) * A2_DS_STRIDE + t * 4;
#pragma unroll 1
      for (int qb = 0; qb < AM_TILES; ++qb) {
        float st[2][4], dpt[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f;
          dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f;
          am_mma_nt(st[nt], kf, sQ, qb * 16 + nt * 8, lane);     // S^T tile: rows = keys, cols = queries
          am_mma_nt(dpt[nt], vf, sdO, qb * 16 + nt * 8, lane);   // dP^T tile
        }
        uint32_t ap[4], ads[4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int q0 = qb * 16 + nt * 8 + 2 * t;
          const float2 Lq = *reinterpret_cast<const float2*>(Ls + q0);
          const float2 Dq = *reinterpret_cast<const float2*>(Ds + q0);
          const float p0 = ex2_approx(fmaf(st[nt][0], sl2, -Lq.x)), p1 = ex2_approx(fmaf(st[nt][1], sl2, -Lq.y));
          const float p2 = ex2_approx(fmaf(st[nt][2], sl2, -Lq.x)), p3 = ex2_approx(fmaf(st[nt][3], sl2, -Lq.y));
          ap[2 * nt] = pack_bf16x2(p0, p1);
          ap[2 * nt + 1] = pack_bf16x2(p2, p3);
          ads[2 * nt] = pack_bf16x2(p0 * (dpt[nt][0] - Dq.x), p1 * (dpt[nt][1] - Dq.y));
          ads[2 * nt + 1] = pack_bf16x2(p2 * (dpt[nt][2] - Dq.x), p3 * (dpt[nt][3] - Dq.y));
        }
        const uint32_t dsa = ds_w + qb * 32;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dsa), "r"(ads[0]) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dsa + 8 * A2_DS_STRIDE), "r"(ads[1]) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dsa + 16), "r"(ads[2]) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dsa + 8 * A2_DS_STRIDE + 16), "r"(ads[3]) : "memory");
        am_mma_tn(dv, ap, sdO, qb * 16, lane);   // padded query rows of dO / Q are zero -> no masking needed
        am_mma_tn(dk, ads, sQ, qb * 16, lane);
      }
      a2_store_tile(dqkv + hm + hstride, AM_D, warp * 16, lane, dk, AM_SCALE, AM_SCALE);
      a2_store_tile(dqkv + hm + 2 * hstride, AM_D, warp * 16, lane, dv, 1.0f, 1.0f);
    }
    __syncthreads();   // dS^T complete; Q and dO buffers free
    if (has_next) {
      a2_stage(sQ, qkv + hm_n, AM_D);
      a2_stage(sdO, dout + tok_n, VITK_DIM);
      fetch_o(next);
    }
    cp_async_commit();

    // ---------------- phase 2: dQ for this warp's 16 queries ----------------
    {
      float dq[8][4];
#pragma unroll
      for (int dn = 0; dn < 8; ++dn) dq[dn][0] = dq[dn][1] = dq[dn][2] = dq[dn][3] = 0.f;
      const uint32_t ds_r = sDS + (uint32_t)((lane & 7) + ((lane >> 4) & 1) * 8) * A2_DS_STRIDE +
                            (uint32_t)(warp * 16 + ((lane >> 3) & 1) * 8) * 2;
#pragma unroll 1
      for (int kb = 0; kb < AM_TILES; ++kb) {
        uint32_t a[4];
        ldsm_x4_t(ds_r + kb * 16 * A2_DS_STRIDE, a);
        am_mma_tn(dq, a, sK, kb * 16, lane);   // padded key rows of K are zero -> no masking needed
      }
      a2_store_tile(dqkv + hm, AM_D, warp * 16, lane, dq, AM_SCALE, AM_SCALE);
    }
    if (!has_next) break;
    item = next;
  }
}

constexpr size_t AM_FWD_SMEM = 3 * AM_MAT_BYTES;
constexpr size_t AM_BWD_SMEM = 4 * AM_MAT_BYTES + (2 * AM_NP + 3 * AM_D) * sizeof(float);

int colsum_headmajor(const void* x, int dtype, int M, int C, float* db, cudaStream_t st);

int attn_fwd_tc(const void* qkv, void* out, float* lse, int batch, cudaStream_t st);

int attn_fwd_mma(const void* qkv, void* out, float* lse, int batch, cudaStream_t st) {
  if (attn_debug_variant() == 0) return attn_fwd_tc(qkv, out, lse, batch, st);   // default: tcgen05 kernel (attention_tc.cu)
  static bool configured = false;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(attn_fwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AM_FWD_SMEM));
    VITK_CUDA(cudaFuncSetAttribute(attn_fwd_mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_FWD_SMEM));
    configured = true;
  }
  if (attn_debug_variant() == 2) {   // debug knob 3 == 2: the first-generation kernel (one CTA per item)
    VITK_LAUNCH((attn_fwd_mma_kernel), batch * VITK_HEADS, AM_THREADS, AM_FWD_SMEM, st, (const bf16*)qkv, (bf16*)out, lse, batch);
  } else {
    const int items = batch * VITK_HEADS, sms = sm_count();
    VITK_LAUNCH((attn_fwd_mma2_kernel), (items < sms ? items : sms), A2_THREADS, A2_FWD_SMEM, st, (const bf16*)qkv, (bf16*)out, lse, batch, items);
  }
  return VITK_OK;
}

int attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int batch, cudaStream_t st);

int attn_bwd_mma(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* dqkv_colsum,
                 int batch, cudaStream_t st) {
  if (attn_debug_variant() == 0) {   // default: tcgen05 kernel (attention_tc.cu)
    VITK_TRY(attn_bwd_tc(qkv, out, dout, lse, dqkv, batch, st));
    if (dqkv_colsum) VITK_TRY(colsum_headmajor(dqkv, VITK_BF16, batch * VITK_NTOK, 3 * VITK_DIM, dqkv_colsum, st));
    return VITK_OK;
  }
  static bool configured = false;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(attn_bwd_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AM_BWD_SMEM));
    VITK_CUDA(cudaFuncSetAttribute(attn_bwd_mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_BWD_SMEM));
    configured = true;
  }
  if (attn_debug_variant() == 2) {
    VITK_LAUNCH((attn_bwd_mma_kernel<1>), batch * VITK_HEADS, AM_THREADS, AM_BWD_SMEM, st, (const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv, dqkv_colsum, batch);
    return VITK_OK;
  }
  const int items = batch * VITK_HEADS, sms = sm_count();
  VITK_LAUNCH((attn_bwd_mma2_kernel), (items < sms ? items : sms), A2_THREADS, A2_BWD_SMEM, st, (const bf16*)qkv, (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv, batch, items);
  // the qkv bias gradient: a stand-alone coalesced column-sum pass (fusing it into the mma.sync kernel measured slower)
  if (dqkv_colsum) VITK_TRY(colsum_headmajor(dqkv, VITK_BF16, batch * VITK_NTOK, 3 * VITK_DIM, dqkv_colsum, st));
  return VITK_OK;
}

}  // namespace vitk
