// mma.sync (m16n8k16) attention kernels for sequence length 197, d = 64 -- the register-tiled generation that the
// tcgen05 kernels of attention_tc.cu replaced (53 / 127 us vs 32 / 75 us per layer at batch 64).  Kept selectable
// (vitk_debug_set(3, 1)) as the A/B baseline: the legacy HMMA pipe peaks at ~550 TFLOP/s on B200 and every MMA needs a
// 256-byte B fragment from shared memory, which is exactly the 128 B/clk shared-memory port at full rate.
// Replaces F.scaled_dot_product_attention reached from timm Attention.forward
// (/root/reference/train_advanced.py:203 -> self.vit(x)).
//
// Q/K/V (and dO in backward) of one (batch, head) item are 197x64 contiguous tiles (head-major storage), staged in
// shared memory with cp.async in a 128-byte-row XOR swizzle (16-byte chunk c of row r lives at chunk c ^ (r & 7)) so
// every ldmatrix is bank-conflict free.  Rows 197..207 are zero padding (13 tiles of 16).
#include "common.cuh"

namespace vitk {

constexpr int AM_N = VITK_NTOK;          // 197
constexpr int AM_NP = 208;               // padded rows
constexpr int AM_TILES = AM_NP / 16;     // 13
constexpr int AM_D = VITK_HEAD_DIM;      // 64
constexpr int AM_MAT_BYTES = AM_NP * 128;  // 26,624
constexpr float AM_SCALE = 0.125f;
constexpr float AM_LOG2E = 1.4426950408889634f;

int attn_debug_variant();  // gemm_tc.cu: vitk_debug_set(3, v)
int debug_knob(int key);     // gemm_tc.cu: vitk_debug_set(key, v)

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

// ================================================================================================
// persistent CTAs (one per SM), 13 warps = one 16-row tile per warp, operands of the NEXT
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

int colsum_headmajor(const void* x, int dtype, int M, int C, float* db, cudaStream_t st);
int attn_fwd_tc(const void* qkv, void* out, float* lse, int batch, cudaStream_t st);
int attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* dqkv_colsum, int batch,
                cudaStream_t st, int cs_sections);

// vitk_debug_set(3, 0): tcgen05 kernels (attention_tc.cu, default); (3, 1): the mma.sync kernels of this file
int attn_fwd_mma(const void* qkv, void* out, float* lse, int batch, cudaStream_t st) {
  if (attn_debug_variant() == 0) return attn_fwd_tc(qkv, out, lse, batch, st);
  static bool configured = false;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(attn_fwd_mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_FWD_SMEM));
    configured = true;
  }
  const int items = batch * VITK_HEADS, sms = sm_count();
  VITK_LAUNCH((attn_fwd_mma2_kernel), (items < sms ? items : sms), A2_THREADS, A2_FWD_SMEM, st, (const bf16*)qkv, (bf16*)out, lse, batch, items);
  return VITK_OK;
}

int attn_bwd_mma(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* dqkv_colsum,
                 int batch, cudaStream_t st, int cs_sections) {
  if (attn_debug_variant() == 0) {
    // the tcgen05 kernel's epilogue warps add the requested sections of the qkv bias gradient from their staging tiles
    // (cs_sections: which of the q | k | v sections -- bits 0 | 1 | 2 -- this launch is asked for)
    return attn_bwd_tc(qkv, out, dout, lse, dqkv, dqkv_colsum, batch, st, cs_sections);
  } else {
    static bool configured = false;
    if (!configured) {
      VITK_CUDA(cudaFuncSetAttribute(attn_bwd_mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A2_BWD_SMEM));
      configured = true;
    }
    const int items = batch * VITK_HEADS, sms = sm_count();
    VITK_LAUNCH((attn_bwd_mma2_kernel), (items < sms ? items : sms), A2_THREADS, A2_BWD_SMEM, st, (const bf16*)qkv, (const bf16*)out,
                (const bf16*)dout, lse, (bf16*)dqkv, batch, items);
  }
  // the qkv bias gradient: a stand-alone coalesced column-sum pass
  if (dqkv_colsum) VITK_TRY(colsum_headmajor(dqkv, VITK_BF16, batch * VITK_NTOK, 3 * VITK_DIM, dqkv_colsum, st));
  return VITK_OK;
}

}  // namespace vitk
