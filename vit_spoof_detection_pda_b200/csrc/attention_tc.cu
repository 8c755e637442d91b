// Fused multi-head self-attention for sequence length 197, d = 64 on the 5th-generation tensor cores
// (tcgen05.mma, accumulators and the probability matrix in tensor memory, operands staged by TMA).
// Replaces F.scaled_dot_product_attention reached from timm Attention.forward
// (/root/reference/train_advanced.py:203 -> self.vit(x)).
//
// One persistent CTA per SM walks (batch, head) items; 320 threads:
//   warp 0      TMA producer: Q (256-row box), K, V (208-row boxes) of the NEXT item land in the other smem stage
//               while the current item computes (head-major q/k/v: one item is three dense [197][64] tiles)
//   warp 1      MMA issuer (one elected lane):
//                 S_t  = Q_t K^T        t = 0,1: query rows 128t..128t+127, N = 208 keys, K = 64   -> TMEM cols 208t..
//                 O_t  = P_t V          A operand P_t read from TENSOR MEMORY (bf16, written over S_t by the
//                                       softmax warps), B = V as an MN-major shared-memory operand, K = 208 keys
//   warps 2..9  softmax / epilogue: lane = query row (tcgen05.ld 32x32b); two passes over the 208 scores (row max, then
//               exp2 + row sum + bf16 pack -> tcgen05.st in place), later O_t * (1/l) -> global, log-sum-exp saved.
// Rows / keys 197..255 of the boxes are other items' rows (or TMA zero fill at the end of the tensor): rows only
// pollute their own (never stored) output rows, key columns >= 197 are masked to p = 0.
// Algorithmic FLOPs per item: 4 * 197^2 * 64; exps: 197 * 208 (MUFU-bound: 16/clk/SM).
#include <cuda.h>
#include <string.h>
#include <type_traits>

#include "common.cuh"

namespace vitk {

int tune_knob(int key);     // gemm_tc.cu: vitk_debug_set(key, v); key 7 (development build) = timing experiments, results invalid

namespace atc {

constexpr int N_TOK = VITK_NTOK;            // 197
constexpr int NK = 208;                     // key columns of S (multiple of 16)
constexpr uint32_t O0_COL = 2 * NK;         // O of the 128-row tile: the 64 tensor-memory columns behind the two S tiles
constexpr int THREADS = 320;
constexpr uint32_t Q_BYTES = 256 * 128;     // 32 KB  (two 128-row A tiles)
constexpr uint32_t KV_BYTES = NK * 128;     // 26 KB
constexpr uint32_t STAGE_BYTES = Q_BYTES + 2 * KV_BYTES;   // 86,016
constexpr uint32_t STG_OFF = 2 * STAGE_BYTES;                 // 8 epilogue warps x 2 KB staging tiles
constexpr uint32_t BAR_OFF = STG_OFF + 8 * 2048;
constexpr size_t FWD_SMEM = 1024 + BAR_OFF + 256;
constexpr float SCALE = 0.125f;
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                       uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %2};\n"
      "mov.b64 db, {%3, %4};\n"
      "setp.ne.b32 p, %6, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
      "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]      (A: 128 lanes x K bf16, two per 32-bit column)
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 db;\n"
      "mov.b64 db, {%2, %3};\n"
      "setp.ne.b32 p, %5, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
      "}" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tm_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tm_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tm_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Output rows leave the SM coalesced: tcgen05.ld gives every lane one ROW (32 fp32 columns); the lane packs it to 64 B of
// bf16 and writes it into a swizzled 32 x 64 B staging tile, then 4 lanes cooperate on each row, so one store
// instruction covers 8 rows x 64 contiguous bytes instead of 32 rows x 16 bytes (4x fewer cache lines per instruction:
// row-per-thread stores measured ~17 us of a 92 us backward launch).
//   stg: this warp's 2 KB staging tile; row_ptr(r): global address of column 0 of this 32-column slab in row r, or
//   nullptr when row r must not be written.
//   colsum (optional, 32 floats, 32-byte aligned): += column sums of the bf16 values of the rows that are written (the qkv
//   bias gradient), read back from the staged tile.
__device__ __forceinline__ void pack_slab32(const uint32_t (&o)[32], float scale, uint32_t (&pk)[16]) {
#pragma unroll
  for (int e = 0; e < 16; ++e) pk[e] = pack_bf16x2(__uint_as_float(o[2 * e]) * scale, __uint_as_float(o[2 * e + 1]) * scale);
}
template <typename RowPtr>
__device__ __forceinline__ void store_slab_packed(uint32_t stg, int lane, const uint32_t (&pk)[16], RowPtr row_ptr, float* colsum = nullptr) {
#pragma unroll
  for (int k4 = 0; k4 < 4; ++k4) {
    const uint32_t a = stg + (uint32_t)(lane * 64 + ((k4 ^ ((lane >> 1) & 3)) << 4));
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pk[4 * k4]), "r"(pk[4 * k4 + 1]), "r"(pk[4 * k4 + 2]),
                 "r"(pk[4 * k4 + 3]) : "memory");
  }
  __syncwarp();
  if (colsum) {
    // lane reads the 16-byte chunk (lane & 3) -- 8 columns -- of rows (lane >> 2) + 8 j: four conflict-free 128-bit loads
    // cover the tile; the 8 lanes that share a chunk are then summed with three shuffle stages and lanes 0..3 add their
    // 8 columns to global memory with two vector reductions.
    float cs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cs[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = (lane >> 2) + 8 * j;
      uint32_t w[4];
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                   : "r"(stg + (uint32_t)(r * 64 + (((lane & 3) ^ ((r >> 1) & 3)) << 4))));
      if (row_ptr(r) != nullptr) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          cs[2 * i] += __uint_as_float(w[i] << 16);
          cs[2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 4);
      cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 8);
      cs[i] += __shfl_xor_sync(0xffffffffu, cs[i], 16);
    }
    if (lane < 4) {
      float* d = colsum + lane * 8;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(cs[0]), "f"(cs[1]), "f"(cs[2]), "f"(cs[3]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d + 4), "f"(cs[4]), "f"(cs[5]), "f"(cs[6]), "f"(cs[7]) : "memory");
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = i * 8 + (lane >> 2), k4 = lane & 3;
    uint4 u;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                 : "r"(stg + (uint32_t)(r * 64 + ((k4 ^ ((r >> 1) & 3)) << 4))));
    bf16* dst = row_ptr(r);
    if (dst) *reinterpret_cast<uint4*>(dst + k4 * 8) = u;
  }
  __syncwarp();
}
template <typename RowPtr>
__device__ __forceinline__ void store_slab32(uint32_t stg, int lane, const uint32_t (&o)[32], float scale, RowPtr row_ptr,
                                             float* colsum = nullptr) {
  uint32_t pk[16];
  pack_slab32(o, scale, pk);
  store_slab_packed(stg, lane, pk, row_ptr, colsum);
}

// instruction descriptor: D = f32, A = B = bf16, majors, N >> 3, M = 128
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
// shared-memory matrix descriptor halves (SWIZZLE_128B, 8-row groups 1024 B apart)
constexpr uint32_t DESC_HI = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }

__global__ void __launch_bounds__(THREADS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                   bf16* __restrict__ out, float* __restrict__ lse, int batch, int n_items) {
  trace_mark(TK_ATTN_FWD, 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + BAR_OFF;
  // barriers: ld_full[2] | ld_empty[2] | s_full[2] | p_full[2] | o_full[2] | t_free[2] | tmem ptr
  auto ld_full = [&](int s) { return bars + 8u * s; };
  auto ld_empty = [&](int s) { return bars + 8u * (2 + s); };
  auto s_full = [&](int t) { return bars + 8u * (4 + t); };
  auto p_full = [&](int t) { return bars + 8u * (6 + t); };
  auto o_full = [&](int t) { return bars + 8u * (8 + t); };
  auto t_free = [&](int t) { return bars + 8u * (10 + t); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + BAR_OFF + 8 * 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_kv)) : "memory");
    for (int s = 0; s < 2; ++s) {
      mbar_init(ld_full(s), 1);
      mbar_init(ld_empty(s), 1);
      mbar_init(s_full(s), 1);
      mbar_init(p_full(s), 4);     // the four softmax warps of a 128-row tile
      mbar_init(o_full(s), 1);
      mbar_init(t_free(s), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();
  trace_mark(TK_ATTN_FWD, 1);

  const int64_t M = (int64_t)batch * N_TOK;
  const int n_my = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // items of this CTA

  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    for (int it = 0; it < n_my; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int b = item / VITK_HEADS, h = item % VITK_HEADS;
      const int s = it & 1;
      mbar_wait(ld_empty(s), ((it >> 1) & 1) ^ 1);
      if (leader) {
        mbar_expect_tx(ld_full(s), STAGE_BYTES);
        const uint32_t dst = sbase + s * STAGE_BYTES;
        tma_load_3d(dst, &map_q, ld_full(s), 0, b * N_TOK, h);
        tma_load_3d(dst + Q_BYTES, &map_kv, ld_full(s), 0, b * N_TOK, VITK_HEADS + h);
        tma_load_3d(dst + Q_BYTES + KV_BYTES, &map_kv, ld_full(s), 0, b * N_TOK, 2 * VITK_HEADS + h);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t IDESC_S = make_idesc(NK, 0, 0);    // S = Q K^T : both K-major, N = 208
    constexpr uint32_t IDESC_O = make_idesc(64, 0, 1);    // O = P V   : A in TMEM, B = V MN-major, N = 64
    unsigned long long* trd = lane == 0 ? trace_detail_base(TK_ATTN_FWD) : nullptr;
    for (int it = 0; it < n_my; ++it) {
      const int s = it & 1;
      const uint32_t par = it & 1;
      const uint32_t sq = sbase + s * STAGE_BYTES, sk = sq + Q_BYTES, sv = sk + KV_BYTES;
      trace_detail(trd, 8, it);              // MMA warp reaches item
      mbar_wait(ld_full(s), (it >> 1) & 1);
      trace_detail(trd, 9, it);              // operands landed
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        // S_1 overwrites the columns of the previous item's O_1; O_0 lives in the spare columns (416..479) so that S_0 of the
        // next item -- the 128-row tile is the critical chain -- does not wait for the previous O_0 to be stored
        if (t == 1) mbar_wait(t_free(1), par ^ 1);
        tc_fence_after();
        trace_detail(trd, 10 + t, it);         // S_t issued
        if (leader) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_ss(tmem_base + t * NK, desc_lo(sq + t * 16384 + k * 32), DESC_HI, desc_lo(sk + k * 32), DESC_HI, IDESC_S, k > 0 ? 1u : 0u);
          tc_commit(s_full(t));
        }
        __syncwarp();
      }
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        mbar_wait(p_full(t), par);         // P_t (bf16) is in tensor memory columns [208t, 208t + 104)
        if (t == 0) mbar_wait(t_free(0), par ^ 1);     // the previous item's O_0 has been read out of the spare columns
        tc_fence_after();
        trace_detail(trd, 12 + t, it);         // O_t issued
        if (leader) {
#pragma unroll
          for (int ks = 0; ks < NK / 16; ++ks)
            mma_ts(tmem_base + (t == 0 ? O0_COL : NK + 128), tmem_base + t * NK + ks * 8, desc_lo(sv + ks * 2048), DESC_HI, IDESC_O, ks > 0 ? 1u : 0u);
          tc_commit(o_full(t));
          if (t == 1) tc_commit(ld_empty(s));   // every MMA reading this stage has retired
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax / epilogue warps =====================
    const int we = warp - 2;
    const int t = we >> 2;               // 128-row tile of this warp
    const int qr = warp & 3;             // TMEM lane quarter this warp may touch
    const int row = qr * 32 + lane;      // row inside the tile
    const int q = t * 128 + row;         // query index
    const bool warp_valid = (t * 128 + qr * 32) < N_TOK;
    const uint32_t lane_base = tmem_base + ((uint32_t)(qr * 32) << 16);
    const uint32_t taddr = lane_base + (uint32_t)(t * NK);
    const float sl2 = SCALE * LOG2E;
    unsigned long long* trd = (qr == 0 && lane == 0) ? trace_detail_base(TK_ATTN_FWD) : nullptr;
    for (int it = 0; it < n_my; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int b = item / VITK_HEADS, h = item % VITK_HEADS;
      const uint32_t par = it & 1;
      mbar_wait(s_full(t), par);
      tc_fence_after();
      trace_detail(trd, 14 + t, it);           // S_t complete: softmax starts
      float m = -INFINITY, l = 0.f;
      if (warp_valid) {
        // ---- ONE pass over the 208 scores (tensor-memory reads are the scarce resource: ~64 B/clk/SM).  p = 2^((s - m)*c)
        // with a LAZY row maximum m: it is only raised -- and the probabilities already written rescaled -- when a chunk
        // exceeds it by more than 2^8 (p <= 256 is harmless in bf16 / fp32; 1/l and the log-sum-exp absorb the shift).
        // The next chunk's tcgen05.ld is in flight while the current one is exponentiated.
        uint32_t v[2][32];
        tm_ld32(taddr, v[0]);
#pragma unroll
        for (int c = 0; c < 7; ++c) {
          uint32_t (&cur)[32] = v[c & 1];
          tm_ld_wait();
          if (c < 5) tm_ld32(taddr + (c + 1) * 32, v[(c + 1) & 1]);
          else if (c == 5) {
            uint32_t (&nx)[32] = v[0];
            uint32_t w16[16];
            tm_ld16(taddr + 192, w16);
            tm_ld_wait();   // the last (16-column) chunk is short: fetch it synchronously
#pragma unroll
            for (int j = 0; j < 16; ++j) nx[j] = w16[j];
#pragma unroll
            for (int j = 16; j < 32; ++j) nx[j] = 0xff800000u;   // -inf
          }
          const int nval = c < 6 ? 32 : N_TOK - 192;      // valid keys of this chunk
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nval) cm = fmaxf(cm, __uint_as_float(cur[j]));
          if (c == 0) {
            m = cm;
          } else {
            const bool raise = (cm - m) * sl2 > 8.0f;
            if (__any_sync(0xffffffffu, raise)) {
              const float mn = raise ? cm : m;
              const float f = ex2((m - mn) * sl2);     // 1 for the lanes that keep their maximum
              l *= f;
              tm_st_wait();                              // earlier probability chunks have landed in tensor memory
              for (int cc = 0; cc < c; ++cc) {           // rescale the probabilities already in tensor memory
                uint32_t pk[16];
                tm_ld16(taddr + cc * 16, pk);
                tm_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float2 pr = unpack_bf16x2(pk[j]);
                  pk[j] = pack_bf16x2(pr.x * f, pr.y * f);
                }
                tm_st16(taddr + cc * 16, pk);
              }
              m = mn;
            }
          }
          const float mb = m * sl2;
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float p0 = (2 * j < nval) ? ex2(fmaf(__uint_as_float(cur[2 * j]), sl2, -mb)) : 0.f;
            const float p1 = (2 * j + 1 < nval) ? ex2(fmaf(__uint_as_float(cur[2 * j + 1]), sl2, -mb)) : 0.f;
            l += p0 + p1;
            pk[j] = pack_bf16x2(p0, p1);
          }
          if (c < 6) {
            tm_st16(taddr + c * 16, pk);
          } else {
            uint32_t pk8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) pk8[j] = pk[j];
            tm_st8(taddr + 96, pk8);
          }
        }
        tm_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(t));
      trace_detail(trd, 16 + t, it);           // P_t handed over
      // ---- O_t = P_t V is accumulated into columns [416, 480) (t = 0) / [208 + 128, 208 + 192) (t = 1)
      mbar_wait(o_full(t), par);
      tc_fence_after();
      trace_detail(trd, 18 + t, it);           // O_t complete
      if (warp_valid) {
        uint32_t o[32];
        // 1/l of every row of this warp (the staging transpose hands rows to other lanes)
        const float inv = 1.0f / l;
        const int q0w = t * 128 + qr * 32;        // first query of this warp
        bf16* obase = out + ((int64_t)b * N_TOK) * VITK_DIM + h * VITK_HEAD_DIM;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          tm_ld32((t == 0 ? lane_base + O0_COL : taddr + 128) + half * 32, o);
          tm_ld_wait();
          store_slab32(sbase + STG_OFF + (uint32_t)we * 2048, lane, o, inv,
                       [&](int r) -> bf16* { return (q0w + r < N_TOK) ? obase + (int64_t)(q0w + r) * VITK_DIM + half * 32 : nullptr; });
        }
        if (lse && q < N_TOK) lse[(int64_t)h * M + (int64_t)b * N_TOK + q] = m * SCALE + logf(l);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_free(t));
      trace_detail(trd, 20 + t, it);           // O_t stored, tensor memory released
    }
  }

  tc_fence_before();
  __syncthreads();
  trace_mark(TK_ATTN_FWD, 2);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// head-major [n_blk][M][64] bf16 -> 3-D map (64, M, n_blk), box (64, box_rows, 1), SWIZZLE_128B
static int make_hm_map(const void* base, int64_t M, int n_blk, int box_rows, CUtensorMap* map) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return VITK_ERR_DRIVER; }
  cuuint64_t dims[3] = {64, (cuuint64_t)M, (cuuint64_t)n_blk};
  cuuint64_t strides[2] = {128, (cuuint64_t)M * 128};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1}, estr[3] = {1, 1, 1};
  if ((uintptr_t)base & 15) { set_error("attention: qkv must be 16-byte aligned"); return VITK_ERR_ARG; }
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.rank = 3; key.dtype = VITK_BF16; key.swizzle = 128; key.l2promo = 128;
  for (int i = 0; i < 3; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
  key.strides[0] = strides[0]; key.strides[1] = strides[1];
  if (tmap_cache_get(key, map)) return VITK_OK;
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (attention) failed: CUresult %d", (int)r); return VITK_ERR_DRIVER; }
  tmap_cache_put(key, map);
  return VITK_OK;
}


// ================================================================================================
// backward:  dQ, dK, dV from dO with P recomputed from the saved log-sum-exp -- no recomputation of S beyond that,
// every product on tcgen05.  Key-major formulation, per item two key tiles t (128 keys) x four query chunks c
// (64, 64, 64, 16 queries) = 8 steps; per step
//     MMA1  S^T  = K_t Q_c^T          MMA2  dP^T = V_t dO_c^T            (A, B from smem, D = TMEM buffer s & 1)
//     softmax warps (lane = key): P^T = 2^(S^T*scale*log2e - L_q), dS^T = P^T o (dP^T - D_q)  -> bf16, written in place
//           into tensor memory (A operands of MMA3/4) and, dS^T only, into shared memory in the UMMA MN-major layout
//     MMA3  dV_t += P^T dO_c          MMA4  dK_t += dS^T Q_c             (A from TMEM, B = MN-major smem operand)
//     MMA5  dQ_m += dS_(m,t) K_t      after chunks 2m, 2m+1              (A = dS^T smem read MN-major, M = queries)
// Tensor memory (512 columns): 2 x (S^T 64 | dP^T 64) | dV 64 | dK 64 | dQ 2 x 64.
// Shared memory operands are single-buffered but recycled at box granularity (K/V per key tile, Q/dO per chunk): the
// producer refills a box with the next item's rows as soon as the last MMA reading it has retired, so loads overlap
// the second key tile of the current item.
// ================================================================================================
constexpr uint32_t B_SK = 0, B_SV = 32768, B_SQ = 65536, B_SDO = 98304, B_SDS = 131072;   // byte offsets
constexpr uint32_t B_SL = B_SDS + 65536, B_SD = B_SL + 2048, B_STG = B_SD + 2048, B_BAR = B_STG + 4 * 2048;
constexpr int BWD_THREADS = 64 + 8 * 32 + 4 * 32 + 2 * 32;   // producer, MMA issuer, 8 softmax warps, 4 epilogue warps, 2 delta warps
constexpr size_t BWD_SMEM = 1024 + B_BAR + 512;
constexpr uint32_t T_DV = 256, T_DK = 320, T_DQ = 384;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t saddr, uint32_t lbo) { return ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16); }

// lean MMA issue: 64-bit descriptors = (running low word, constant high word); accumulate flag known at compile time
template <bool ACC>
__device__ __forceinline__ void mma_ss_c(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %3};\n"
      "setp.ne.b32 p, %5, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
      "}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "n"(ACC ? 1 : 0) : "memory");
}
template <bool ACC>
__device__ __forceinline__ void mma_ts_c(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t hi, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 db;\n"
      "mov.b64 db, {%2, %3};\n"
      "setp.ne.b32 p, %5, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
      "}" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(hi), "r"(idesc), "n"(ACC ? 1 : 0) : "memory");
}
// The MMAs of one backward step, step index S = 4 t + c known at compile time (key tile t, query chunk c, buffer S & 1).
struct BwdMma {
  uint32_t tmem, hi, kK, kV, kQ, kdO, kDS;     // descriptor low words of the operand bases (byte offsets add as (off >> 4))
  template <int S, int K = 0>
  __device__ __forceinline__ void front() {    // S^T = K_t Q_c^T, dP^T = V_t dO_c^T (two interleaved accumulation chains)
    constexpr int t = S >> 2, c = S & 3, buf = S & 1;
    constexpr uint32_t idesc = c < 3 ? make_idesc(64, 0, 0) : make_idesc(16, 0, 0);
    mma_ss_c<(K > 0)>(tmem + buf * 128, kK + ((t * 16384 + K * 32) >> 4), kQ + ((c * 8192 + K * 32) >> 4), hi, idesc);
    mma_ss_c<(K > 0)>(tmem + buf * 128 + 64, kV + ((t * 16384 + K * 32) >> 4), kdO + ((c * 8192 + K * 32) >> 4), hi, idesc);
    if constexpr (K < 3) front<S, K + 1>();
  }
  template <int S, int KS = 0>
  __device__ __forceinline__ void back_vk() {  // dV_t += P^T dO_c, dK_t += dS^T Q_c (A from tensor memory)
    constexpr int c = S & 3, buf = S & 1, nks = c < 3 ? 4 : 1;
    constexpr uint32_t idesc = make_idesc(64, 0, 1);
    mma_ts_c<(c > 0 || KS > 0)>(tmem + T_DV, tmem + buf * 128 + KS * 8, kdO + ((c * 8192 + KS * 2048) >> 4), hi, idesc);
    mma_ts_c<(c > 0 || KS > 0)>(tmem + T_DK, tmem + buf * 128 + 64 + KS * 8, kQ + ((c * 8192 + KS * 2048) >> 4), hi, idesc);
    if constexpr (KS + 1 < nks) back_vk<S, KS + 1>();
  }
  template <int S, int KS = 0>
  __device__ __forceinline__ void back_q() {   // dQ_m += dS_(m,t) K_t (A = dS^T in shared memory, read MN-major)
    constexpr int t = S >> 2, m = (S & 3) >> 1;
    constexpr uint32_t idesc = make_idesc(64, 1, 1);
    mma_ss_c<(t > 0 || KS > 0)>(tmem + T_DQ + m * 64, kDS + ((m * 32768 + KS * 2048) >> 4), kK + ((t * 16384 + KS * 2048) >> 4), hi, idesc);
    if constexpr (KS < 7) back_q<S, KS + 1>();
  }
};

__global__ void __launch_bounds__(BWD_THREADS, 1)   // 512 threads x 128 registers = the whole register file
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ CUtensorMap map_q,
                   const __grid_constant__ CUtensorMap map_do, const bf16* __restrict__ out, const bf16* __restrict__ dout_g,
                   const float* __restrict__ lse,
                   bf16* __restrict__ dqkv, float* __restrict__ dqkv_colsum, int cs_sections, int batch, int n_items, int dbg_in) {
#ifdef VITK_DEV
  const int dbg = dbg_in;      // timing experiments (bit 0 no MMAs, 1 no softmax math, 2 no draining, 3 no delta, 4 no loads)
#else
  constexpr int dbg = 0;       // the release build has no result-invalidating path
  (void)dbg_in;
#endif
  trace_mark(TK_ATTN_BWD, 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + B_BAR;
  // barriers (8 B each)
  auto kv_full = [&](int t) { return bars + 8u * t; };          // 0,1
  auto kv_free = [&](int t) { return bars + 8u * (2 + t); };    // 2,3
  auto qd_full = [&](int c) { return bars + 8u * (4 + c); };    // 4..7
  auto qd_free = [&](int c) { return bars + 8u * (8 + c); };    // 8..11
  auto st_full = [&](int b) { return bars + 8u * (12 + b); };   // 12,13
  auto p_full = [&](int b) { return bars + 8u * (14 + b); };    // 14,15
  auto ds_free = [&](int m) { return bars + 8u * (16 + m); };   // 16,17
  const uint32_t dvk_full = bars + 8u * 18, dvk_free = bars + 8u * 19, dq_full = bars + 8u * 20, dq_free = bars + 8u * 21;
  auto delta_ready = [&](int par, int c) { return bars + 8u * (22 + par * 4 + c); };   // 22..29: item parity x chunk
  auto ld_free = [&](int par) { return bars + 8u * (30 + par); };   // 30,31: the L / delta buffer of that item parity has been consumed
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + B_BAR + 8 * 32);
  // L_q = lse_q * log2(e) and delta_q = dO_q . O_q, double-buffered by item parity: [2][256] floats each
  float* sL = reinterpret_cast<float*>(smem + B_SL);
  float* sD = reinterpret_cast<float*>(smem + B_SD);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_kv)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_do)) : "memory");
    for (int i = 0; i < 2; ++i) { mbar_init(kv_full(i), 1); mbar_init(kv_free(i), 1); mbar_init(st_full(i), 1); mbar_init(p_full(i), 4); mbar_init(ds_free(i), 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(qd_full(i), 1); mbar_init(qd_free(i), 1); mbar_init(delta_ready(0, i), 2); mbar_init(delta_ready(1, i), 2); }
    mbar_init(dvk_full, 1); mbar_init(dvk_free, 4); mbar_init(dq_full, 1); mbar_init(dq_free, 4);
    mbar_init(ld_free(0), 8); mbar_init(ld_free(1), 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();
  trace_mark(TK_ATTN_BWD, 1);

  const int64_t M = (int64_t)batch * N_TOK;
  const int n_my = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    for (int it = 0; it < n_my; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int b = item / VITK_HEADS, h = item % VITK_HEADS;
      const uint32_t fpar = (it & 1) ^ 1;     // the box was released once per previous item
      auto load_kv = [&](int t) {
        mbar_wait(kv_free(t), fpar);
        if (leader) {
          if (dbg & 16) { mbar_arrive(kv_full(t)); }
          else {
            mbar_expect_tx(kv_full(t), 32768);
            tma_load_3d(sbase + B_SK + t * 16384, &map_kv, kv_full(t), 0, b * N_TOK + t * 128, VITK_HEADS + h);
            tma_load_3d(sbase + B_SV + t * 16384, &map_kv, kv_full(t), 0, b * N_TOK + t * 128, 2 * VITK_HEADS + h);
          }
        }
        __syncwarp();
      };
      load_kv(0);
      for (int c = 0; c < 4; ++c) {
        mbar_wait(qd_free(c), fpar);
        if (leader) {
          if (dbg & 16) { mbar_arrive(qd_full(c)); }
          else {
            mbar_expect_tx(qd_full(c), 16384);
            tma_load_3d(sbase + B_SQ + c * 8192, &map_q, qd_full(c), 0, b * N_TOK + c * 64, h);
            tma_load_2d(sbase + B_SDO + c * 8192, &map_do, qd_full(c), h * VITK_HEAD_DIM, b * N_TOK + c * 64);
          }
        }
        __syncwarp();
      }
      load_kv(1);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // One elected lane issues ~20 MMAs per step; what it costs per MMA in instructions IS the step time (measured with the
    // kernel-internal trace marks of the development build: with run-time step indices -- multiplies, shifts and branches per
    // descriptor -- the issue thread needed ~4,000 clocks per step while the softmax warps needed ~2,300 and the tensor
    // pipe ~650).  The eight steps of an item are therefore UNROLLED AT COMPILE TIME (template step index): every
    // descriptor low word is a base register plus an immediate, the accumulate predicates are constants.
    const bool leader = elect_one();
    const bool do_mma = !(dbg & 1);
    unsigned long long* trd = lane == 0 ? trace_detail_base(TK_ATTN_BWD) : nullptr;
    BwdMma mm;
    mm.tmem = tmem_base;
    mm.hi = DESC_HI;
    mm.kK = desc_lo(sbase + B_SK); mm.kV = desc_lo(sbase + B_SV); mm.kQ = desc_lo(sbase + B_SQ); mm.kdO = desc_lo(sbase + B_SDO);
    mm.kDS = desc_lo_mn(sbase + B_SDS, 16384);
    for (int it = 0; it < n_my; ++it) {
      const uint32_t ipar = it & 1;
      // front<S>: MMA1, MMA2 of step S;  back<S>: MMA3, MMA4 (and MMA5 after odd chunks) of step S
      auto front = [&](auto S) {
        constexpr int s = decltype(S)::value, t = s >> 2, c = s & 3, buf = s & 1;
        if (c == 0) mbar_wait(kv_full(t), ipar);
        if (t == 0) mbar_wait(qd_full(c), ipar);
        tc_fence_after();
        trace_detail(trd, 8, it * 8 + s);      // front(s): operands landed, MMA1/2 issued now
        if (leader) {
          if (do_mma) mm.front<s>();
          tc_commit(st_full(buf));
        }
        __syncwarp();
      };
      auto back = [&](auto S) {
        constexpr int s = decltype(S)::value, t = s >> 2, c = s & 3, buf = s & 1;
        const int n_b = it * 4 + (s >> 1);                 // completions of this buffer's barriers so far
        mbar_wait(p_full(buf), n_b & 1);
        trace_detail(trd, 9, it * 8 + s);      // back(s): P^T / dS^T ready
        if (c == 0) mbar_wait(dvk_free, ((it * 2 + t) & 1) ^ 1);   // the previous dV / dK tile has been read out
        tc_fence_after();
        trace_detail(trd, 10, it * 8 + s);     // back(s): accumulators free, MMA3/4(/5) issued now
        if (leader) {
          if (do_mma) mm.back_vk<s>();
          if (t == 1) tc_commit(qd_free(c));       // Q_c / dO_c no longer needed by this item
          if (c == 3) tc_commit(dvk_full);         // dV_t, dK_t complete
        }
        __syncwarp();
        if (c & 1) {
          constexpr int m = c >> 1;
          if (t == 0) mbar_wait(dq_free, ipar ^ 1);          // the previous item's dQ has been read out
          tc_fence_after();
          if (leader) {
            if (do_mma) mm.back_q<s>();
            tc_commit(ds_free(m));
            if (m == 1) tc_commit(kv_free(t));     // K_t / V_t no longer needed
            if (t == 1 && m == 1) tc_commit(dq_full);
          }
          __syncwarp();
        }
      };
      using std::integral_constant;
      front(integral_constant<int, 0>{});
      front(integral_constant<int, 1>{}); back(integral_constant<int, 0>{});
      front(integral_constant<int, 2>{}); back(integral_constant<int, 1>{});
      front(integral_constant<int, 3>{}); back(integral_constant<int, 2>{});
      front(integral_constant<int, 4>{}); back(integral_constant<int, 3>{});
      front(integral_constant<int, 5>{}); back(integral_constant<int, 4>{});
      front(integral_constant<int, 6>{}); back(integral_constant<int, 5>{});
      front(integral_constant<int, 7>{}); back(integral_constant<int, 6>{});
      back(integral_constant<int, 7>{});
    }
  } else if (warp < 10) {
    // ===================== softmax warps: two groups of four, ping-pong over the steps =====================
    // Group g = (warp - 2) / 4 owns the steps with s & 1 == g, i.e. tensor-memory buffer g: while one group is between its
    // tcgen05.ld and its arrive, the other group's MMAs (back of the previous step, front of the next) and its own
    // softmax run -- the step pipeline is two deep in the softmax stage as well, not only in tensor memory.  (With all
    // eight warps on ONE step the warps sat in the ld -> exp -> st -> fence chain of that step while the tensor pipe and
    // the other buffer idled: 4,000 clocks per step measured, for ~600 issue slots of work.)  A lane owns one key row
    // and walks the 64 query columns of the chunk in two halves, half 0 first: the bf16 P^T / dS^T it writes in place
    // (16 packed columns per half) then only ever cover fp32 columns it has already read.
    const int we = warp - 2;
    const int qr = warp & 3;             // TMEM lane quarter
    const int grp = we >> 2;             // step parity / tensor-memory buffer of this group
    const int row = qr * 32 + lane;      // TMEM lane = key inside the 128-key tile
    const float sl2 = SCALE * LOG2E;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qr * 32) << 16);
    const uint32_t sDS = sbase + B_SDS;
    const int buf = grp;
    unsigned long long* trd = (qr == 0 && lane == 0) ? trace_detail_base(TK_ATTN_BWD) : nullptr;
    for (int it = 0; it < n_my; ++it) {
      const uint32_t ipar = it & 1;
      const uint32_t sLb = sbase + B_SL + ipar * 1024, sDb = sbase + B_SD + ipar * 1024;   // this item's L / delta
#pragma unroll 1
      for (int s = grp; s < 8; s += 2) {
        const int t = s >> 2, c = s & 3, m = c >> 1;
        const int n_b = it * 4 + (s >> 1);
        const int key = t * 128 + row;
        trace_detail(trd, 11 + grp, it * 8 + s);   // softmax group: waiting for step s
        mbar_wait(st_full(buf), n_b & 1);      // MMA1/2 of this step retired
        trace_detail(trd, 13 + grp, it * 8 + s);   // S^T / dP^T of step s complete
        mbar_wait(ds_free(m), ((it * 2 + t) & 1) ^ 1);                     // MMA5 of the previous key tile has read dS^T buffer m
        if (t == 0) mbar_wait(delta_ready(ipar, c), (it >> 1) & 1);        // L / delta of this chunk's queries (delta warps)
        tc_fence_after();
        trace_detail(trd, 15 + grp, it * 8 + s);   // all inputs of step s ready: math starts
        const uint32_t ds_row = sDS + m * 32768 + (c & 1) * 16384 + row * 128;
        if (dbg & 2) {
        } else if (c < 3) {
#pragma unroll 1
          for (int ch = 0; ch < 2; ++ch) {
            const uint32_t a_st = lane_addr + buf * 128 + ch * 32, a_dp = a_st + 64;
            uint32_t vs[32], vd[32];
            tm_ld32(a_st, vs);
            tm_ld32(a_dp, vd);
            const int q0 = c * 64 + ch * 32;
            tm_ld_wait();
            uint32_t pp[16], pd[16];
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              // per-query constants of 4 columns (broadcast shared-memory reads, fetched as they are needed: holding all
              // 64 of them next to the 64 accumulator words spills at 128 registers per thread)
              float Lv[4], Dv[4];
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(Lv[0]), "=f"(Lv[1]), "=f"(Lv[2]), "=f"(Lv[3])
                           : "r"(sLb + (q0 + j4 * 4) * 4));
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(Dv[0]), "=f"(Dv[1]), "=f"(Dv[2]), "=f"(Dv[3])
                           : "r"(sDb + (q0 + j4 * 4) * 4));
              float p[4], d[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                p[e] = key < N_TOK ? ex2(fmaf(__uint_as_float(vs[j4 * 4 + e]), sl2, -Lv[e])) : 0.f;
                d[e] = p[e] * (__uint_as_float(vd[j4 * 4 + e]) - Dv[e]);
              }
              pp[j4 * 2] = pack_bf16x2(p[0], p[1]); pp[j4 * 2 + 1] = pack_bf16x2(p[2], p[3]);
              pd[j4 * 2] = pack_bf16x2(d[0], d[1]); pd[j4 * 2 + 1] = pack_bf16x2(d[2], d[3]);
            }
            tm_st16(lane_addr + buf * 128 + ch * 16, pp);
            tm_st16(lane_addr + buf * 128 + 64 + ch * 16, pd);
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_row + (((ch * 4 + g4) ^ (row & 7)) << 4)),
                           "r"(pd[g4 * 4]), "r"(pd[g4 * 4 + 1]), "r"(pd[g4 * 4 + 2]), "r"(pd[g4 * 4 + 3]) : "memory");
          }
          tm_st_wait();
        } else {
          // last chunk: queries 192..207 (16 columns), valid up to 196
          uint32_t vs[16], vd[16];
          tm_ld16(lane_addr + buf * 128, vs);
          tm_ld16(lane_addr + buf * 128 + 64, vd);
          float Lq[16], Dq[16];
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(Lq[j4 * 4]), "=f"(Lq[j4 * 4 + 1]), "=f"(Lq[j4 * 4 + 2]), "=f"(Lq[j4 * 4 + 3])
                         : "r"(sLb + (192 + j4 * 4) * 4));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(Dq[j4 * 4]), "=f"(Dq[j4 * 4 + 1]), "=f"(Dq[j4 * 4 + 2]), "=f"(Dq[j4 * 4 + 3])
                         : "r"(sDb + (192 + j4 * 4) * 4));
          }
          tm_ld_wait();
          uint32_t pp[8], pd[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float p[2], d[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int qi = 2 * j + e;
              p[e] = (key < N_TOK && 192 + qi < N_TOK) ? ex2(fmaf(__uint_as_float(vs[qi]), sl2, -Lq[qi])) : 0.f;
              d[e] = p[e] * (__uint_as_float(vd[qi]) - Dq[qi]);
            }
            pp[j] = pack_bf16x2(p[0], p[1]);
            pd[j] = pack_bf16x2(d[0], d[1]);
          }
          tm_st8(lane_addr + buf * 128, pp);
          tm_st8(lane_addr + buf * 128 + 64, pd);
#pragma unroll
          for (int g4 = 0; g4 < 2; ++g4)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ds_row + ((g4 ^ (row & 7)) << 4)),
                         "r"(pd[g4 * 4]), "r"(pd[g4 * 4 + 1]), "r"(pd[g4 * 4 + 2]), "r"(pd[g4 * 4 + 3]) : "memory");
          tm_st_wait();
        }
        fence_proxy_async_smem();     // dS^T rows in shared memory are read by the tensor core (async proxy)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(buf));
        trace_detail(trd, 17 + grp, it * 8 + s);   // step s handed to the MMA issuer
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(ld_free(ipar));   // this warp is done with the item's L / delta values
    }
  } else if (warp < 14) {
    // ===================== epilogue warps (10..13): dV / dK after each key tile, dQ after the item =====================
    // The accumulators are single-buffered in tensor memory and gate the next key tile / item (dvk_free, dq_free), so each
    // drain first pulls its accumulator into registers (packed to bf16), hands the tensor memory back, and only then sends
    // the rows through the staging tile to global memory.  (Measured with the kernel-internal marks: when these warps also
    // computed delta for the next item -- three global-memory round trips per item -- the next item's first back() and its
    // first dQ MMA stalled 3 + 9 us per item behind them.)
    const int ew = warp - 10;            // 0..3
    const int qr = warp & 3;             // TMEM lane quarter this warp may touch
    const uint32_t lane_addr = tmem_base + ((uint32_t)(qr * 32) << 16);
    const uint32_t stg = sbase + B_STG + (uint32_t)ew * 2048;
    const int64_t hstride = (int64_t)VITK_HEADS * M * 64;
    unsigned long long* trd = (ew == 0 && lane == 0) ? trace_detail_base(TK_ATTN_BWD) : nullptr;
    for (int it = 0; it < n_my; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int b = item / VITK_HEADS, h = item % VITK_HEADS;
      const uint32_t ipar = it & 1;
      const int64_t hm = ((int64_t)h * M + (int64_t)b * N_TOK) * 64;
      float* csb = dqkv_colsum ? dqkv_colsum + h * 64 : nullptr;      // [3][12][64]: q | k | v sections (cs_sections: bits 0 | 1 | 2)
      for (int t = 0; t < 2; ++t) {
        // ---- dV_t, dK_t (lane = key)
        mbar_wait(dvk_full, (it * 2 + t) & 1);
        tc_fence_after();
        trace_detail(trd, 19, it * 2 + t);        // dV / dK of key tile t complete
        const bool rows_exist = t * 128 + qr * 32 < N_TOK;
        uint32_t pv[2][16], pkk[2][16];
        if (rows_exist && !(dbg & 4)) {
          uint32_t o[32];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            tm_ld32(lane_addr + T_DV + half * 32, o);
            tm_ld_wait();
            pack_slab32(o, 1.0f, pv[half]);
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            tm_ld32(lane_addr + T_DK + half * 32, o);
            tm_ld_wait();
            pack_slab32(o, SCALE, pkk[half]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(dvk_free);
        trace_detail(trd, 20, it * 2 + t);        // dV / dK pulled out of tensor memory
        if (rows_exist && !(dbg & 4)) {
          const int r0 = t * 128 + qr * 32;
          bf16* dv = dqkv + hm + 2 * hstride;
          bf16* dk = dqkv + hm + hstride;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            store_slab_packed(stg, lane, pv[half],
                              [&](int r) -> bf16* { return (r0 + r < N_TOK) ? dv + (int64_t)(r0 + r) * 64 + half * 32 : nullptr; },
                              (csb && (cs_sections & 4)) ? csb + 2 * VITK_DIM + half * 32 : nullptr);
            store_slab_packed(stg, lane, pkk[half],
                              [&](int r) -> bf16* { return (r0 + r < N_TOK) ? dk + (int64_t)(r0 + r) * 64 + half * 32 : nullptr; },
                              (csb && (cs_sections & 2)) ? csb + VITK_DIM + half * 32 : nullptr);
          }
        }
      }
      // ---- dQ (lane = query), two 128-query tiles: same order -- registers first, tensor memory released, then the stores
      trace_detail(trd, 21, it);                  // dV / dK rows stored; waiting for dQ
      mbar_wait(dq_full, ipar);
      tc_fence_after();
      trace_detail(trd, 22, it);
      uint32_t pq[2][2][16];
      const bool q_exist[2] = {true, 128 + qr * 32 < N_TOK};
      if (!(dbg & 4)) {
        uint32_t o[32];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (!q_exist[mt]) continue;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            tm_ld32(lane_addr + T_DQ + mt * 64 + half * 32, o);
            tm_ld_wait();
            pack_slab32(o, SCALE, pq[mt][half]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      trace_detail(trd, 23, it);
      if (!(dbg & 4)) {
        float* csq = (csb && (cs_sections & 1)) ? csb : nullptr;
        bf16* dq = dqkv + hm;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (!q_exist[mt]) continue;
          const int r0 = mt * 128 + qr * 32;
#pragma unroll
          for (int half = 0; half < 2; ++half)
            store_slab_packed(stg, lane, pq[mt][half],
                              [&](int r) -> bf16* { return (r0 + r < N_TOK) ? dq + (int64_t)(r0 + r) * 64 + half * 32 : nullptr; },
                              csq ? csq + half * 32 : nullptr);
        }
      }
      trace_detail(trd, 24, it);                  // dQ rows stored
    }
  } else {
    // ===================== delta warps (14, 15): delta_q = dO_q . O_q and L_q = lse_q log2(e), ONE ITEM AHEAD =====================
    // Both operands are read from global memory (L2: dO was just written by the proj dgrad), not from the dO boxes in shared
    // memory -- those are recycled per chunk only when the previous item's second key tile retires, so anything that waits for
    // them is late for the next item's first key tile (measured: 1.5 - 2.7 us per step on three of its four steps).
    // Every warp takes 32 of the 64 queries of a chunk, 8 lanes per query.  L / delta are double-buffered by item parity
    // ([2][256] floats, one ready barrier per parity and chunk, one free barrier per parity on which every softmax warp
    // arrives when it has finished an item): these warps run up to two items ahead and never a barrier phase too far.
    const int dw = warp - 14;            // 0..1
    for (int it = 0; it < n_my; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int b = item / VITK_HEADS, h = item % VITK_HEADS;
      const int64_t rowbase = ((int64_t)b * N_TOK) * VITK_DIM + h * VITK_HEAD_DIM + (lane & 7) * 8;
      float* dD = sD + (it & 1) * 256;
      float* dL = sL + (it & 1) * 256;
      if (it >= 2) mbar_wait(ld_free(it & 1), ((it >> 1) - 1) & 1);      // item it - 2 no longer reads this buffer
      if (dbg & 8) {
        for (int c = 0; c < 4; ++c) { __syncwarp(); if (lane == 0) mbar_arrive(delta_ready(it & 1, c)); }
        continue;
      }
      // 7 half-chunks of 16 queries per warp (chunk 3 holds 16 queries: one half).  The 8 row fetches of half-chunk k + 1 are
      // in flight while half-chunk k is reduced: these warps are bound by the latency of their global loads (O comes from
      // HBM), and only when they outrun the item rate do they get -- and stay -- ahead of the softmax warps.
      uint4 ov[2][4], av[2][4];
      auto fetch = [&](int hc, int slot) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int q = (hc >> 1) * 64 + dw * 32 + ((hc & 1) * 4 + g) * 4 + (lane >> 3);
          const int64_t off = rowbase + (int64_t)(q < N_TOK ? q : 0) * VITK_DIM;
          ov[slot][g] = __ldg(reinterpret_cast<const uint4*>(out + off));
          av[slot][g] = __ldg(reinterpret_cast<const uint4*>(dout_g + off));
        }
      };
      fetch(0, 0);
#pragma unroll
      for (int hc = 0; hc < 7; ++hc) {
        const int c = hc >> 1, slot = hc & 1;
        if (hc + 1 < 7) fetch(hc + 1, slot ^ 1);
        const int ql = c * 64 + dw * 32 + lane;     // the log-sum-exp of this warp's 32 queries of the chunk
        float lv = 0.f;
        if ((hc & 1) == 0 && ql < N_TOK) lv = __ldg(lse + (int64_t)h * M + (int64_t)b * N_TOK + ql);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int q = c * 64 + dw * 32 + ((hc & 1) * 4 + g) * 4 + (lane >> 3);
          const uint4 a = av[slot][g], o = ov[slot][g];
          const float2 a0 = unpack_bf16x2(a.x), a1 = unpack_bf16x2(a.y), a2 = unpack_bf16x2(a.z), a3 = unpack_bf16x2(a.w);
          const float2 o0 = unpack_bf16x2(o.x), o1 = unpack_bf16x2(o.y), o2 = unpack_bf16x2(o.z), o3 = unpack_bf16x2(o.w);
          float d = a0.x * o0.x + a0.y * o0.y + a1.x * o1.x + a1.y * o1.y + a2.x * o2.x + a2.y * o2.y + a3.x * o3.x + a3.y * o3.y;
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          if ((lane & 7) == 0) dD[q] = q < N_TOK ? d : 0.f;
        }
        if ((hc & 1) == 0) dL[ql] = lv * LOG2E;
        if ((hc & 1) == 1 || hc == 6) {
          __syncwarp();
          if (lane == 0) mbar_arrive(delta_ready(it & 1, c));     // release: the stores above are visible to the waiting softmax warps
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  trace_mark(TK_ATTN_BWD, 2);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// dout [M][768] bf16 -> 2-D map (768, M), box (64, 64), SWIZZLE_128B
static int make_dout_map(const void* base, int64_t M, CUtensorMap* map) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return VITK_ERR_DRIVER; }
  cuuint64_t dims[2] = {(cuuint64_t)VITK_DIM, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)VITK_DIM * 2};
  cuuint32_t box[2] = {64, 64}, estr[2] = {1, 1};
  if ((uintptr_t)base & 15) { set_error("attention: dout must be 16-byte aligned"); return VITK_ERR_ARG; }
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.rank = 2; key.dtype = VITK_BF16; key.swizzle = 128; key.l2promo = 129;
  key.dims[0] = dims[0]; key.dims[1] = dims[1]; key.box[0] = box[0]; key.box[1] = box[1];
  key.strides[0] = strides[0];
  if (tmap_cache_get(key, map)) return VITK_OK;
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (attention dout) failed: CUresult %d", (int)r); return VITK_ERR_DRIVER; }
  tmap_cache_put(key, map);
  return VITK_OK;
}

}  // namespace atc

int attn_fwd_tc(const void* qkv, void* out, float* lse, int batch, cudaStream_t st) {
  VITK_TRY(set_max_dyn_smem_once((const void*)atc::attn_fwd_tc_kernel, (int)atc::FWD_SMEM));
  const int64_t M = (int64_t)batch * VITK_NTOK;
  CUtensorMap map_q, map_kv;
  VITK_TRY(atc::make_hm_map(qkv, M, 3 * VITK_HEADS, 256, &map_q));
  VITK_TRY(atc::make_hm_map(qkv, M, 3 * VITK_HEADS, atc::NK, &map_kv));
  const int items = batch * VITK_HEADS, sms = sm_count();
  VITK_LAUNCH((atc::attn_fwd_tc_kernel), (items < sms ? items : sms), atc::THREADS, atc::FWD_SMEM, st, map_q, map_kv, (bf16*)out, lse, batch, items);
  return VITK_OK;
}

int attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* dqkv_colsum, int batch,
                cudaStream_t st, int cs_sections) {
  VITK_TRY(set_max_dyn_smem_once((const void*)atc::attn_bwd_tc_kernel, (int)atc::BWD_SMEM));
  const int64_t M = (int64_t)batch * VITK_NTOK;
  CUtensorMap map_kv, map_q, map_do;
  VITK_TRY(atc::make_hm_map(qkv, M, 3 * VITK_HEADS, 128, &map_kv));
  VITK_TRY(atc::make_hm_map(qkv, M, 3 * VITK_HEADS, 64, &map_q));
  VITK_TRY(atc::make_dout_map(dout, M, &map_do));
  const int items = batch * VITK_HEADS, sms = sm_count();
  VITK_LAUNCH((atc::attn_bwd_tc_kernel), (items < sms ? items : sms), atc::BWD_THREADS, atc::BWD_SMEM, st, map_kv, map_q, map_do,
              (const bf16*)out, (const bf16*)dout, lse, (bf16*)dqkv, dqkv_colsum, cs_sections, batch, items, tune_knob(7));
  return VITK_OK;
}

}  // namespace vitk
