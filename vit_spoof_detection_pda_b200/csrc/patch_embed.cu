// Patch embedding (timm PatchEmbed.proj = Conv2d(3, 768, k16, s16) + cls token + pos_embed; SURVEY.md 2.1 K1/K2,
// reached from /root/reference/train_advanced.py:190/203) as an IM2COL-FREE GEMM FED BY TMA.
//
// The patch matrix  P[(b, py, px)][k = c*256 + i*16 + j] = image[b][c][16 py + i][16 px + j]  is never materialised in
// global memory: a 4-D tensor map over the NCHW fp32 image -- dims (x 224, i 16, py 14, b*3+c) with image row y = 16 py + i,
// strides (4, 896, 14336, 200704) bytes -- lets ONE TMA box {224, NI, NPY, 1} deliver NI image rows of each of NPY patch
// rows into shared memory as [py][i][224 pixels]: the K-slice (c, i0..i0+NI) of 14 * NPY patches.  Four
// converter warps turn each raw box into a bf16 operand tile in the canonical 128-byte-swizzled layout ([row][64 k] =
// 128 B rows, 16-byte chunk c at c ^ (row & 7)) -- byte for byte what a TMA load of a materialised bf16 patch matrix
// would have produced -- and tcgen05.mma.kind::f16 consumes it next to the bf16 weights / gradients that arrive by TMA.
//   forward   x0[b, 1+p, :] = P[b, p, :] . Wpe^T + bpe + pos[1+p] ;  x0[b, 0, :] = cls + pos[0]
//             CTA = (image b, half h, 256 output columns): 128 accumulator rows = patches py in [9h, 9h+9) (126 rows; the box
//             of h = 1 runs past py = 13 and is zero-filled); 12 k-blocks of 64 in (image-row group, channel) order.
//   uint8 edge: the same kernel, raw boxes {672 B = 224 px x 3 ch, 4 rows, 9, 1} of the HWC uint8 image;
//             the converter applies ToTensor + Normalize ((u / 255 - mean[c]) / std[c], fp32, torchvision's order:
//             train_advanced.py:174-175) before the bf16 cast: the same values, in the same k order, as the fp32 edge.
//   wgrad     dWpe[n][k] += sum_tokens dx0[token][n] * P[token][k]: both operands MN-major (token = reduction = slow axis
//             of both arrays).  A = bf16 dx0 by TMA (3-D map [b][t][768], 64-token boxes), B = raw boxes {224, 16, 4, 1}
//             (56 patches x the 256 k of one channel) through the converter (rows 56..63 of the tile are zero);
//             accumulator 128 (n) x 256 (k) in tensor memory, images split over CTAs, partial tiles summed with
//             red.global.add.v4.f32.
// The fp32-validate / SIMT path reads the image through MatLayout modes 3 / 4 (gemm.cuh): no im2col there either.
#include <cuda.h>
#include <string.h>

#include "gemm.cuh"

namespace vitk {
namespace pe {

constexpr int BN = 256;                    // output columns per CTA (forward) / k columns per CTA (wgrad)
constexpr int N_KB = VITK_DIM / 64;        // 12 k-blocks of 64
constexpr int STAGES = 3;
constexpr uint32_t A_BYTES = 128 * 128;    // bf16 operand tile of the patches: 128 rows x 64 k = 16 KB
constexpr uint32_t B_BYTES = BN * 128;     // bf16 weights: 256 rows x 64 k = 32 KB
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;          // 48 KB
constexpr uint32_t RAW_BYTES = 32768;      // fp32 box {224, 4, 9, 1}: 36 image rows x 896 B = 32,256; uint8 box: 36 x 672 B = 24,192
constexpr uint32_t RAW_OFF = STAGES * STAGE_BYTES;           // 147,456
constexpr uint32_t BAR_OFF = RAW_OFF + 2 * RAW_BYTES;        // 212,992
constexpr size_t FWD_SMEM = 1024 + BAR_OFF + 256;
constexpr int NCW = 8;                     // converter / epilogue warps (two per tensor-memory lane quarter)
constexpr int FWD_THREADS = 64 + 32 * NCW; // warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 converter + epilogue
constexpr int ROWS_H0 = 126, ROWS_H1 = 70;   // patches of the two halves of an image (9 and 5 patch rows)

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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
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
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// instruction descriptor: D = f32, A = B = bf16, majors, N >> 3, M = 128
__host__ __device__ constexpr uint32_t idesc_bf16(int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
// shared-memory matrix descriptor, SWIZZLE_128B: 8-row groups 1024 B apart; LBO = distance of the 64-wide MN atoms (MN-major)
constexpr uint32_t DESC_HI = ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}

// ---- converter: 64 k values (4 image rows x 16 pixels of one channel) of one patch -> 128 bytes of bf16 in a SWIZZLE_128B tile --
// The raw box holds WHOLE image rows -- [patch row py][image row i][224 pixels] (fp32: 896 B rows; uint8 HWC: 672 B rows) --
// because TMA moves a box as one request per contiguous inner row: boxes cut into 16-pixel rows (64 B) arrived at ~12
// clocks per row and were the kernel's bottleneck (504 rows per k-block), whole image rows are 36.
// Task = (patch r = py * 14 + px, chunk ch of 4 pixels: image row i = ch / 4, pixels 4 (ch % 4) ..); a warp takes 2 patches x 16
// chunks per pass.
//   fp32 raw : 4 floats at ((py * n_i + i0 + i) * 224 + px * 16 + (ch % 4) * 4) * 4
//   uint8 raw: 12 bytes (4 pixels x 3 channels) at (py * 4 + i) * 672 + px * 48 + (ch % 4) * 12; channel c is picked out
template <bool U8>
__device__ __forceinline__ void convert_rows(uint32_t raw, int n_i, int i0, uint32_t tile, int n_rows, int n_tile_rows, int cw,
                                             int lane, int c, float mean, float sd) {
  const int ch = lane & 15;
  const int r_first = cw * 2 + (lane >> 4);     // NCW warps x 2 rows per pass
  // rows r_first + 8 k, four at a time: the four shared-memory fetches are issued before any of them is used (a
  // one-row-per-iteration loop was bound by the load -> convert -> store latency: 160 clocks per row, 2,500 per k-block)
#pragma unroll 1
  for (int rb = r_first; rb < n_tile_rows; rb += 8 * NCW) {
    uint32_t w[4][3];
    float4 f[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = rb + 2 * NCW * u;
      const int rr = r < n_rows ? r : 0;
      const int py = rr / 14, px = rr - py * 14;
      if (U8) {
        const uint32_t src = raw + (uint32_t)((py * 4 + (ch >> 2)) * 672 + px * 48 + (ch & 3) * 12);
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[u][0]) : "r"(src));
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[u][1]) : "r"(src + 4));
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w[u][2]) : "r"(src + 8));
      } else {
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[u].x), "=f"(f[u].y), "=f"(f[u].z), "=f"(f[u].w)
                     : "r"(raw + (uint32_t)(((py * n_i + i0 + (ch >> 2)) * 224 + px * 16 + (ch & 3) * 4) * 4)));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = rb + 2 * NCW * u;
      if (r >= n_tile_rows) break;
      uint32_t lo, hi;
      if (U8) {
        const uint64_t lo64 = (uint64_t)w[u][0] | ((uint64_t)w[u][1] << 32);
        float v[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          const int byte = x * 3 + c;     // 0..11
          const uint32_t uu = byte < 8 ? (uint32_t)((lo64 >> (8 * byte)) & 0xffu) : ((w[u][2] >> (8 * (byte - 8))) & 0xffu);
          v[x] = ((float)uu / 255.0f - mean) / sd;
        }
        lo = pack_bf16x2(v[0], v[1]);
        hi = pack_bf16x2(v[2], v[3]);
      } else {
        lo = pack_bf16x2(f[u].x, f[u].y);
        hi = pack_bf16x2(f[u].z, f[u].w);
      }
      if (r >= n_rows) { lo = 0u; hi = 0u; }     // rows past n_rows: zeros
      const uint32_t dst = tile + (uint32_t)r * 128 + ((((uint32_t)ch >> 1) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)ch & 1u) * 8;
      asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst), "r"(lo), "r"(hi) : "memory");
    }
  }
}

struct FwdParams {
  const float* bpe;
  const float* cls;
  const float* pos;
  float* x0;
  int batch;
  float mean[3], stdv[3];
};

template <bool U8>
__global__ void __launch_bounds__(FWD_THREADS, 1)
patch_embed_fwd_kernel(const __grid_constant__ CUtensorMap map_img, const __grid_constant__ CUtensorMap map_w, const FwdParams p) {
  trace_mark(TK_PATCH_EMBED, 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + BAR_OFF;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto raw_full = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto raw_empty = [&](int s) { return bars + 8u * (2 * STAGES + 2 + s); };
  const uint32_t acc_full = bars + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + BAR_OFF + 8 * (2 * STAGES + 5));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // work item: 256 output columns x one half of one image
  const int tn = blockIdx.x % (VITK_DIM / BN), half = (blockIdx.x / (VITK_DIM / BN)) & 1, b = blockIdx.x / (2 * (VITK_DIM / BN));
  const int py0 = half * 9;
  // k-block q covers k = c*256 + ig*64 .. +64 with ig = q / 3 (image rows 4 ig .. 4 ig + 3), c = q % 3: the uint8 raw box of
  // an image-row group holds all three channels; the fp32 edge walks the same order (bit-identical accumulators)
  constexpr int RAW_PER_KB = U8 ? 3 : 1;                     // k-blocks fed by one raw box
  constexpr uint32_t RAW_TX = U8 ? 672u * 4u * 9u : 224u * 4u * 4u * 9u;      // whole image rows: [9 py][4 i][224 px (x 3 ch | x 4 B)]

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_img)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_w)) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1 + NCW);        // producer's expect_tx arrive (weights) + the converter warps (patches)
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) { mbar_init(raw_full(s), 1); mbar_init(raw_empty(s), NCW); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();
  trace_mark(TK_PATCH_EMBED, 1);

  if (warp == 0) {
    // ===================== TMA producer: raw image boxes + weight tiles =====================
    const bool leader = elect_one();
    unsigned long long* trd = lane == 0 ? trace_detail_base(TK_PATCH_EMBED) : nullptr;
    for (int q = 0; q < N_KB; ++q) {
      const int s = q % STAGES, ig = q / 3, c = q % 3;
      trace_detail(trd, 9, q);                  // producer reaches k-block q
      if (q % RAW_PER_KB == 0) {
        const int rq = q / RAW_PER_KB, rs = rq & 1;
        mbar_wait(raw_empty(rs), ((rq >> 1) & 1) ^ 1);
        if (leader) {
          mbar_expect_tx(raw_full(rs), RAW_TX);
          tma_load_4d(sbase + RAW_OFF + rs * RAW_BYTES, &map_img, raw_full(rs), 0, ig * 4, py0, U8 ? b : b * 3 + c);
        }
      }
      mbar_wait(empty_bar(s), ((q / STAGES) & 1) ^ 1);
      trace_detail(trd, 10, q);                 // stage free: weights issued
      if (leader) {
        mbar_expect_tx(full_bar(s), B_BYTES);
        tma_load_2d(sbase + s * STAGE_BYTES + A_BYTES, &map_w, full_bar(s), c * 256 + ig * 64, tn * BN);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool leader = elect_one();
    constexpr uint32_t IDESC = idesc_bf16(BN, 0, 0);
    unsigned long long* trd = lane == 0 ? trace_detail_base(TK_PATCH_EMBED) : nullptr;
    for (int q = 0; q < N_KB; ++q) {
      const int s = q % STAGES;
      mbar_wait(full_bar(s), (q / STAGES) & 1);
      tc_fence_after();
      trace_detail(trd, 11, q);                 // operands of k-block q complete: MMAs issued
      if (leader) {
        const uint32_t sa = sbase + s * STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)        // UMMA_K = 16 bf16 = 32 B inside the 128-byte swizzle row
          mma_bf16(tmem_base, desc_lo(sa + ks * 32, 16), DESC_HI, desc_lo(sa + A_BYTES + ks * 32, 16), DESC_HI, IDESC,
                   (q > 0 || ks > 0) ? 1u : 0u);
        tc_commit(empty_bar(s));
        if (q + 1 == N_KB) tc_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    const int cw = warp - 2;                 // 0..3
    // ===================== converter: raw rows -> bf16 operand tiles =====================
    unsigned long long* trd = (cw == 0 && lane == 0) ? trace_detail_base(TK_PATCH_EMBED) : nullptr;
    for (int q = 0; q < N_KB; ++q) {
      const int s = q % STAGES, c = q % 3;
      const int rq = q / RAW_PER_KB, rs = rq & 1;
      if (q % RAW_PER_KB == 0) mbar_wait(raw_full(rs), (rq >> 1) & 1);
      trace_detail(trd, 12, q);                 // raw box of k-block q landed
      mbar_wait(empty_bar(s), ((q / STAGES) & 1) ^ 1);
      trace_detail(trd, 13, q);                 // stage free: conversion starts
      convert_rows<U8>(sbase + RAW_OFF + rs * RAW_BYTES, 4, 0, sbase + s * STAGE_BYTES, ROWS_H0, 128, cw, lane, c, p.mean[c], p.stdv[c]);
      fence_proxy_async_smem();              // generic-proxy writes -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(full_bar(s));
        if (q % RAW_PER_KB == RAW_PER_KB - 1) mbar_arrive(raw_empty(rs));   // the raw box has been consumed
      }
      trace_detail(trd, 14, q);                 // conversion of k-block q done
    }
    // ===================== epilogue: + bias + pos_embed, token rows of this half; the CLS row =====================
    // tcgen05.ld hands a lane one accumulator ROW; the 32 x 32 chunk goes through a swizzled transpose buffer (the raw ring,
    // idle by now) so that 8 lanes cover 128 contiguous bytes of ONE token row: coalesced pos reads and x0 stores.
    const int q4 = warp & 3;                 // TMEM lane quarter this warp may touch
    const int chs = cw >> 2;                 // which interleaved half of the eight 32-column chunks
    const int nvalid = half ? ROWS_H1 : ROWS_H0;
    const uint32_t tb = sbase + RAW_OFF + (uint32_t)cw * 4096;
    const int sub_row = lane >> 3, c4 = lane & 7;
    const int t_base = 1 + half * ROWS_H0 + q4 * 32;       // token of this warp's accumulator row 0
    // pos_embed / bias of the NEXT chunk are fetched while the current one is written (those of the first chunk under the
    // end of the main loop): fetched inside the chunk they cost one global-memory round trip per chunk.  The loop is kept
    // rolled: this code runs once per CTA, straight-line copies of it only add instruction-cache misses.
    float4 pn[8], bn;
    auto fetch_pos = [&](int ch) {
      const int col = tn * BN + ch * 32 + c4 * 4;
      bn = __ldg(reinterpret_cast<const float4*>(p.bpe + col));
#pragma unroll
      for (int itr = 0; itr < 8; ++itr) {
        const int rr = itr * 4 + sub_row;
        const int t = (q4 * 32 + rr < nvalid) ? t_base + rr : 0;
        pn[itr] = __ldg(reinterpret_cast<const float4*>(p.pos + (int64_t)t * VITK_DIM + col));
      }
    };
    fetch_pos(chs);
    mbar_wait(acc_full, 0);
    tc_fence_after();
    trace_detail(trd, 8);                        // main loop done, epilogue starts
#pragma unroll 1
    for (int ch = chs; ch < BN / 32; ch += 2) {
      uint32_t v[32];
      tm_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + ch * 32, v);
      float4 pc[8];
      const float4 b4 = bn;
#pragma unroll
      for (int itr = 0; itr < 8; ++itr) pc[itr] = pn[itr];
      if (ch + 2 < BN / 32) fetch_pos(ch + 2);
#pragma unroll
      for (int g = 0; g < 8; ++g)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tb + (uint32_t)(lane * 128 + ((g ^ (lane & 7)) << 4))),
                     "r"(v[g * 4]), "r"(v[g * 4 + 1]), "r"(v[g * 4 + 2]), "r"(v[g * 4 + 3]) : "memory");
      __syncwarp();
      const int col = tn * BN + ch * 32 + c4 * 4;
#pragma unroll
      for (int itr = 0; itr < 8; ++itr) {
        const int rr = itr * 4 + sub_row;
        float4 a;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w)
                     : "r"(tb + (uint32_t)(rr * 128 + ((c4 ^ (rr & 7)) << 4))));
        if (q4 * 32 + rr < nvalid) {
          const int t = t_base + rr;
          *reinterpret_cast<float4*>(p.x0 + ((int64_t)b * VITK_NTOK + t) * VITK_DIM + col) =
              make_float4(a.x + b4.x + pc[itr].x, a.y + b4.y + pc[itr].y, a.z + b4.z + pc[itr].z, a.w + b4.w + pc[itr].w);
        }
      }
      __syncwarp();
    }
    trace_detail(trd, 15);                       // epilogue rows stored
    if (half == 0 && cw == 0) {              // x0[b, 0, :] = cls + pos[0]
      float* crow = p.x0 + (int64_t)b * VITK_NTOK * VITK_DIM + tn * BN;
      for (int j = lane * 4; j < BN; j += 128) {
        const float4 cc = __ldg(reinterpret_cast<const float4*>(p.cls + tn * BN + j));
        const float4 pp = __ldg(reinterpret_cast<const float4*>(p.pos + tn * BN + j));
        *reinterpret_cast<float4*>(crow + j) = make_float4(cc.x + pp.x, cc.y + pp.y, cc.z + pp.z, cc.w + pp.w);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  trace_mark(TK_PATCH_EMBED, 2);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
constexpr int WG_ROWS = 56;                                  // patches per reduction chunk: 4 patch rows (py) x 14
constexpr uint32_t WG_A_BYTES = 2 * 8192;                    // bf16 dx0: 2 MN atoms (64 features) x 64 tokens x 128 B
constexpr uint32_t WG_B_BYTES = 4 * 8192;                    // bf16 patches: 4 MN atoms (64 k) x 64 tokens x 128 B
constexpr uint32_t WG_STAGE = WG_A_BYTES + WG_B_BYTES;       // 48 KB
constexpr int WG_STAGES = 2;
constexpr uint32_t WG_RAW_BYTES = WG_ROWS * 1024;            // fp32 box {224, 16, 4, 1}: 64 image rows x 896 B = 57,344 B
constexpr uint32_t WG_RAW_OFF = WG_STAGES * WG_STAGE;        // 98,304
constexpr uint32_t WG_BAR_OFF = WG_RAW_OFF + 2 * WG_RAW_BYTES;   // 212,992
constexpr size_t WG_SMEM = 1024 + WG_BAR_OFF + 128;
constexpr int WG_THREADS = 64 + 32 * NCW;
constexpr int WG_TILES = (VITK_DIM / 128) * 3;               // 6 n tiles x 3 channels

__global__ void __launch_bounds__(WG_THREADS, 1)
patch_embed_wgrad_kernel(const __grid_constant__ CUtensorMap map_img, const __grid_constant__ CUtensorMap map_dx,
                         float* __restrict__ dw, int batch, int n_splits) {
  trace_mark(TK_PATCH_WGRAD, 0);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + WG_BAR_OFF;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (WG_STAGES + s); };
  auto raw_full = [&](int s) { return bars + 8u * (2 * WG_STAGES + s); };
  auto raw_empty = [&](int s) { return bars + 8u * (2 * WG_STAGES + 2 + s); };
  const uint32_t acc_full = bars + 8u * (2 * WG_STAGES + 4);
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + WG_BAR_OFF + 8 * (2 * WG_STAGES + 5));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x % WG_TILES, split = blockIdx.x / WG_TILES;
  const int tn = tile / 3, c = tile % 3;                     // output features [128 tn, +128) x k columns of channel c
  const int n_img = (batch - split + n_splits - 1) / n_splits;   // images split, split + n_splits, ...
  const int n_iters = n_img * 4;                             // 4 chunks of 4 patch rows per image (the last: 2 real + 2 zero-filled)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_img)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_dx)) : "memory");
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(full_bar(s), 1 + NCW); mbar_init(empty_bar(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(raw_full(s), 1); mbar_init(raw_empty(s), NCW); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_sync();
  trace_mark(TK_PATCH_WGRAD, 1);

  if (warp == 0) {
    const bool leader = elect_one();
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % WG_STAGES, rs = it & 1;
      const int b = split + (it >> 2) * n_splits, py0 = (it & 3) * 4;
      mbar_wait(raw_empty(rs), ((it >> 1) & 1) ^ 1);
      if (leader) {
        mbar_expect_tx(raw_full(rs), WG_RAW_BYTES);
        tma_load_4d(sbase + WG_RAW_OFF + rs * WG_RAW_BYTES, &map_img, raw_full(rs), 0, 0, py0, b * 3 + c);
      }
      mbar_wait(empty_bar(s), ((it / WG_STAGES) & 1) ^ 1);
      if (leader) {
        mbar_expect_tx(full_bar(s), WG_A_BYTES);
        const uint32_t sa = sbase + s * WG_STAGE;
        // bf16 dx0[b][1 + 14 py0 .. +64][128 tn + 64 a .. +64]: tokens past 196 are zero-filled; the 8 tokens past the chunk's
        // 56 meet zero rows of the patch tile
#pragma unroll
        for (int a = 0; a < 2; ++a) tma_load_3d(sa + a * 8192, &map_dx, full_bar(s), tn * 128 + a * 64, 1 + py0 * 14, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t IDESC = idesc_bf16(BN, 1, 1);          // both operands MN-major
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % WG_STAGES;
      mbar_wait(full_bar(s), (it / WG_STAGES) & 1);
      tc_fence_after();
      if (leader) {
        const uint32_t sa = sbase + s * WG_STAGE, sb = sa + WG_A_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)                        // UMMA_K = 16 tokens = 16 rows x 128 B
          mma_bf16(tmem_base, desc_lo(sa + ks * 2048, 8192), DESC_HI, desc_lo(sb + ks * 2048, 8192), DESC_HI, IDESC,
                   (it > 0 || ks > 0) ? 1u : 0u);
        tc_commit(empty_bar(s));
        if (it + 1 == n_iters) tc_commit(acc_full);
      }
      __syncwarp();
    }
  } else {
    const int cw = warp - 2;
    // converter: raw [56 patches][256 k fp32] -> four [64 tokens][64 k] bf16 tiles (MN-major B operand; rows 56..63 zero)
    for (int it = 0; it < n_iters; ++it) {
      const int s = it % WG_STAGES, rs = it & 1;
      mbar_wait(raw_full(rs), (it >> 1) & 1);
      mbar_wait(empty_bar(s), ((it / WG_STAGES) & 1) ^ 1);
      const uint32_t sb = sbase + s * WG_STAGE + WG_A_BYTES;
#pragma unroll 1
      for (int ig = 0; ig < 4; ++ig)
        convert_rows<false>(sbase + WG_RAW_OFF + rs * WG_RAW_BYTES, 16, ig * 4, sb + ig * 8192, WG_ROWS, 64, cw, lane, 0, 0.f, 1.f);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(full_bar(s));
        mbar_arrive(raw_empty(rs));
      }
    }
    if (n_iters > 0) {
      // epilogue: lane = output feature n, 256 consecutive k columns -> red.global.add.v4.f32 into the fp32 gradient
      const int q4 = warp & 3;
      mbar_wait(acc_full, 0);
      tc_fence_after();
      float* drow = dw + (int64_t)(tn * 128 + q4 * 32 + lane) * VITK_DIM + c * 256;
#pragma unroll 1
      for (int ch = cw >> 2; ch < BN / 32; ch += 2) {
        uint32_t v[32];
        tm_ld32(tmem_base + ((uint32_t)(q4 * 32) << 16) + ch * 32, v);
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + ch * 32 + c4 * 4), "f"(__uint_as_float(v[c4 * 4 + 0])),
                       "f"(__uint_as_float(v[c4 * 4 + 1])), "f"(__uint_as_float(v[c4 * 4 + 2])), "f"(__uint_as_float(v[c4 * 4 + 3])) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  trace_mark(TK_PATCH_WGRAD, 2);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
  }
}

// uint8 HWC -> fp32 NCHW with ToTensor + Normalize (fp32-validate / SIMT path and the weight gradient of uint8 inputs only:
// the bf16 forward never materialises this)
__global__ void __launch_bounds__(256)
u8_to_nchw_kernel(const uint8_t* __restrict__ img, float m0, float m1, float m2, float s0, float s1, float s2,
                  float* __restrict__ out, int batch) {
  pdl_sync_traced(TK_PATCH_EMBED);
  const int64_t total = (int64_t)batch * VITK_IMG * VITK_IMG;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = idx / (VITK_IMG * VITK_IMG), px = idx % (VITK_IMG * VITK_IMG);
    const uint8_t* src = img + idx * 3;
    float* dst = out + b * 3 * VITK_IMG * VITK_IMG + px;
    dst[0] = ((float)src[0] / 255.0f - m0) / s0;
    dst[(int64_t)VITK_IMG * VITK_IMG] = ((float)src[1] / 255.0f - m1) / s1;
    dst[(int64_t)2 * VITK_IMG * VITK_IMG] = ((float)src[2] / 255.0f - m2) / s2;
  }
  trace_end(TK_PATCH_EMBED);
}

// dpos[t][j] += sum_b dx0[b][t][j]; dcls[j] += that at t = 0; dbpe[j] += sum over t >= 1.  grid (197, batch groups)
constexpr int EG_BGROUP = 8;
__global__ void __launch_bounds__(192)
embed_param_grads_kernel(const float* __restrict__ dx0, int batch, float* __restrict__ dpos, float* __restrict__ dcls,
                         float* __restrict__ dbpe) {
  pdl_sync_traced(TK_EMBED_GRADS);
  const int t = blockIdx.x, b0 = blockIdx.y * EG_BGROUP, b1 = min(batch, b0 + EG_BGROUP);
  const int j = threadIdx.x * 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int b = b0; b < b1; ++b) {
    const float4 v = *reinterpret_cast<const float4*>(dx0 + ((int64_t)b * VITK_NTOK + t) * VITK_DIM + j);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dpos + (int64_t)t * VITK_DIM + j), "f"(s.x), "f"(s.y), "f"(s.z), "f"(s.w) : "memory");
  float* other = t == 0 ? dcls : dbpe;
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(other + j), "f"(s.x), "f"(s.y), "f"(s.z), "f"(s.w) : "memory");
  trace_end(TK_EMBED_GRADS);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
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

static int encode(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, CUtensorMapSwizzle sw, const char* what) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return VITK_ERR_DRIVER; }
  if ((uintptr_t)base & 15) { set_error("patch embedding: %s must be 16-byte aligned", what); return VITK_ERR_ARG; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (%s) failed: CUresult %d", what, (int)r); return VITK_ERR_DRIVER; }
  return VITK_OK;
}
// NCHW fp32 image as (x 224, i 16, py 14, b*3 + c) -- image row y = 16 py + i; box {224, box_i, box_py, 1}: whole image rows,
// no swizzle (the converter warps read it)
static int image_map_f32(const float* img, int batch, int box_i, int box_py, CUtensorMap* map) {
  const cuuint64_t dims[4] = {VITK_IMG, 16, 14, (cuuint64_t)batch * 3};
  const cuuint64_t strides[3] = {VITK_IMG * 4, 16 * VITK_IMG * 4, (cuuint64_t)VITK_IMG * VITK_IMG * 4};
  const cuuint32_t box[4] = {VITK_IMG, (cuuint32_t)box_i, (cuuint32_t)box_py, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, img, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, "image");
}
// HWC uint8 image as (168 words = 224 px x 3 ch, i 16, py 14, b); box {168, 4, 9, 1}, no swizzle
static int image_map_u8(const uint8_t* img, int batch, CUtensorMap* map) {
  const cuuint64_t dims[4] = {VITK_IMG * 3 / 4, 16, 14, (cuuint64_t)batch};
  const cuuint64_t strides[3] = {VITK_IMG * 3, 16 * VITK_IMG * 3, (cuuint64_t)VITK_IMG * VITK_IMG * 3};
  const cuuint32_t box[4] = {VITK_IMG * 3 / 4, 4, 9, 1};
  return encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, img, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, "uint8 image");
}

static int fwd_tc(const void* images, bool u8, const float* mean3, const float* std3, const void* wpe16, const float* bpe, const float* cls,
                  const float* pos, float* x0, int batch, cudaStream_t st) {
  CUtensorMap map_img, map_w;
  if (u8) VITK_TRY(image_map_u8((const uint8_t*)images, batch, &map_img));
  else VITK_TRY(image_map_f32((const float*)images, batch, 4, 9, &map_img));
  {
    const cuuint64_t dims[2] = {VITK_DIM, VITK_DIM};
    const cuuint64_t strides[1] = {VITK_DIM * 2};
    const cuuint32_t box[2] = {64, BN};
    VITK_TRY(encode(&map_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wpe16, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "patch weights (bf16)"));
  }
  FwdParams p{};
  p.bpe = bpe; p.cls = cls; p.pos = pos; p.x0 = x0; p.batch = batch;
  for (int c = 0; c < 3; ++c) { p.mean[c] = u8 ? mean3[c] : 0.f; p.stdv[c] = u8 ? std3[c] : 1.f; }
  const int grid = batch * 2 * (VITK_DIM / BN);
  if (u8) {
    VITK_TRY(set_max_dyn_smem_once((const void*)patch_embed_fwd_kernel<true>, (int)FWD_SMEM));
    VITK_LAUNCH((patch_embed_fwd_kernel<true>), grid, FWD_THREADS, FWD_SMEM, st, map_img, map_w, p);
  } else {
    VITK_TRY(set_max_dyn_smem_once((const void*)patch_embed_fwd_kernel<false>, (int)FWD_SMEM));
    VITK_LAUNCH((patch_embed_fwd_kernel<false>), grid, FWD_THREADS, FWD_SMEM, st, map_img, map_w, p);
  }
  return VITK_OK;
}

static int fwd_simt(const float* images, const float* wpe, const float* bpe, const float* cls, const float* pos, float* x0, int batch,
                    cudaStream_t st) {
  GemmProblem p{};
  p.I = batch * VITK_NTOK; p.J = VITK_DIM; p.R = VITK_DIM;
  p.A = images; p.B = wpe; p.in_dtype = VITK_F32;
  p.la = layout_patches_rows_tok(); p.lb = layout_rowmajor(VITK_DIM);
  p.ep.mode = E_PATCH; p.ep.out = x0; p.ep.out_dtype = VITK_F32; p.ep.bias = bpe; p.ep.residual = pos;
  p.ep.aux = const_cast<float*>(cls); p.ep.ldc = VITK_DIM;
  return gemm_simt(p, 1, st);
}

static bool use_tc(int precision, int engine) { return precision == VITK_PREC_BF16 && engine != VITK_ENGINE_SIMT; }

static int u8_to_nchw(const uint8_t* images_hwc, const float* mean3, const float* std3, float* out, int batch, cudaStream_t st) {
  const int64_t total = (int64_t)batch * VITK_IMG * VITK_IMG;
  const int grid = (int)((total + 255) / 256 < (int64_t)sm_count() * 16 ? (total + 255) / 256 : (int64_t)sm_count() * 16);
  VITK_LAUNCH((u8_to_nchw_kernel), grid, 256, 0, st, images_hwc, mean3[0], mean3[1], mean3[2], std3[0], std3[1], std3[2], out, batch);
  return VITK_OK;
}

}  // namespace pe
}  // namespace vitk

using namespace vitk;

extern "C" int vitk_patch_embed_fwd(const float* images, const float* wpe, const void* wpe16, const float* bpe, const float* cls,
                                    const float* pos, float* x0, int batch, int precision, int engine, void* stream) {
  VITK_CHECK_ARG(images && wpe && bpe && cls && pos && x0 && batch > 0);
  VITK_CHECK_ARG(precision == VITK_PREC_FP32_VALIDATE || precision == VITK_PREC_BF16);
  if (pe::use_tc(precision, engine)) {
    VITK_CHECK_ARG(wpe16);
    return pe::fwd_tc(images, false, nullptr, nullptr, wpe16, bpe, cls, pos, x0, batch, (cudaStream_t)stream);
  }
  return pe::fwd_simt(images, wpe, bpe, cls, pos, x0, batch, (cudaStream_t)stream);
}

extern "C" int vitk_patch_embed_fwd_u8(const uint8_t* images_hwc, const float* mean3, const float* std3, const float* wpe,
                                       const void* wpe16, const float* bpe, const float* cls, const float* pos, float* x0,
                                       float* nchw_scratch, int batch, int precision, int engine, void* stream) {
  VITK_CHECK_ARG(images_hwc && mean3 && std3 && wpe && bpe && cls && pos && x0 && batch > 0);
  VITK_CHECK_ARG(precision == VITK_PREC_FP32_VALIDATE || precision == VITK_PREC_BF16);
  cudaStream_t st = (cudaStream_t)stream;
  if (pe::use_tc(precision, engine)) {
    VITK_CHECK_ARG(wpe16);
    return pe::fwd_tc(images_hwc, true, mean3, std3, wpe16, bpe, cls, pos, x0, batch, st);
  }
  VITK_CHECK_ARG(nchw_scratch);      // fp32-validate / SIMT: ToTensor + Normalize into the scratch, then the in-place patch GEMM
  VITK_TRY(pe::u8_to_nchw(images_hwc, mean3, std3, nchw_scratch, batch, st));
  return pe::fwd_simt(nchw_scratch, wpe, bpe, cls, pos, x0, batch, st);
}

extern "C" int vitk_u8_to_nchw(const uint8_t* images_hwc, const float* mean3, const float* std3, float* out, int batch, void* stream) {
  VITK_CHECK_ARG(images_hwc && mean3 && std3 && out && batch > 0);
  return pe::u8_to_nchw(images_hwc, mean3, std3, out, batch, (cudaStream_t)stream);
}

extern "C" int vitk_patch_embed_wgrad(const float* dx0, const void* dx0_bf16, const float* images, float* dwpe, float* dbpe,
                                      float* dcls, float* dpos, int batch, int precision, int engine, void* stream) {
  VITK_CHECK_ARG(dx0 && images && dwpe && dbpe && dcls && dpos && batch > 0);
  cudaStream_t st = (cudaStream_t)stream;
  if (pe::use_tc(precision, engine)) {
    VITK_CHECK_ARG(dx0_bf16);
    CUtensorMap map_img, map_dx;
    VITK_TRY(pe::image_map_f32(images, batch, 16, 4, &map_img));
    const cuuint64_t dims[3] = {VITK_DIM, VITK_NTOK, (cuuint64_t)batch};
    const cuuint64_t strides[2] = {VITK_DIM * 2, (cuuint64_t)VITK_NTOK * VITK_DIM * 2};
    const cuuint32_t box[3] = {64, 64, 1};
    VITK_TRY(pe::encode(&map_dx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dx0_bf16, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "dx0 (bf16)"));
    int n_splits = sm_count() / pe::WG_TILES;
    if (n_splits < 1) n_splits = 1;
    if (n_splits > batch) n_splits = batch;
    VITK_TRY(set_max_dyn_smem_once((const void*)pe::patch_embed_wgrad_kernel, (int)pe::WG_SMEM));
    VITK_LAUNCH((pe::patch_embed_wgrad_kernel), pe::WG_TILES * n_splits, pe::WG_THREADS, pe::WG_SMEM, st, map_img, map_dx, dwpe, batch, n_splits);
  } else {
    GemmProblem p{};
    p.I = VITK_DIM; p.J = VITK_DIM; p.R = batch * VITK_NTOK;
    p.A = dx0; p.B = images; p.in_dtype = VITK_F32;
    p.la = layout_transposed(VITK_DIM); p.lb = layout_patches_rows_k();
    p.ep.mode = E_ACCUM; p.ep.out = dwpe; p.ep.ldc = VITK_DIM; p.ep.out_dtype = VITK_F32;
    int splits = (4 * sm_count() + 143) / 144;
    if (splits > (p.R + 255) / 256) splits = (p.R + 255) / 256;
    VITK_TRY(gemm_simt(p, splits, st));
  }
  VITK_LAUNCH((pe::embed_param_grads_kernel), dim3(VITK_NTOK, (batch + pe::EG_BGROUP - 1) / pe::EG_BGROUP), 192, 0, st, dx0, batch, dpos, dcls, dbpe);
  return VITK_OK;
}
