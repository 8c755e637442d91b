// tcgen05 / TMEM / TMA GEMM engine for sm_100a (bf16 operands, fp32 accumulation in tensor memory).
//
//   C[i][j] = sum_r A(i,r) * B(j,r)  + fused epilogue (gemm.cuh)
//
// One persistent CTA per SM, 320 threads:
//   warp 0        TMA producer   (cp.async.bulk.tensor.3d -> 128B-swizzled smem ring, mbarrier expect_tx)
//   warp 1        MMA issuer     (one elected lane: tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16,
//                                 accumulators double-buffered in TMEM; tcgen05.commit -> mbarriers)
//   warps 2..9    epilogue       (tcgen05.ld 32x32b -> registers -> swizzled smem transpose -> coalesced
//                                 bias/GELU/residual/scatter global IO; two warps per TMEM lane quarter)
// Both operands may be K-major (reduction index contiguous: forward X W^T) or MN-major (row index
// contiguous: dgrad's W, wgrad's dY^T and X) -- the major-ness is a bit in the instruction descriptor
// plus the canonical 128B-swizzle shared-memory layout the TMA boxes are written in -- and may be stored
// head-major ([C/64][M][64], q/k/v and their gradients), which is just a different 3-D tensor map.
// CTA pairs (template CG = 2, the default for every shape with >= 2 x 128 rows): the two CTAs of a 2-cluster own
// one 256 x BN tile -- each stages its own 128 rows of A and HALF of the B tile, the leader's elected lane issues
// tcgen05.mma.cta_group::2 (M = 256) against both shared memories and multicasts its commits to both CTAs' barriers.
// Per CTA and k-block that is 128 + BN/2 operand rows from L2 instead of 128 + BN: the kernel is bound by L2->SM
// bandwidth (~6.3 KB/clk chip-wide), not by the tensor pipe, so this is what moves it.
// Work: output tiles strided over the persistent CTAs (CTA pairs); accumulate epilogues (wgrad) instead use stream-K --
// the (tile, k-block) space is cut into one equal contiguous range per CTA and partial tiles are summed
// with fp32 vector atomics, so 18..72-tile weight-gradient GEMMs still load all 148 SMs evenly.
//
// Replaces the cuBLASLt GEMMs behind timm's nn.Linear layers (SURVEY.md 2.1 K4,K6,K7,K8 and their
// autograd backward), reached from /root/reference/train_advanced.py:327-330.
#include <cuda.h>

#include "gemm.cuh"

namespace vitk {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;              // 64 bf16 = 128 B = one swizzle row
constexpr int TC_EPI_WARPS = 8;            // two warps per TMEM lane quarter, interleaved over 32-column chunks
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_EPI_WARP0 = 2;
constexpr uint32_t TC_A_STAGE_BYTES = TC_BM * TC_BK * 2;  // 16 KB

// operand addressing modes (see header comment)
enum { OP_KM_FLAT = 0, OP_KM_SPLIT = 1, OP_MN_FLAT = 2, OP_MN_SPLIT = 3 };

struct TcParams {
  int32_t I, J, R;
  int32_t a_mode, b_mode;
  int32_t n_tiles_m, n_tiles_n, kb_total;
  int32_t streamk;          // 0: tile-strided, full K per tile; 1: contiguous (tile, k-block) unit range per CTA
  int64_t units_per_cta;    // stream-K: ceil(tiles * kb_total / gridDim.x)
  uint32_t idesc;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;  // bytes
  uint32_t a_kstep, b_kstep;            // bytes advanced per UMMA_K=16 step
  EpiParams ep;
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
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
// 2-CTA variant: the transaction bytes land on the mbarrier of the pair's leader CTA (shared::cluster address)
__device__ __forceinline__ void tma_load_3d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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

// 64-bit shared-memory matrix descriptor (SWIZZLE_128B, version 1 = Blackwell)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// ------------------------------------------------------------------------------------------------
// epilogue on 4 consecutive columns (j..j+3) of row i.  The accumulator chunk is transposed through a
// small swizzled shared-memory buffer first, so that 8 consecutive lanes cover 32 consecutive columns of
// ONE row: every global load/store below is a fully used 64/128-byte segment (coalesced), instead of the
// row-per-thread pattern tcgen05.ld hands out.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint2 pack_bf16x4(float4 v) { return make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w)); }
__device__ __forceinline__ float4 unpack_bf16x4(uint2 u) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// 8 rows (i, i+4, ..., i+28) x 4 columns (j..j+3) per lane.  All global loads of the 8 rows are issued
// before any dependent math/store so 8 requests per lane are in flight (no asm/memory-clobber fences here).
__device__ __forceinline__ void epilogue_rows8(const EpiParams& ep, int i, int j, int I, float4 (&v)[8]) {
  switch (ep.mode) {
    case E_STORE: {
      if (ep.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + j));
#pragma unroll
        for (int it = 0; it < 8; ++it) v[it] = f4_add(v[it], b);
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = i + it * 4;
        if (row >= I) continue;
        const int64_t o = (int64_t)row * ep.ldc + j;
        if (ep.out_dtype == VITK_BF16) *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out) + o) = pack_bf16x4(v[it]);
        else *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + o) = v[it];
      }
    } break;
    case E_BIAS_GELU: {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + j));
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = i + it * 4;
        if (row >= I) continue;
        const int64_t o = (int64_t)row * ep.ldc + j;
        const float4 ur = unpack_bf16x4(pack_bf16x4(f4_add(v[it], b)));  // GELU acts on the 16-bit fc1 output (autocast)
        float4 g, dg;
        gelu_fast_both(ur.x, g.x, dg.x); gelu_fast_both(ur.y, g.y, dg.y);
        gelu_fast_both(ur.z, g.z, dg.z); gelu_fast_both(ur.w, g.w, dg.w);
        if (ep.aux) *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.aux) + o) = pack_bf16x4(dg);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out) + o) = pack_bf16x4(g);
      }
    } break;
    case E_BIAS_RESIDUAL: {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + j));
      float4 r[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = i + it * 4;
        r[it] = (row < I) ? *reinterpret_cast<const float4*>(ep.residual + (int64_t)row * ep.ldc + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = i + it * 4;
        if (row < I) *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (int64_t)row * ep.ldc + j) = f4_add(r[it], f4_add(v[it], b));
      }
    } break;
    case E_QKV_SCATTER: {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + j));
      const int64_t base = (int64_t)(j >> 6) * ep.hm_rows * 64 + (j & 63);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = i + it * 4;
        if (row < I) *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out) + base + (int64_t)row * 64) = pack_bf16x4(f4_add(v[it], b));
      }
    } break;
    case E_GELU_BWD: {
      uint2 u[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = i + it * 4;
        u[it] = (row < I) ? *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(ep.aux) + (int64_t)row * ep.ldc + j) : make_uint2(0u, 0u);
      }
      float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = i + it * 4;
        if (row >= I) continue;
        const float4 uu = unpack_bf16x4(u[it]);
        float4 o = v[it];
        o.x *= uu.x; o.y *= uu.y; o.z *= uu.z; o.w *= uu.w;   // aux = gelu'(u), saved by the forward epilogue
        const uint2 ob = pack_bf16x4(o);
        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.out) + (int64_t)row * ep.ldc + j) = ob;
        cs = f4_add(cs, unpack_bf16x4(ob));   // bias gradient sums the values the wgrad GEMM will read
      }
      if (ep.colsum) {
        // column sums of this 32x32 chunk: lanes l, l^8, l^16, l^24 hold the same 4 columns (different rows)
        cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 8);  cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 8);
        cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 8);  cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 8);
        cs.x += __shfl_xor_sync(0xffffffffu, cs.x, 16); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, 16);
        cs.z += __shfl_xor_sync(0xffffffffu, cs.z, 16); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, 16);
        if ((threadIdx.x & 31) < 8)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(ep.colsum + j), "f"(cs.x), "f"(cs.y), "f"(cs.z), "f"(cs.w));
      }
    } break;
    case E_ACCUM: {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = i + it * 4;
        if (row >= I) continue;
        float* d = reinterpret_cast<float*>(ep.out) + (int64_t)row * ep.ldc + j;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d), "f"(v[it].x), "f"(v[it].y), "f"(v[it].z), "f"(v[it].w));
      }
    } break;
    case E_PATCH: {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + j));
      const float4 cls = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(ep.aux) + j));
      float4 pe[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = i + it * 4;
        pe[it] = (row < I) ? __ldg(reinterpret_cast<const float4*>(ep.residual + (int64_t)(row % VITK_NTOK) * ep.ldc + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = i + it * 4;
        if (row >= I) continue;
        const float4 val = (row % VITK_NTOK == 0) ? cls : f4_add(v[it], b);
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + (int64_t)row * ep.ldc + j) = f4_add(val, pe[it]);
      }
    } break;
    default: break;
  }
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
template <int BN, int CG> struct TcCfg {
  static constexpr int B_ROWS = BN / CG;                                   // rows of the B tile this CTA stages
  static constexpr uint32_t B_STAGE_BYTES = B_ROWS * TC_BK * 2;
  static constexpr uint32_t STAGE_BYTES = TC_A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES_FIT = (227 * 1024 - 1024 - 256 - TC_EPI_WARPS * 4096) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;  // double-buffered accumulator, power of two
  static constexpr uint32_t EPI_OFF = STAGES * STAGE_BYTES + 256;  // TC_EPI_WARPS x 4 KB transpose buffers after the barriers
  static constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (size_t)EPI_OFF + TC_EPI_WARPS * 4096;
};

template <int CG>
__device__ __forceinline__ void tc_tma(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  if constexpr (CG == 2) tma_load_3d_2cta(dst, map, bar, c0, c1, c2);
  else tma_load_3d(dst, map, bar, c0, c1, c2);
}
template <int CG>
__device__ __forceinline__ void tc_issue_operand_loads(const CUtensorMap* map, int mode, uint32_t dst, uint32_t bar,
                                                       int row0, int rows_in_tile, int r0) {
  // K-major: one box [rows_in_tile][64 r]; MN-major: rows_in_tile/64 boxes [64 r][64 rows], 8 KB apart
  if (mode == OP_KM_FLAT) {
    tc_tma<CG>(dst, map, bar, r0, row0, 0);
  } else if (mode == OP_KM_SPLIT) {
    tc_tma<CG>(dst, map, bar, 0, row0, r0 >> 6);
  } else if (mode == OP_MN_FLAT) {
    for (int a = 0; a < rows_in_tile / 64; ++a) tc_tma<CG>(dst + a * 8192, map, bar, row0 + a * 64, r0, 0);
  } else {
    for (int a = 0; a < rows_in_tile / 64; ++a) tc_tma<CG>(dst + a * 8192, map, bar, 0, r0, (row0 >> 6) + a);
  }
}

// Work iterator shared by the three warp roles (they must walk identical sequences).
struct TcWork {
  int tile, kb0, kb1;
};
struct TcWorkIter {
  int64_t cur, end;  // stream-K: unit cursor / end ; tile-strided: next tile / number of tiles
  int stride;
  // `cg` CTAs (one cluster) walk the same sequence
  __device__ __forceinline__ void init(const TcParams& p, int cg) {
    const int64_t tiles = (int64_t)p.n_tiles_m * p.n_tiles_n;
    const int cluster = blockIdx.x / cg;
    stride = gridDim.x / cg;
    if (p.streamk) {
      const int64_t total = tiles * p.kb_total;
      cur = (int64_t)cluster * p.units_per_cta;
      end = cur + p.units_per_cta < total ? cur + p.units_per_cta : total;
    } else {
      cur = cluster;
      end = tiles;
    }
  }
  __device__ __forceinline__ bool next(const TcParams& p, TcWork& w) {
    if (cur >= end) return false;
    if (p.streamk) {
      w.tile = (int)(cur / p.kb_total);
      w.kb0 = (int)(cur % p.kb_total);
      const int64_t left = end - cur;
      w.kb1 = (int)((int64_t)(p.kb_total - w.kb0) < left ? p.kb_total : w.kb0 + left);
      cur += w.kb1 - w.kb0;
    } else {
      w.tile = (int)cur;
      w.kb0 = 0;
      w.kb1 = p.kb_total;
      cur += stride;
    }
    return true;
  }
};

template <int BN, int CG>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const TcParams p) {
  using Cfg = TcCfg<BN, CG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  // barrier layout: full[S] | empty[S] | tmem_full[2] | tmem_empty[2] | tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * (2 * Cfg::STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTA pair: rank 0 (leader) owns the barriers the pair shares -- `full` (TMA bytes of both CTAs) and `tmem_empty`
  // (epilogue warps of both CTAs); `empty` and `tmem_full` exist in both CTAs and are signalled by multicast commits.
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), CG * TC_EPI_WARPS);  // one arrive per epilogue warp of every CTA of the pair
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)),
                   "r"(Cfg::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_ptr_smem)),
                   "r"(Cfg::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // peer barriers initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer (every CTA loads its own A rows and its share of the B rows) ==========
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      TcWorkIter wi;
      wi.init(p, CG);
      TcWork w;
      while (wi.next(p, w)) {
        const int tile = w.tile;
        const int i0 = (tile / p.n_tiles_n) * (TC_BM * CG) + (int)rank * TC_BM;
        const int j0 = (tile % p.n_tiles_n) * BN + (int)rank * Cfg::B_ROWS;
        const int kb0 = w.kb0, kb1 = w.kb1;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (rank == 0) mbar_expect_tx(full_bar(stage), CG * Cfg::STAGE_BYTES);
          const uint32_t fb = CG == 2 ? mapa_u32(full_bar(stage), 0) : full_bar(stage);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          tc_issue_operand_loads<CG>(&map_a, p.a_mode, sa, fb, i0, TC_BM, kb * TC_BK);
          tc_issue_operand_loads<CG>(&map_b, p.b_mode, sa + TC_A_STAGE_BYTES, fb, j0, Cfg::B_ROWS, kb * TC_BK);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if constexpr (CG == 2) {
        // tail: the leader's last multicast commits must have landed in this CTA's `empty` barriers before it may exit
        for (int s = 0; s < Cfg::STAGES; ++s) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      TcWorkIter wi;
      wi.init(p, CG);
      TcWork w;
      while (wi.next(p, w)) {
        const int kb0 = w.kb0, kb1 = w.kb1;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + TC_A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * p.a_kstep, p.a_lbo, p.a_sbo);
            const uint64_t bdesc = make_smem_desc(sb + k * p.b_kstep, p.b_lbo, p.b_sbo);
            if constexpr (CG == 2) tc_mma_bf16_2cta(d_tmem, adesc, bdesc, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else tc_mma_bf16(d_tmem, adesc, bdesc, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // smem slot free (in both CTAs) once these MMAs retire
          if constexpr (CG == 2) tc_commit_2cta(empty_bar(stage)); else tc_commit(empty_bar(stage));
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete (both CTAs' epilogues)
        if constexpr (CG == 2) tc_commit_2cta(tfull_bar(acc)); else tc_commit(tfull_bar(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    const int half = (warp - TC_EPI_WARP0) >> 2;  // which interleaved set of column chunks
    int acc = 0;
    uint32_t acc_phase = 0;
    TcWorkIter wi;
    wi.init(p, CG);
    TcWork w;
    while (wi.next(p, w)) {
      const int tile = w.tile;
      const int i0 = (tile / p.n_tiles_n) * (TC_BM * CG) + (int)rank * TC_BM, j0 = (tile % p.n_tiles_n) * BN;
      // While the MMAs of this tile still run: pull the epilogue's second operand (saved gelu' / fp32 residual)
      // for this warp's 32 rows into L2, so the dependent loads below are L2 hits instead of HBM round trips.
      if (p.ep.mode == E_GELU_BWD || p.ep.mode == E_BIAS_RESIDUAL) {
        const int elt = p.ep.mode == E_GELU_BWD ? 2 : 4;
        const char* src = p.ep.mode == E_GELU_BWD ? reinterpret_cast<const char*>(p.ep.aux)
                                                  : reinterpret_cast<const char*>(p.ep.residual);
        const int lines = BN * elt / 128;
        for (int idx = half * 32 + lane; idx < 32 * lines; idx += 64) {
          const int row = i0 + q * 32 + idx / lines;
          if (row < p.I)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(src + ((int64_t)row * p.ep.ldc + j0) * elt + (idx % lines) * 128));
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      float4* tb = reinterpret_cast<float4*>(smem + Cfg::EPI_OFF + (uint32_t)(warp - TC_EPI_WARP0) * 4096u);
      const int sub_row = lane >> 3, c4 = lane & 7;
      if (i0 + q * 32 < p.I) {   // warps whose 32 rows are all past the end of the matrix have nothing to write
#pragma unroll 1
        for (int c = half; c < BN / 32; c += TC_EPI_WARPS / 4) {
          uint32_t raw[32];
          tc_ld32(taddr + c * 32, raw);
          // row `lane` of the 32x32 chunk -> swizzled smem (16-byte column group cg at cg ^ (row & 7))
#pragma unroll
          for (int cg = 0; cg < 8; ++cg)
            tb[lane * 8 + (cg ^ (lane & 7))] = make_float4(__uint_as_float(raw[cg * 4]), __uint_as_float(raw[cg * 4 + 1]),
                                                           __uint_as_float(raw[cg * 4 + 2]), __uint_as_float(raw[cg * 4 + 3]));
          __syncwarp();
          float4 v[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + sub_row;
            v[it] = tb[r * 8 + (c4 ^ (r & 7))];
          }
          __syncwarp();
          epilogue_rows8(p.ep, i0 + q * 32 + sub_row, j0 + c * 32 + c4 * 4, p.I, v);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(tempty_bar(acc), 0));
        else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // nobody exits while its peer may still touch it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
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

// Builds the 3-D map of one operand and returns its addressing mode.
static int make_operand_map(const void* base, const MatLayout& l, int rows, int R, int tile_rows, CUtensorMap* map, int* mode) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return VITK_ERR_DRIVER; }
  cuuint64_t dims[3], strides[2];
  cuuint32_t box[3], estr[3] = {1, 1, 1};
  if (l.split == 0 && l.s_col == 1) {             // K-major flat: (r, row)
    *mode = OP_KM_FLAT;
    dims[0] = R; dims[1] = rows; dims[2] = 1;
    strides[0] = (cuuint64_t)l.s_row * 2; strides[1] = strides[0] * rows;
    box[0] = TC_BK; box[1] = tile_rows; box[2] = 1;
  } else if (l.split == 2 && l.s_col == 1 && l.s_row == 64) {  // K-major, reduction index stored in 64-blocks
    *mode = OP_KM_SPLIT;
    dims[0] = 64; dims[1] = rows; dims[2] = R / 64;
    strides[0] = 128; strides[1] = (cuuint64_t)l.s_blk * 2;
    box[0] = 64; box[1] = tile_rows; box[2] = 1;
  } else if (l.split == 0 && l.s_row == 1) {      // MN-major flat: (row, r)
    *mode = OP_MN_FLAT;
    dims[0] = rows; dims[1] = R; dims[2] = 1;
    strides[0] = (cuuint64_t)l.s_col * 2; strides[1] = strides[0] * R;
    box[0] = 64; box[1] = TC_BK; box[2] = 1;
  } else if (l.split == 1 && l.s_row == 1 && l.s_col == 64) {  // MN-major, row index stored in 64-blocks
    *mode = OP_MN_SPLIT;
    dims[0] = 64; dims[1] = R; dims[2] = rows / 64;
    strides[0] = 128; strides[1] = (cuuint64_t)l.s_blk * 2;
    box[0] = 64; box[1] = TC_BK; box[2] = 1;
  } else {
    set_error("gemm_tc: operand layout not expressible as a TMA tensor map");
    return VITK_ERR_UNSUPPORTED;
  }
  if (((uintptr_t)base & 15) || (strides[0] & 15) || (strides[1] & 15)) {
    set_error("gemm_tc: operand base/strides must be 16-byte aligned");
    return VITK_ERR_ARG;
  }
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (dims %llu,%llu,%llu strides %llu,%llu box %u,%u,%u)", (int)r,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              (unsigned long long)strides[0], (unsigned long long)strides[1], box[0], box[1], box[2]);
    return VITK_ERR_DRIVER;
  }
  return VITK_OK;
}

static int g_tc_debug[8] = {0, 0, 0, 0, 0, 0, 0, 0};
int attn_debug_variant() { return g_tc_debug[3]; }

template <int BN, int CG>
static int launch_tc(const GemmProblem& pr, cudaStream_t st) {
  using Cfg = TcCfg<BN, CG>;
  static bool configured = false;
  if (!configured) {
    VITK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES));
    configured = true;
  }
  CUtensorMap map_a, map_b;
  TcParams p{};
  p.I = pr.I; p.J = pr.J; p.R = pr.R;
  VITK_TRY(make_operand_map(pr.A, pr.la, pr.I, pr.R, TC_BM, &map_a, &p.a_mode));
  VITK_TRY(make_operand_map(pr.B, pr.lb, pr.J, pr.R, Cfg::B_ROWS, &map_b, &p.b_mode));
  const bool a_mn = p.a_mode >= OP_MN_FLAT, b_mn = p.b_mode >= OP_MN_FLAT;
  // instruction descriptor: D=f32, A=B=bf16, majors, N>>3, M>>4 (M = 128 per CTA: 256 for a CTA pair)
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
            ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((TC_BM * CG) >> 4) << 24);
  // canonical SWIZZLE_128B layouts: K-major: 8-row groups 1024 B apart (SBO), LBO unused;
  // MN-major: 8-r groups 1024 B apart (SBO), 64-wide MN atoms TC_BK*128 B apart (LBO)
  p.a_sbo = 1024; p.a_lbo = a_mn ? TC_BK * 128 : 16; p.a_kstep = a_mn ? 16 * 128 : 32;
  p.b_sbo = 1024; p.b_lbo = b_mn ? TC_BK * 128 : 16; p.b_kstep = b_mn ? 16 * 128 : 32;
  if (g_tc_debug[0] == 1) {  // debug variant: swap LBO/SBO roles of MN-major operands
    if (a_mn) { p.a_sbo = TC_BK * 128; p.a_lbo = 1024; }
    if (b_mn) { p.b_sbo = TC_BK * 128; p.b_lbo = 1024; }
  }
  p.n_tiles_m = (pr.I + TC_BM * CG - 1) / (TC_BM * CG);
  p.n_tiles_n = pr.J / BN;
  p.kb_total = (pr.R + TC_BK - 1) / TC_BK;
  const int tiles = p.n_tiles_m * p.n_tiles_n;
  const int slots = sm_count() / CG;            // persistent CTAs (CTA pairs)
  int grid = tiles < slots ? tiles : slots;
  p.streamk = 0;
  p.units_per_cta = 0;
  if (pr.ep.mode == E_ACCUM && g_tc_debug[1] != 1) {
    const int64_t total = (int64_t)tiles * p.kb_total;
    grid = total < slots ? (int)total : slots;
    p.streamk = 1;
    p.units_per_cta = (total + grid - 1) / grid;
    grid = (int)((total + p.units_per_cta - 1) / p.units_per_cta);
  }
  p.ep = pr.ep;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid * CG, 1, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CG == 2 ? 1 : 0;
  VITK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CG>, map_a, map_b, p));
  VITK_LAUNCH_CHECK();
  return VITK_OK;
}

int gemm_tc(const GemmProblem& pr, cudaStream_t st) {
  VITK_CHECK_ARG(pr.I > 0 && pr.J > 0 && pr.R > 0 && pr.A && pr.B && pr.ep.out);
  if (pr.in_dtype != VITK_BF16) { set_error("gemm_tc: bf16 operands only"); return VITK_ERR_UNSUPPORTED; }
  if (pr.ep.mode != E_STORE && pr.ep.mode != E_ACCUM && pr.ep.mode != E_BIAS_RESIDUAL && pr.ep.mode != E_PATCH &&
      pr.ep.out_dtype != VITK_BF16) { set_error("gemm_tc: epilogue expects bf16 output"); return VITK_ERR_UNSUPPORTED; }
  if (pr.J % 128 != 0 || pr.ep.ldc % 8 != 0) { set_error("gemm_tc: J must be a multiple of 128 (got %d)", pr.J); return VITK_ERR_UNSUPPORTED; }
  // Tile shape: CTA pair (CG = 2, 256 x BN) or single CTA (128 x BN), BN in {256, 192, 128}.  The kernel is bound by
  // L2->SM operand traffic long before the tensor pipe: per CTA and 64-deep k-block it pulls (128 + BN/CG) rows of
  // 128 B at ~42.5 B/clk/SM (6.3 KB/clk chip-wide, B300_MICROARCH.md) while the MMAs need 2*BN clk.  Pick the
  // candidate that minimises waves x k-blocks x max(L2, MMA) clocks; with M = B*197 the tile count is rarely a
  // multiple of the slot count, so the wave term matters (12608 x 768 on 74 pairs: 150 tiles of 256x256 = 3 waves for
  // 2.03 waves of work, 200 tiles of 256x192 = 3 waves of 3/4 the cost).
  const bool b_mn = pr.lb.s_row == 1 && pr.lb.s_col != 1;   // MN-major B: staged in 64-row atoms
  int cg = 0, bn = 0;
  const int forced_bn = (g_tc_debug[2] == 128 || g_tc_debug[2] == 192 || g_tc_debug[2] == 256) ? g_tc_debug[2] : 0;
  if (pr.ep.mode == E_ACCUM) {
    bn = forced_bn ? forced_bn : (pr.J % 256 == 0 ? 256 : 128);  // stream-K balances by itself
    cg = (pr.I > TC_BM && g_tc_debug[4] != 1 && !(b_mn && (bn / 2) % 64 != 0)) ? 2 : 1;
  } else {
    const long kb = (pr.R + TC_BK - 1) / TC_BK;
    double best = -1.0;
    for (int c : {2, 1}) {
      if (c == 2 && (pr.I <= TC_BM || g_tc_debug[4] == 1)) continue;
      if (c == 1 && g_tc_debug[4] == 2 && best >= 0.0) continue;   // debug: pairs forced (when expressible)
      const long slots = sm_count() / c;
      const long tiles_m = (pr.I + TC_BM * c - 1) / (TC_BM * c);
      for (int cand : {256, 192, 128}) {
        if (pr.J % cand != 0) continue;
        if (b_mn && (cand / c) % 64 != 0) continue;
        if (forced_bn && forced_bn != cand) continue;
        const long tiles = tiles_m * (pr.J / cand);
        const long waves = (tiles + slots - 1) / slots;
        const double l2 = (128.0 + cand / c) * 128.0 / 42.5, mma = 2.0 * cand;
        const double cost = (double)waves * ((double)kb * (l2 > mma ? l2 : mma) + 600.0);
        if (best < 0.0 || cost < best) { best = cost; bn = cand; cg = c; }
      }
    }
  }
  if (bn == 0 || pr.J % bn != 0) { set_error("gemm_tc: no BLOCK_N divides J=%d", pr.J); return VITK_ERR_UNSUPPORTED; }
  if (cg == 2) {
    if (bn == 256) return launch_tc<256, 2>(pr, st);
    if (bn == 192) return launch_tc<192, 2>(pr, st);
    return launch_tc<128, 2>(pr, st);
  }
  if (bn == 256) return launch_tc<256, 1>(pr, st);
  if (bn == 192) return launch_tc<192, 1>(pr, st);
  return launch_tc<128, 1>(pr, st);
}

}  // namespace vitk

extern "C" int vitk_debug_set(int key, int value) {
  if (key < 0 || key >= 8) return VITK_ERR_ARG;
  vitk::g_tc_debug[key] = value;
  return VITK_OK;
}
