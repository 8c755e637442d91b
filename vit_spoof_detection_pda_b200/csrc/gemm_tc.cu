// tcgen05 / TMEM / TMA GEMM engine for sm_100a (bf16 operands, fp32 accumulation in tensor memory).
//
//   C[i][j] = sum_r A(i,r) * B(j,r)  + fused epilogue (gemm.cuh)
//
// One persistent CTA per SM, 320 threads; the whole warp of each role walks the work list (warp-uniform control flow keeps
// addresses in uniform registers), one elected lane issues:
//   warp 0        TMA producer   (cp.async.bulk.tensor.3d -> 128B-swizzled smem ring of 5..8 stages, mbarrier expect_tx;
//                                 coordinates advance by additions only)
//   warp 1        MMA issuer     (tcgen05.mma.kind::f16, 128|256 x BN x 16, accumulators double-buffered in TMEM, smem
//                                 descriptors = constant high word + running low word; tcgen05.commit -> mbarriers)
//   warps 2..9    epilogue       (tcgen05.ld 32x32b gives each lane an accumulator ROW -> epilogue math -> swizzled
//                                 32 x 32 staging tile -> TMA store; second operands TMA-loaded into the staging tile;
//                                 compile-time epilogue kinds EK_*; stream-K accumulate keeps a transposing red.global path)
// Both operands may be K-major (reduction index contiguous: forward X W^T) or MN-major (row index
// contiguous: dgrad's W, wgrad's dY^T and X) -- the major-ness is a bit in the instruction descriptor
// plus the canonical 128B-swizzle shared-memory layout the TMA boxes are written in -- and may be stored
// head-major ([C/64][M][64], q/k/v and their gradients), which is just a different 3-D tensor map.
// CTA pairs (template CG = 2, the default for every shape with >= 2 x 128 rows): the two CTAs of a 2-cluster own
// one 256 x BN tile -- each stages its own 128 rows of A and HALF of the B tile, the leader's elected lane issues
// tcgen05.mma.cta_group::2 (M = 256) against both shared memories and multicasts its commits to both CTAs' barriers.
// Per CTA and k-block that is 128 + BN/2 operand rows instead of 128 + BN: what bounds the kernel is the rate at which
// one SM can take operand bytes in (~45 B/clk through TMA, tools/micro/tma_ingest.cu), not the tensor
// pipe, so halving the B traffic is what moves it.
// Work: output tiles strided over the persistent CTAs (CTA pairs).  Accumulate epilogues (wgrad: 9..36 output tiles, a
// 12,608-deep reduction) split K and sum partial tiles with fp32 vector atomics: sliced split-K -- cluster c owns k-slice
// c / tiles of tile c % tiles, so all clusters of a slice sweep the same k range in lock-step and drain one accumulator
// each -- when tiles x slices fill >= 90 % of the clusters, contiguous stream-K ranges over the (tile, k-block) space
// otherwise (tc_decompose; the plan is queryable on the host: vitk_gemm_plan).  Non-accumulating GEMMs whose last wave would be
// nearly empty (M = B*197: 150 tiles on 74 pairs) deal the k-blocks of the last tiles out to all clusters of the same launch;
// partial accumulators meet in an fp32 scratch and the last cluster to arrive at a region's ticket applies the epilogue
// (decomposition mode 3, tc_tail_plan; vitk_gemm_tail_plan).
// Every launch carries the programmatic-dependent-launch attribute: barrier init, TMEM allocation and tensor-map
// prefetch run before pdl_sync(), i.e. under the tail of the previous kernel.
//
// Replaces the cuBLASLt GEMMs behind timm's nn.Linear layers (SURVEY.md 2.1 K4,K6,K7,K8 and their
// autograd backward), reached from /root/reference/train_advanced.py:327-330.
#include <cuda.h>
#include <string.h>

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

// Branch-free TMA coordinates of one operand: box `a` of k-block `kb` of the tile whose first row is `row0` is at
//   c0 = row0*fr0 + kb*d0 + a*e0,  c1 = row0*fr1 + kb*d1,  c2 = (row0>>6)*fr2 + kb*d2 + a*e2     (boxes 8 KB apart in smem)
struct OpCoord {
  int32_t fr0, fr1, fr2, d0, d1, d2, e0, e2, nbox;
};

struct TcParams {
#ifdef VITK_DEV
  int32_t dbg;              // development build, timing experiments only (vitk_debug_set(7, v)): bit 0 skip the epilogue
                            // body, bit 1 no operand loads -- results invalid; the release build has no such path
#endif
  int32_t I, J, R;
  int32_t a_mode, b_mode;
  OpCoord ca, cb;
  int32_t n_tiles_m, n_tiles_n, kb_total;
  int32_t streamk;          // 0: tile-strided, full K per tile; 1: contiguous (tile, k-block) unit range per CTA;
                            // 2: sliced split-K, CTA c owns k-slice c / tiles of tile c % tiles;
                            // 3: split tail -- the (tile, k-block) units of the last tiles [n_whole, tiles) in contiguous ranges
                            //    over ALL clusters, then tiles [0, n_whole) tile-strided with the fused epilogue (tc_tail_plan)
  int64_t units_per_cta;    // stream-K: ceil(tiles * kb_total / gridDim.x); sliced split-K: number of k-slices; split tail:
                            // ceil((tiles - n_whole) * kb_total / gridDim.x)
  int32_t n_whole;          // split tail: number of whole-K tiles
  int32_t tail_row0;        // split tail: first matrix row of the first tail tile (scratch row 0)
  float* tail_acc;          // split tail: fp32 [I - tail_row0][J] partial sums, all zero between launches
  int* tail_tickets;        // split tail: arrival counters, one per (tail tile, CTA rank, epilogue warp); zero between launches
  uint32_t idesc;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;  // bytes
  uint32_t a_kstep, b_kstep;            // bytes advanced per UMMA_K=16 step
  EpiParams ep;
};

#ifdef VITK_DEV
#define TC_DBG(p) ((p).dbg)
#else
#define TC_DBG(p) 0
#endif

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
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_commit_2cta(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
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
// tcgen05.mma with the 64-bit descriptors assembled from a constant high word and a running low word
__device__ __forceinline__ void tc_mma_desc(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accum, bool pair) {
  if (pair)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "setp.ne.b32 p, %6, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n"
        "}" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accum) : "memory");
  else
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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

__device__ __forceinline__ void tc_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
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
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// TMA store / load of epilogue tiles (2-D row-major, 3-D head-major) and bulk-group bookkeeping
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
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
// Epilogue kinds (compile-time).  All but EK_LEGACY move the output through shared memory and TMA:
// tcgen05.ld hands every lane one accumulator ROW (32 consecutive columns); the lane applies the epilogue math,
// writes its row into a swizzled 32 x 32 staging tile and one elected lane issues a cp.async.bulk.tensor store
// (rows past the end of the matrix are clipped by the tensor map).  Second operands (fp32 residual, saved gelu')
// arrive the same way: a TMA load into the staging tile issued BEFORE the accumulator is ready, combined in place.
// No per-thread global loads/stores and no global-memory latency remain on the epilogue warps' critical path.
enum { EK_LEGACY = 0, EK_STORE_BF16 = 1, EK_STORE_F32 = 2, EK_GELU = 3, EK_RESIDUAL = 4, EK_SCATTER = 5, EK_GELU_BWD = 6 };
// staging per epilogue warp: two 4 KB slots for fp32 tiles, two 2 KB slots for bf16 tiles (EK_GELU: one slot pair g | u).
// Every KB not spent here is operand-ring depth: the mainloop needs ~1.5 us of loads in flight to ride out DRAM latency.
// Under 256-wide CTA-pair tiles (32 KB stages) the second fp32 slot would cost the ring its fifth stage -- measured 20 % per
// k-block (fc2 forward at bs 64: 35 us per wave with 4 stages against 29 us of equal work with 256 x 192 tiles and 5) -- so
// those kernels stage through ONE slot per warp: the chunks of a tile then serialise on their TMA store / load round trips,
// which the 20+ us mainloop of the next tile hides (double-buffered accumulators).
__host__ __device__ constexpr uint32_t tc_epi_warp_bytes(int ek, int bn, int cg) {
  return (ek == EK_STORE_F32 || ek == EK_RESIDUAL) ? ((bn == 256 && cg == 2) ? 4096u : 8192u) : 4096u;
}

template <int BN, int CG, int EK> struct TcCfg {
  static constexpr int B_ROWS = BN / CG;                                   // rows of the B tile this CTA stages
  static constexpr uint32_t B_STAGE_BYTES = B_ROWS * TC_BK * 2;
  static constexpr uint32_t STAGE_BYTES = TC_A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr uint32_t EPI_WARP_BYTES = tc_epi_warp_bytes(EK, BN, CG);
  static constexpr uint32_t TAIL_BYTES = TC_EPI_WARPS * EPI_WARP_BYTES + TC_EPI_WARPS * 512 + 512;  // staging | bias | barriers
  static constexpr int STAGES_FIT = (227 * 1024 - 1024 - (int)TAIL_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  static constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;  // double-buffered accumulator, power of two
  static constexpr uint32_t EPI_OFF = STAGES * STAGE_BYTES;                                  // 1024-aligned staging tiles
  static constexpr uint32_t BIAS_OFF = EPI_OFF + TC_EPI_WARPS * EPI_WARP_BYTES;              // 512 B per epilogue warp
  static constexpr uint32_t BAR_OFF = BIAS_OFF + TC_EPI_WARPS * 512;
  static constexpr size_t SMEM_BYTES = 1024 /*align slack*/ + (size_t)STAGES * STAGE_BYTES + TAIL_BYTES;
};

template <int CG>
__device__ __forceinline__ void tc_tma(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  if constexpr (CG == 2) tma_load_3d_2cta(dst, map, bar, c0, c1, c2);
  else tma_load_3d(dst, map, bar, c0, c1, c2);
}
// Work iterator shared by the three warp roles (they must walk identical sequences).
struct TcWork {
  int tile, kb0, kb1;
  int tm, tn;   // tile row / tile column
};
struct TcWorkIter {
  int64_t cur, end;  // stream-K: unit cursor / end ; tile-strided: next tile / number of tiles
  int64_t tcur, tend;   // split tail: this cluster's unit range in the (tail tile, k-block) space
  int stride;
  // the CTAs of one cluster (CTA pair) walk the same sequence; `cluster` of `n_clusters`.  Host-callable: the plan query of
  // the C-ABI (vitk_gemm_plan_items) walks the very same code, and tests/test_host_logic.py checks that the items of all
  // clusters cover every (tile, k-block) unit exactly once.
  __host__ __device__ __forceinline__ void init(const TcParams& p, int cluster, int n_clusters) {
    const int64_t tiles = (int64_t)p.n_tiles_m * p.n_tiles_n;
    stride = n_clusters;
    tcur = tend = 0;
    if (p.streamk == 3) {          // one contiguous range of the tail's units, then whole tiles
      cur = cluster;
      end = p.n_whole;
      const int64_t total = (tiles - p.n_whole) * p.kb_total;
      tcur = (int64_t)cluster * p.units_per_cta;
      tend = tcur + p.units_per_cta < total ? tcur + p.units_per_cta : total;
    } else if (p.streamk == 2) {   // sliced split-K: exactly one (tile, k-slice) item per CTA (pair)
      cur = cluster;
      end = cur + 1;
    } else if (p.streamk) {
      const int64_t total = tiles * p.kb_total;
      cur = (int64_t)cluster * p.units_per_cta;
      end = cur + p.units_per_cta < total ? cur + p.units_per_cta : total;
    } else {
      cur = cluster;
      end = tiles;
    }
  }
  __host__ __device__ __forceinline__ bool next(const TcParams& p, TcWork& w) {
    if (!next_raw(p, w)) return false;
    w.tm = w.tile / p.n_tiles_n;     // consecutive CTAs walk along a tile row (column-first order measured identical)
    w.tn = w.tile % p.n_tiles_n;
    return true;
  }
  __host__ __device__ __forceinline__ bool next_raw(const TcParams& p, TcWork& w) {
    if (p.streamk == 3) {
      // this cluster's share of the tail tiles FIRST (w.tile >= p.n_whole): the fix-up behind a partial item -- fence, ticket,
      // the last arriver's read-back and epilogue -- then runs under the mainloop of the cluster's first whole tile
      if (tcur < tend) {
        w.tile = p.n_whole + (int)(tcur / p.kb_total);
        w.kb0 = (int)(tcur % p.kb_total);
        const int64_t left = tend - tcur;
        w.kb1 = (int)((int64_t)(p.kb_total - w.kb0) < left ? p.kb_total : w.kb0 + left);
        tcur += w.kb1 - w.kb0;
        return true;
      }
      if (cur >= end) return false;
      w.tile = (int)cur;             // whole-K tile, fused epilogue
      w.kb0 = 0;
      w.kb1 = p.kb_total;
      cur += stride;
      return true;
    }
    if (cur >= end) return false;
    if (p.streamk == 2) {
      const int tiles = p.n_tiles_m * p.n_tiles_n, S = (int)p.units_per_cta;
      const int slice = (int)cur / tiles;
      w.tile = (int)cur % tiles;
      w.kb0 = (int)((int64_t)slice * p.kb_total / S);
      w.kb1 = (int)((int64_t)(slice + 1) * p.kb_total / S);
      cur = end;
    } else if (p.streamk) {
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

// shared-memory offsets of lane `r`'s row inside a 32-row staging tile
__device__ __forceinline__ uint32_t row_off_64(int r, int k4) { return (uint32_t)(r * 64 + ((k4 ^ ((r >> 1) & 3)) << 4)); }   // bf16, SWIZZLE_64B
__device__ __forceinline__ uint32_t row_off_128(int r, int k8) { return (uint32_t)(r * 128 + ((k8 ^ (r & 7)) << 4)); }        // fp32, SWIZZLE_128B

// += column sums of one staged 32 x 32 bf16 tile (SWIZZLE_64B rows, row_off_64) into colsum[0..31] (32-byte aligned):
// lane reads the 16-byte chunk (lane & 3) of rows (lane >> 2) + 8 j -- four conflict-free 128-bit loads cover the tile --
// the 8 lanes sharing a chunk are summed with three shuffle stages, lanes 0..3 issue two vector reductions each.
__device__ __forceinline__ void tile_colsum_bf16(uint32_t t0, int lane, float* colsum) {
  float cs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) cs[i] = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 w = lds128(t0 + row_off_64((lane >> 2) + 8 * j, lane & 3));
    cs[0] += __uint_as_float(w.x << 16); cs[1] += __uint_as_float(w.x & 0xFFFF0000u);
    cs[2] += __uint_as_float(w.y << 16); cs[3] += __uint_as_float(w.y & 0xFFFF0000u);
    cs[4] += __uint_as_float(w.z << 16); cs[5] += __uint_as_float(w.z & 0xFFFF0000u);
    cs[6] += __uint_as_float(w.w << 16); cs[7] += __uint_as_float(w.w & 0xFFFF0000u);
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

template <int BN, int CG, int EK>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_d, const TcParams p) {
  using Cfg = TcCfg<BN, CG, EK>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_base + Cfg::BAR_OFF;
  // barrier layout: full[S] | empty[S] | tmem_full[2] | tmem_empty[2] | epi_load[8 warps][2] | tmem_ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (Cfg::STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * Cfg::STAGES + 2 + s); };
  auto eload_bar = [&](int we, int s) { return bar_base + 8u * (2 * Cfg::STAGES + 4 + we * 2 + s); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + Cfg::BAR_OFF + 8 * (2 * Cfg::STAGES + 4 + 2 * TC_EPI_WARPS));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dbg = TC_DBG(p);   // 0 (compile time) in the release build
  const unsigned long long tr_aux = ((unsigned long long)p.I << 40) | ((unsigned long long)p.J << 20) | (unsigned long long)p.R;
  trace_mark(TK_GEMM_TC, 0, tr_aux);
  // CTA pair: rank 0 (leader) owns the barriers the pair shares -- `full` (TMA bytes of both CTAs) and `tmem_empty`
  // (epilogue warps of both CTAs); `empty` and `tmem_full` exist in both CTAs and are signalled by multicast commits.
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    if constexpr (EK != EK_LEGACY) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_c)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_d)) : "memory");
    }
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), CG * TC_EPI_WARPS);  // one arrive per epilogue warp of every CTA of the pair
    }
    for (int s = 0; s < 2 * TC_EPI_WARPS; ++s) mbar_init(eload_bar(0, s), 1);
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
  pdl_sync();   // everything above overlapped the previous kernel's tail; global memory is touched only from here on
  trace_mark(TK_GEMM_TC, 1, tr_aux);

  if (warp == 0) {
    // ===================== TMA producer (every CTA loads its own A rows and its share of the B rows) ==========
    // The whole warp walks the loop (warp-uniform control flow keeps addresses in uniform registers); one elected
    // lane arms the barrier and issues the copies.  Coordinates advance by additions only (OpCoord).
    {
      int stage = 0;
      uint32_t phase = 0;
      TcWorkIter wi;
      wi.init(p, (int)blockIdx.x / CG, (int)gridDim.x / CG);
      TcWork w;
      const bool leader = elect_one();
      while (wi.next(p, w)) {
        const int i0 = w.tm * (TC_BM * CG) + (int)rank * TC_BM;
        const int j0 = w.tn * BN + (int)rank * Cfg::B_ROWS;
        const int kb0 = w.kb0, kb1 = w.kb1;
        int a0 = i0 * p.ca.fr0 + kb0 * p.ca.d0, a1 = i0 * p.ca.fr1 + kb0 * p.ca.d1, a2 = (i0 >> 6) * p.ca.fr2 + kb0 * p.ca.d2;
        int b0 = j0 * p.cb.fr0 + kb0 * p.cb.d0, b1 = j0 * p.cb.fr1 + kb0 * p.cb.d1, b2 = (j0 >> 6) * p.cb.fr2 + kb0 * p.cb.d2;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (dbg & 2) continue;
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (leader) {
            if (rank == 0) mbar_expect_tx(full_bar(stage), CG * Cfg::STAGE_BYTES);
            const uint32_t fb = CG == 2 ? mapa_u32(full_bar(stage), 0) : full_bar(stage);
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
#pragma unroll
            for (int a = 0; a < 2; ++a)
              if (a < p.ca.nbox) tc_tma<CG>(sa + a * 8192, &map_a, fb, a0 + a * p.ca.e0, a1, a2 + a * p.ca.e2);
#pragma unroll
            for (int a = 0; a < 4; ++a)
              if (a < p.cb.nbox) tc_tma<CG>(sa + TC_A_STAGE_BYTES + a * 8192, &map_b, fb, b0 + a * p.cb.e0, b1, b2 + a * p.cb.e2);
          }
          __syncwarp();
          a0 += p.ca.d0; a1 += p.ca.d1; a2 += p.ca.d2;
          b0 += p.cb.d0; b1 += p.cb.d1; b2 += p.cb.d2;
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (CG == 2 && !(dbg & 2)) {
        // tail: the leader's last multicast commits must have landed in this CTA's `empty` barriers before it may exit
        for (int s = 0; s < Cfg::STAGES; ++s) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // One elected lane issues; the warp stays converged so descriptor arithmetic lives in uniform registers.  The
    // 64-bit smem descriptors are a constant high word plus a low word that advances by (bytes >> 4) per UMMA_K step:
    // the issue loop must stay far below the 2*BN tensor clocks one k-block takes, or it -- not the tensor pipe --
    // bounds the kernel.
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      TcWorkIter wi;
      wi.init(p, (int)blockIdx.x / CG, (int)gridDim.x / CG);
      TcWork w;
      const bool leader = elect_one();
      const uint32_t a_hi = ((p.a_sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);   // SBO | version 1 | SWIZZLE_128B
      const uint32_t b_hi = ((p.b_sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = ((smem_base >> 4) & 0x3FFFu) | (((p.a_lbo >> 4) & 0x3FFFu) << 16);
      const uint32_t b_lo0 = (((smem_base + TC_A_STAGE_BYTES) >> 4) & 0x3FFFu) | (((p.b_lbo >> 4) & 0x3FFFu) << 16);
      const uint32_t a_ks = p.a_kstep >> 4, b_ks = p.b_kstep >> 4;
      while (wi.next(p, w)) {
        const int nkb = w.kb1 - w.kb0;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          if (!(dbg & 2)) mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (leader) {
            const uint32_t so = (uint32_t)stage * (Cfg::STAGE_BYTES >> 4);
            const uint32_t al = a_lo0 + so, bl = b_lo0 + so;
            tc_mma_desc(d_tmem, al, a_hi, bl, b_hi, p.idesc, kb > 0 ? 1u : 0u, CG == 2);
            tc_mma_desc(d_tmem, al + a_ks, a_hi, bl + b_ks, b_hi, p.idesc, 1u, CG == 2);
            tc_mma_desc(d_tmem, al + 2 * a_ks, a_hi, bl + 2 * b_ks, b_hi, p.idesc, 1u, CG == 2);
            tc_mma_desc(d_tmem, al + 3 * a_ks, a_hi, bl + 3 * b_ks, b_hi, p.idesc, 1u, CG == 2);
            // smem slot free (in both CTAs) once these MMAs retire
            if (!(dbg & 2)) { if constexpr (CG == 2) tc_commit_2cta(empty_bar(stage)); else tc_commit(empty_bar(stage)); }
            if (kb + 1 == nkb) {   // accumulator complete (both CTAs' epilogues)
              if constexpr (CG == 2) tc_commit_2cta(tfull_bar(acc)); else tc_commit(tfull_bar(acc));
            }
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int we = warp - TC_EPI_WARP0;
    const int q = warp & 3;      // TMEM lane quarter this warp may touch
    const int half = we >> 2;    // which interleaved set of 32-column chunks
    int acc = 0;
    uint32_t acc_phase = 0;
    TcWorkIter wi;
    wi.init(p, (int)blockIdx.x / CG, (int)gridDim.x / CG);
    TcWork w;
    auto release_tmem = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(tempty_bar(acc), 0));
        else mbar_arrive(tempty_bar(acc));
      }
    };

    if constexpr (EK == EK_LEGACY) {
      while (wi.next(p, w)) {
        const int i0 = w.tm * (TC_BM * CG) + (int)rank * TC_BM, j0 = w.tn * BN;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
        const uint32_t tb = smem_base + Cfg::EPI_OFF + (uint32_t)we * Cfg::EPI_WARP_BYTES;   // 4 KB transpose buffer
        const int sub_row = lane >> 3, c4 = lane & 7;
        if (i0 + q * 32 < p.I && !(dbg & 1)) {   // warps whose 32 rows are all past the end of the matrix have nothing to write
#pragma unroll 1
          for (int c = half; c < BN / 32; c += TC_EPI_WARPS / 4) {
            uint32_t raw[32];
            tc_ld32(taddr + c * 32, raw);
            // row `lane` of the 32x32 chunk -> swizzled smem (16-byte column group cg at cg ^ (row & 7)), then 8 lanes
            // re-read 32 consecutive columns of ONE row: coalesced global IO in epilogue_rows8
#pragma unroll
            for (int cg = 0; cg < 8; ++cg)
              sts128(tb + row_off_128(lane, cg), raw[cg * 4], raw[cg * 4 + 1], raw[cg * 4 + 2], raw[cg * 4 + 3]);
            __syncwarp();
            float4 v[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int r = it * 4 + sub_row;
              const uint4 u = lds128(tb + row_off_128(r, c4));
              v[it] = make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
            }
            __syncwarp();
            epilogue_rows8(p.ep, i0 + q * 32 + sub_row, j0 + c * 32 + c4 * 4, p.I, v);
          }
        }
        release_tmem();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    } else {
      // ---------------- TMA epilogue ----------------
      constexpr bool kLoads = (EK == EK_RESIDUAL || EK == EK_GELU_BWD);       // second operand TMA-loaded into the staging tile
      constexpr bool kF32 = (EK == EK_STORE_F32 || EK == EK_RESIDUAL);   // fp32 tiles: 4 KB, 128-byte rows
      constexpr uint32_t kTileBytes = kF32 ? 4096u : 2048u;
      constexpr int NCH = BN / 64;                                             // 32-column chunks per warp and tile
      constexpr uint32_t kSlotBytes = kF32 ? 4096u : 2048u;
      constexpr bool kOneSlot = kF32 && Cfg::EPI_WARP_BYTES == 4096u;          // see tc_epi_warp_bytes
      const uint32_t stg = smem_base + Cfg::EPI_OFF + (uint32_t)we * Cfg::EPI_WARP_BYTES;   // slot s at stg + s * kSlotBytes
      const uint32_t sbias = smem_base + Cfg::BIAS_OFF + (uint32_t)we * 512;
      uint32_t eph = 0;            // phase bits of this warp's two load barriers
      uint32_t slot = 0;           // staging slot of the next chunk (alternates across tiles too)
      while (wi.next(p, w)) {
        const int i0 = w.tm * (TC_BM * CG) + (int)rank * TC_BM, j0 = w.tn * BN;
        const int row0 = i0 + q * 32;
        const bool active = row0 < p.I;   // warps whose 32 rows are all past the end of the matrix have nothing to do
        if (p.streamk == 3 && w.tile >= p.n_whole) {
          // ---- partial item of a tail tile (always before this cluster's whole tiles): this warp's 32 rows x BN/2 columns
          // of the partial accumulator are added to the fp32 scratch; the warp that completes a region -- the last of the
          // clusters sharing the tile to arrive at the region's ticket -- reads the sums back, clears them and applies
          // the epilogue with per-thread global IO (epilogue_rows8: the same arithmetic as the fused TMA epilogue).
          if (lane == 0) bulk_wait_read<0>();      // the staging tiles double as the transpose buffer (no store is reading them)
          __syncwarp();
          mbar_wait(tfull_bar(acc), acc_phase);
          tc_fence_after();
          if (active) {
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            const int sub_row = lane >> 3, c4 = lane & 7;
            float* const srow = p.tail_acc + (int64_t)(row0 + sub_row - p.tail_row0) * p.J + j0 + c4 * 4;
#pragma unroll 1
            for (int c = half; c < BN / 32; c += TC_EPI_WARPS / 4) {
              uint32_t raw[32];
              tc_ld32(taddr + c * 32, raw);
#pragma unroll
              for (int cg = 0; cg < 8; ++cg)
                sts128(stg + row_off_128(lane, cg), raw[cg * 4], raw[cg * 4 + 1], raw[cg * 4 + 2], raw[cg * 4 + 3]);
              __syncwarp();
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const uint4 u = lds128(stg + row_off_128(it * 4 + sub_row, c4));
                if (row0 + sub_row + it * 4 < p.I)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(srow + (int64_t)(it * 4) * p.J + c * 32),
                               "f"(__uint_as_float(u.x)), "f"(__uint_as_float(u.y)), "f"(__uint_as_float(u.z)), "f"(__uint_as_float(u.w)) : "memory");
              }
              __syncwarp();
            }
          }
          release_tmem();
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          if (active) {
            // contributors of this tile: the clusters whose unit ranges intersect [tt * kb_total, (tt + 1) * kb_total)
            const int tt = w.tile - p.n_whole;
            const int64_t u0 = (int64_t)tt * p.kb_total;
            const int ncontrib = (int)((u0 + p.kb_total - 1) / p.units_per_cta - u0 / p.units_per_cta) + 1;
            __threadfence();
            __syncwarp();
            int last = 0;
            if (lane == 0) {
              int* ticket = p.tail_tickets + (tt * CG + (int)rank) * TC_EPI_WARPS + we;
              last = (atomicAdd(ticket, 1) + 1 == ncontrib);
              if (last) *ticket = 0;
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
              __threadfence();
              const int sub_row = lane >> 3, c4 = lane & 7;
              float* const srow = p.tail_acc + (int64_t)(row0 + sub_row - p.tail_row0) * p.J + j0 + c4 * 4;
#pragma unroll 1
              for (int c = half; c < BN / 32; c += TC_EPI_WARPS / 4) {
                float4 v[8];
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                  float4* sp = reinterpret_cast<float4*>(srow + (int64_t)(it * 4) * p.J + c * 32);
                  if (row0 + sub_row + it * 4 < p.I) {
                    v[it] = __ldcg(sp);
                    __stcg(sp, make_float4(0.f, 0.f, 0.f, 0.f));
                  } else {
                    v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                  }
                }
                epilogue_rows8(p.ep, row0 + sub_row, j0 + c * 32 + c4 * 4, p.I, v);
              }
            }
          }
          fence_proxy_async_smem();   // generic-proxy use of the staging tiles is ordered before the TMA loads / stores that follow
          __syncwarp();
          continue;
        }
        // ---- before the accumulator is ready: bias values of this warp's columns, and the first second-operand tile
        float bv[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k) bv[k] = p.ep.bias ? __ldg(p.ep.bias + j0 + (half + 2 * k) * 32 + lane) : 0.f;
        auto issue_load = [&](int k, uint32_t s) {   // lane 0: second operand of chunk k -> staging slot s
          mbar_expect_tx(eload_bar(we, s), kTileBytes);
          tma_load_2d(stg + s * kSlotBytes, &map_d, eload_bar(we, s), j0 + (half + 2 * k) * 32, row0);
        };
        if constexpr (kLoads) {
          if (active && lane == 0) {
            bulk_wait_read<0>();       // the store that last read this slot has drained it
            issue_load(0, slot);
          }
        }
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
        if (!active || (dbg & 1)) {
          release_tmem();
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
          continue;
        }
#pragma unroll
        for (int k = 0; k < NCH; ++k) asm volatile("st.shared.f32 [%0], %1;" ::"r"(sbias + (k * 32 + lane) * 4), "f"(bv[k]) : "memory");
        __syncwarp();
        uint32_t raw[32];
        tc_ld32_issue(taddr + half * 32, raw);
#pragma unroll 1
        for (int k = 0; k < NCH; ++k) {
          const int col = j0 + (half + 2 * k) * 32;
          const uint32_t s = (EK == EK_GELU || kOneSlot) ? 0u : slot;   // EK_GELU: both 2 KB slots hold one chunk (g | g')
          if constexpr (!kOneSlot) slot ^= 1;
          if constexpr (kLoads && !kOneSlot) {
            if (k + 1 < NCH && lane == 0) {
              bulk_wait_read<0>();     // chunk k-1's store (slot s^1) has been read out
              issue_load(k + 1, s ^ 1);
            }
          }
          tc_ld_wait();
          if (k + 1 == NCH) release_tmem();   // accumulator fully in registers: the MMA warp may reuse this TMEM buffer
          // ---- epilogue math on this lane's row: 32 columns
          float v[32];
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            float4 b;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "r"(sbias + (k * 32 + c4 * 4) * 4));
            v[c4 * 4 + 0] = __uint_as_float(raw[c4 * 4 + 0]) + b.x;
            v[c4 * 4 + 1] = __uint_as_float(raw[c4 * 4 + 1]) + b.y;
            v[c4 * 4 + 2] = __uint_as_float(raw[c4 * 4 + 2]) + b.z;
            v[c4 * 4 + 3] = __uint_as_float(raw[c4 * 4 + 3]) + b.w;
          }
          if (k + 1 < NCH) tc_ld32_issue(taddr + (half + 2 * (k + 1)) * 32, raw);   // next chunk's TMEM read overlaps the rest
          const uint32_t t0 = stg + s * kSlotBytes;
          // the staging slot is reusable once the TMA store that last read it has drained it (two chunks ago; EK_GELU: the
          // previous chunk) -- checked as late as possible, after this chunk's math, so the drain overlaps it
          auto slot_ready = [&]() {
            if (lane == 0) { if constexpr (EK == EK_GELU || kOneSlot) bulk_wait_read<0>(); else bulk_wait_read<1>(); }
            __syncwarp();
          };
          if constexpr (kLoads) {
            mbar_wait(eload_bar(we, s), (eph >> s) & 1u);
            eph ^= 1u << s;
          }
          if constexpr (EK == EK_STORE_BF16 || EK == EK_SCATTER) {
            uint32_t ob[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) ob[e] = pack_bf16x2(v[2 * e], v[2 * e + 1]);
            slot_ready();
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) sts128(t0 + row_off_64(lane, k4), ob[k4 * 4], ob[k4 * 4 + 1], ob[k4 * 4 + 2], ob[k4 * 4 + 3]);
            if constexpr (EK == EK_STORE_BF16) {
              // optional column sums of the bf16 output (no bias in this mode: rows past the end of the matrix are zero)
              if (p.ep.colsum) { __syncwarp(); tile_colsum_bf16(t0, lane, p.ep.colsum + col); }
            }
          } else if constexpr (EK == EK_STORE_F32) {
            slot_ready();
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8)
              sts128(t0 + row_off_128(lane, k8), __float_as_uint(v[k8 * 4 + 0]), __float_as_uint(v[k8 * 4 + 1]),
                     __float_as_uint(v[k8 * 4 + 2]), __float_as_uint(v[k8 * 4 + 3]));
          } else if constexpr (EK == EK_GELU) {
            // GELU acts on the 16-bit fc1 output (autocast); gelu'(u) is kept for the backward epilogue
            uint32_t g[16], dg[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const float2 ur = unpack_bf16x2(pack_bf16x2(v[2 * e], v[2 * e + 1]));
              float g0, d0, g1, d1;
              gelu_fast_both(ur.x, g0, d0);
              gelu_fast_both(ur.y, g1, d1);
              g[e] = pack_bf16x2(g0, g1);
              dg[e] = pack_bf16x2(d0, d1);
            }
            slot_ready();
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              sts128(t0 + row_off_64(lane, k4), g[k4 * 4], g[k4 * 4 + 1], g[k4 * 4 + 2], g[k4 * 4 + 3]);
              sts128(t0 + 2048 + row_off_64(lane, k4), dg[k4 * 4], dg[k4 * 4 + 1], dg[k4 * 4 + 2], dg[k4 * 4 + 3]);
            }
          } else if constexpr (EK == EK_RESIDUAL) {
#pragma unroll
            for (int k8 = 0; k8 < 8; ++k8) {
              const uint32_t a = t0 + row_off_128(lane, k8);
              const uint4 r = lds128(a);
              sts128(a, __float_as_uint(__uint_as_float(r.x) + v[k8 * 4 + 0]), __float_as_uint(__uint_as_float(r.y) + v[k8 * 4 + 1]),
                     __float_as_uint(__uint_as_float(r.z) + v[k8 * 4 + 2]), __float_as_uint(__uint_as_float(r.w) + v[k8 * 4 + 3]));
            }
          } else if constexpr (EK == EK_GELU_BWD) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint32_t a = t0 + row_off_64(lane, k4);
              const uint4 u = lds128(a);   // gelu'(u), saved by the forward epilogue
              const float2 u0 = unpack_bf16x2(u.x), u1 = unpack_bf16x2(u.y), u2 = unpack_bf16x2(u.z), u3 = unpack_bf16x2(u.w);
              sts128(a, pack_bf16x2(v[k4 * 8 + 0] * u0.x, v[k4 * 8 + 1] * u0.y), pack_bf16x2(v[k4 * 8 + 2] * u1.x, v[k4 * 8 + 3] * u1.y),
                     pack_bf16x2(v[k4 * 8 + 4] * u2.x, v[k4 * 8 + 5] * u2.y), pack_bf16x2(v[k4 * 8 + 6] * u3.x, v[k4 * 8 + 7] * u3.y));
            }
            if (p.ep.colsum) {
              // bias gradient of fc1 = column sums of the bf16 values just written (rows past the end are zero: their
              // gelu' tile was zero-filled by the TMA load)
              __syncwarp();
              tile_colsum_bf16(t0, lane, p.ep.colsum + col);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            if constexpr (EK == EK_SCATTER) {
              tma_store_3d(&map_c, t0, col & 63, row0, col >> 6);
            } else {
              tma_store_2d(&map_c, t0, col, row0);
              if constexpr (EK == EK_GELU) {
                if (p.ep.aux) tma_store_2d(&map_d, t0 + 2048, col, row0);
              }
            }
            bulk_commit();
            if constexpr (kLoads && kOneSlot) {
              if (k + 1 < NCH) {         // single slot: the next chunk's second operand may land once this store has read the slot
                bulk_wait_read<0>();
                issue_load(k + 1, 0u);
              }
            }
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (lane == 0) bulk_wait_all();   // shared memory must outlive the last stores' reads; writes complete before exit
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // nobody exits while its peer may still touch it
  trace_mark(TK_GEMM_TC, 2, tr_aux);
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
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.rank = 3; key.dtype = VITK_BF16; key.swizzle = 128; key.l2promo = 256;
  for (int i = 0; i < 3; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
  key.strides[0] = strides[0]; key.strides[1] = strides[1];
  if (tmap_cache_get(key, map)) return VITK_OK;
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS) tmap_cache_put(key, map);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed: CUresult %d (dims %llu,%llu,%llu strides %llu,%llu box %u,%u,%u)", (int)r,
              (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
              (unsigned long long)strides[0], (unsigned long long)strides[1], box[0], box[1], box[2]);
    return VITK_ERR_DRIVER;
  }
  return VITK_OK;
}


// Tuning overrides (vitk_debug_set): process-wide, for tests and A/B timing; every one of them yields valid results.
//   1 whole-K tiles for accumulate GEMMs, 2 forced BLOCK_N, 4 CTA group, 5 per-thread epilogue IO, 6 no programmatic
//   dependent launch, 9 split tail (1: off, n > 1: minimum reduction depth in k-blocks instead of 24), 13 stream-K instead of sliced split-K (> 1: fill threshold in percent).
// Development build only (libvitk_dev.so): 0 swap LBO/SBO of MN-major operands, 7 timing-only bit mask (results INVALID),
//   12 whole qkv bias gradient from the attention kernel.
static int g_tc_debug[16] = {0};
static bool knob_allowed(int key) {
#ifdef VITK_DEV
  return key >= 0 && key < 16;
#else
  return key == 1 || key == 2 || key == 4 || key == 5 || key == 6 || key == 9 || key == 13;
#endif
}
int tune_knob(int key) { return knob_allowed(key) ? g_tc_debug[key] : 0; }

// 32 x 32 element tile maps of the epilogue operands: row-major [rows][ld] (2-D) or head-major [C/64][rows][64] (3-D).
// 4-byte elements -> 128-byte tile rows (SWIZZLE_128B), 2-byte -> 64-byte rows (SWIZZLE_64B).
static int make_tile_map(const void* base, int dtype, int64_t cols, int64_t rows, int64_t ld, int64_t hm_rows, CUtensorMap* map) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return VITK_ERR_DRIVER; }
  const cuuint64_t elt = dtype == VITK_BF16 ? 2 : 4;
  cuuint64_t dims[3], strides[2];
  cuuint32_t box[3] = {32, 32, 1}, estr[3] = {1, 1, 1};
  int rank = 2;
  if (hm_rows > 0) {
    rank = 3;
    dims[0] = 64; dims[1] = (cuuint64_t)hm_rows; dims[2] = (cuuint64_t)(cols / 64);
    strides[0] = 64 * elt; strides[1] = (cuuint64_t)hm_rows * 64 * elt;
  } else {
    dims[0] = (cuuint64_t)cols; dims[1] = (cuuint64_t)rows; dims[2] = 1;
    strides[0] = (cuuint64_t)ld * elt; strides[1] = strides[0] * (cuuint64_t)rows;
  }
  if (((uintptr_t)base & 15) || (strides[0] & 15)) { set_error("gemm_tc: epilogue operand must be 16-byte aligned"); return VITK_ERR_ARG; }
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.rank = (uint32_t)rank; key.dtype = (uint32_t)dtype; key.swizzle = dtype == VITK_BF16 ? 64 : 128; key.l2promo = 128;
  for (int i = 0; i < 3; ++i) { key.dims[i] = i < rank ? dims[i] : 0; key.box[i] = box[i]; }
  key.strides[0] = strides[0]; key.strides[1] = rank == 3 ? strides[1] : 0;
  if (tmap_cache_get(key, map)) return VITK_OK;
  const CUresult r = enc(map, dtype == VITK_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank,
                         const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         dtype == VITK_BF16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (epilogue tile) failed: CUresult %d", (int)r); return VITK_ERR_DRIVER; }
  tmap_cache_put(key, map);
  return VITK_OK;
}

static int epilogue_kind(const EpiParams& ep) {
  if (g_tc_debug[5] == 1) return EK_LEGACY;   // debug: per-thread global IO epilogue for every mode
  const bool ok16 = (((uintptr_t)ep.out & 15) == 0) && (ep.ldc % 8 == 0);
  switch (ep.mode) {
    case E_STORE: return !ok16 ? EK_LEGACY : (ep.out_dtype == VITK_BF16 ? EK_STORE_BF16 : EK_STORE_F32);
    case E_BIAS_GELU: return (ok16 && ep.out_dtype == VITK_BF16 && (((uintptr_t)ep.aux & 15) == 0)) ? EK_GELU : EK_LEGACY;
    case E_BIAS_RESIDUAL: return (ok16 && (((uintptr_t)ep.residual & 15) == 0)) ? EK_RESIDUAL : EK_LEGACY;
    case E_QKV_SCATTER: return (ok16 && ep.out_dtype == VITK_BF16) ? EK_SCATTER : EK_LEGACY;
    // weight-gradient partial tiles: the transposing red.global.add.v4 path (a TMA reduce-add epilogue measured 2-3 %
    // slower on all four shapes in round 1 and was removed)
    case E_ACCUM: return EK_LEGACY;
    case E_GELU_BWD: return (ok16 && ep.out_dtype == VITK_BF16 && ep.aux && (((uintptr_t)ep.aux & 15) == 0)) ? EK_GELU_BWD : EK_LEGACY;
    default: return EK_LEGACY;
  }
}

// ---- split tail (pure host logic).  M = B*197 rows rarely fill the last wave of the persistent grid: 12608 x 768 on 74 CTA
// pairs is 150 tiles of 256 x 256 = two full waves and a third one with TWO tiles in it -- a third of the kernel's time for
// 1.3 % of its work.  When the last wave is at most a quarter full and the reduction is deep enough (>= 24 k-blocks: fc2
// forward, fc1 dgrad, qkv dgrad), the tiles of the full waves stay whole-K items with the fused epilogue and the k-blocks of
// the remaining tiles are dealt out to ALL clusters as contiguous ranges (decomposition mode 3); partial accumulators meet
// in an fp32 scratch and the last cluster to arrive at a region's ticket applies the epilogue (kernel: "partial item").
// The caller provides the scratch (GemmProblem::tail_scratch, all zero between launches): TC_TAIL_TICKETS ints of tickets,
// then fp32 [rows of the tail tiles][J].
constexpr int TC_TAIL_TICKETS = 1024;
struct TcTailPlan { int n_whole, tail_row0; int64_t units_per_cluster; };
// f32_out: the epilogue kind stages fp32 tiles (tc_epi_warp_bytes).  A split tail never justifies an operand ring shallower than
// 5 stages (fc2 forward at bs 64 with a 4-stage ring: 73 us in-step against 66 us for three waves of 256 x 192 tiles).
static bool tc_tail_plan(int I, int J, int R, int BN, int CG, bool f32_out, int64_t scratch_floats, TcTailPlan* tp) {
  if (g_tc_debug[9] == 1 || scratch_floats <= TC_TAIL_TICKETS) return false;
  {
    const int tail_bytes = TC_EPI_WARPS * (int)tc_epi_warp_bytes(f32_out ? EK_STORE_F32 : EK_STORE_BF16, BN, CG) + TC_EPI_WARPS * 512 + 512;
    const int stage_bytes = (int)TC_A_STAGE_BYTES + (BN / CG) * TC_BK * 2;
    if ((227 * 1024 - 1024 - tail_bytes) / stage_bytes < 5) return false;
  }
  const int kb = (R + TC_BK - 1) / TC_BK;
  const int rows_per_tile = TC_BM * CG;
  const int tiles_n = J / BN, tiles_m = (I + rows_per_tile - 1) / rows_per_tile;
  const int tiles = tiles_m * tiles_n, slots = sm_count() / CG;
  const int full = tiles / slots, rem = tiles % slots;
  const int min_kb = g_tc_debug[9] > 1 ? g_tc_debug[9] : 24;     // knob 9 = n > 1: A/B of the depth threshold
  if (kb < min_kb || full < 1 || rem == 0 || rem * 4 > slots) return false;
  const int n_whole = tiles - rem;
  const int tail_row0 = (n_whole / tiles_n) * rows_per_tile;
  if ((int64_t)(I - tail_row0) * J > scratch_floats - TC_TAIL_TICKETS) return false;
  if (rem * CG * TC_EPI_WARPS > TC_TAIL_TICKETS) return false;
  tp->n_whole = n_whole;
  tp->tail_row0 = tail_row0;
  tp->units_per_cluster = ((int64_t)rem * kb + slots - 1) / slots;
  return true;
}

// ---- work decomposition (pure host logic; p.I / p.J / p.R set).  Returns the number of clusters (CTAs / CG) to launch.
// tail_floats > 0: a non-accumulating GEMM whose epilogue mode permits the split tail, with that much scratch.
static int tc_decompose(TcParams& p, bool accumulate, int BN, int CG, int64_t tail_floats = 0, bool f32_out = false) {
  p.n_tiles_m = (p.I + TC_BM * CG - 1) / (TC_BM * CG);
  p.n_tiles_n = p.J / BN;
  p.kb_total = (p.R + TC_BK - 1) / TC_BK;
  const int tiles = p.n_tiles_m * p.n_tiles_n;
  const int slots = sm_count() / CG;            // persistent CTAs (CTA pairs)
  int grid = tiles < slots ? tiles : slots;
  p.streamk = 0;
  p.units_per_cta = 0;
  p.n_whole = tiles;
  p.tail_row0 = 0;
  TcTailPlan tp;
  if (!accumulate && tail_floats > 0 && tc_tail_plan(p.I, p.J, p.R, BN, CG, f32_out, tail_floats, &tp)) {
    p.streamk = 3;
    p.n_whole = tp.n_whole;
    p.tail_row0 = tp.tail_row0;
    p.units_per_cta = tp.units_per_cluster;
    return slots;
  }
  if (accumulate && g_tc_debug[1] != 1) {
    const int64_t total = (int64_t)tiles * p.kb_total;
    grid = total < slots ? (int)total : slots;
    p.streamk = 1;
    p.units_per_cta = (total + grid - 1) / grid;
    grid = (int)((total + p.units_per_cta - 1) / p.units_per_cta);
    // Sliced split-K (knob 13 = 1 disables; 13 = p > 1: fill threshold in percent): when tiles x S fills >= 90 % of the
    // slots, every CTA takes ONE k-slice of ONE tile.  All CTAs of a slice then sweep the same k range in lock-step -- the
    // A panel of a tile row and the B panel of a tile column are fetched from DRAM once and hit in L2 for the other tiles
    // -- and each CTA drains one accumulator instead of the two or three partial tiles a contiguous stream-K range straddles.
    const int S = tiles > 0 ? slots / tiles : 0;
    const int fill_pct = g_tc_debug[13] > 1 ? g_tc_debug[13] : 90;
    if (g_tc_debug[13] != 1 && S >= 1 && tiles * S * 100 >= slots * fill_pct && p.kb_total >= 2 * S) {
      p.streamk = 2;
      p.units_per_cta = S;
      grid = tiles * S;
    }
  }
  return grid;
}

// ---- tile shape (pure host logic): CTA pair (CG = 2, 256 x BN) or single CTA (128 x BN), BN in {256, 192, 128}: minimise
//   (waves of the persistent grid) x (k-blocks x max(tensor clocks, operand-ingest clocks) + per-tile overhead).
// A 64-deep k-block costs 2*BN tensor clocks and pulls (128 + BN/CG) x 128 B into the SM; measured with the epilogue
// and the loads switched off in turn (vitk_debug_set(7, .)), the mainloop sustains ~36 B/clk/SM of operand traffic
// -- it, not the tensor pipe, bounds every shape here -- so wider tiles win unless they cost a whole extra wave: with
// M = B*197 the tile count is rarely a multiple of the slot count (12608 x 768 on 74 CTA pairs: 150 tiles of 256x256 =
// 3 waves for 2.03 waves of work, 200 tiles of 256x192 = 3 waves of 7/8 the bytes).  With a tail scratch (tail_floats > 0) a
// nearly empty last wave costs only its share of k-blocks plus a fixed fix-up (tc_tail_plan): 150 tiles = 2 waves + 2 k-blocks.
static void tc_pick_tile(int I, int J, int R, bool accumulate, bool b_mn, int* bn_out, int* cg_out, int64_t tail_floats = 0,
                         bool f32_out = false) {
  int cg = 0, bn = 0;
  const int forced_bn = (g_tc_debug[2] == 128 || g_tc_debug[2] == 192 || g_tc_debug[2] == 256) ? g_tc_debug[2] : 0;
  if (accumulate) {
    bn = forced_bn ? forced_bn : (J % 256 == 0 ? 256 : 128);  // split-K / stream-K balance by themselves
    cg = (I > TC_BM && g_tc_debug[4] != 1 && !(b_mn && (bn / 2) % 64 != 0)) ? 2 : 1;
  } else {
    const long kb = (R + TC_BK - 1) / TC_BK;
    double best = -1.0;
    for (int c : {2, 1}) {
      if (c == 2 && (I <= TC_BM || g_tc_debug[4] == 1)) continue;
      if (c == 1 && g_tc_debug[4] == 2 && best >= 0.0) continue;   // debug: pairs forced (when expressible)
      const long slots = sm_count() / c;
      const long tiles_m = (I + TC_BM * c - 1) / (TC_BM * c);
      for (int cand : {256, 192, 128}) {
        if (J % cand != 0) continue;
        if (b_mn && (cand / c) % 64 != 0) continue;
        if (forced_bn && forced_bn != cand) continue;
        const long tiles = tiles_m * (J / cand);
        const long waves = (tiles + slots - 1) / slots;
        const double ingest = (128.0 + cand / c) * 128.0 / 36.0, mma = 2.0 * cand;
        const double per_kb = ingest > mma ? ingest : mma;
        double cost = (double)waves * ((double)kb * per_kb + 1200.0);
        TcTailPlan tp;
        if (tail_floats > 0 && tc_tail_plan(I, J, R, cand, c, f32_out, tail_floats, &tp))
          cost = (double)(tiles / slots) * ((double)kb * per_kb + 1200.0) + (double)tp.units_per_cluster * per_kb + 6000.0;
        if (best < 0.0 || cost < best) { best = cost; bn = cand; cg = c; }
      }
    }
  }
  *bn_out = bn;
  *cg_out = cg;
}

// epilogue modes whose partial items the kernel's per-thread fix-up path may finish (epilogue_rows8 implements every mode; the
// epilogues with fused column sums stay whole-tile only)
static int64_t plan_tail_floats(int J) { return (int64_t)TC_TAIL_TICKETS + (int64_t)512 * J; }
static int64_t tail_floats_of(const GemmProblem& pr) {
  if (!pr.tail_scratch || pr.ep.colsum || ((uintptr_t)pr.tail_scratch & 15)) return 0;
  if (!(pr.ep.mode == E_BIAS_RESIDUAL || pr.ep.mode == E_STORE || pr.ep.mode == E_BIAS_GELU || pr.ep.mode == E_QKV_SCATTER)) return 0;
  return pr.tail_scratch_floats;
}

template <int BN, int CG, int EK>
static int launch_tc(const GemmProblem& pr, cudaStream_t st) {
  using Cfg = TcCfg<BN, CG, EK>;
  VITK_TRY(set_max_dyn_smem_once((const void*)gemm_tc_kernel<BN, CG, EK>, (int)Cfg::SMEM_BYTES));
  CUtensorMap map_a, map_b, map_c, map_d;
  TcParams p{};
  p.I = pr.I; p.J = pr.J; p.R = pr.R;
#ifdef VITK_DEV
  p.dbg = g_tc_debug[7];
#endif
  VITK_TRY(make_operand_map(pr.A, pr.la, pr.I, pr.R, TC_BM, &map_a, &p.a_mode));
  VITK_TRY(make_operand_map(pr.B, pr.lb, pr.J, pr.R, Cfg::B_ROWS, &map_b, &p.b_mode));
  const EpiParams& ep = pr.ep;
  if constexpr (EK == EK_LEGACY) {
    map_c = map_a; map_d = map_a;   // unused
  } else if constexpr (EK == EK_SCATTER) {
    VITK_TRY(make_tile_map(ep.out, VITK_BF16, pr.J, pr.I, 0, ep.hm_rows, &map_c));
    map_d = map_c;
  } else {
    const int odt = (EK == EK_STORE_F32 || EK == EK_RESIDUAL) ? VITK_F32 : VITK_BF16;
    VITK_TRY(make_tile_map(ep.out, odt, pr.J, pr.I, ep.ldc, 0, &map_c));
    if constexpr (EK == EK_RESIDUAL) VITK_TRY(make_tile_map(ep.residual, VITK_F32, pr.J, pr.I, ep.ldc, 0, &map_d));
    else if constexpr (EK == EK_GELU_BWD) VITK_TRY(make_tile_map(ep.aux, VITK_BF16, pr.J, pr.I, ep.ldc, 0, &map_d));
    else if constexpr (EK == EK_GELU) { if (ep.aux) VITK_TRY(make_tile_map(ep.aux, VITK_BF16, pr.J, pr.I, ep.ldc, 0, &map_d)); else map_d = map_c; }
    else map_d = map_c;
  }
  auto fill_coord = [](int mode, int rows_in_tile, OpCoord* c) {
    *c = OpCoord{0, 0, 0, 0, 0, 0, 0, 0, 1};
    if (mode == OP_KM_FLAT) { c->fr1 = 1; c->d0 = TC_BK; }
    else if (mode == OP_KM_SPLIT) { c->fr1 = 1; c->d2 = 1; }
    else if (mode == OP_MN_FLAT) { c->fr0 = 1; c->d1 = TC_BK; c->e0 = 64; c->nbox = rows_in_tile / 64; }
    else { c->fr2 = 1; c->d1 = TC_BK; c->e2 = 1; c->nbox = rows_in_tile / 64; }
  };
  fill_coord(p.a_mode, TC_BM, &p.ca);
  fill_coord(p.b_mode, Cfg::B_ROWS, &p.cb);
  const bool a_mn = p.a_mode >= OP_MN_FLAT, b_mn = p.b_mode >= OP_MN_FLAT;
  // instruction descriptor: D=f32, A=B=bf16, majors, N>>3, M>>4 (M = 128 per CTA: 256 for a CTA pair)
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
            ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((TC_BM * CG) >> 4) << 24);
  // canonical SWIZZLE_128B layouts: K-major: 8-row groups 1024 B apart (SBO), LBO unused;
  // MN-major: 8-r groups 1024 B apart (SBO), 64-wide MN atoms TC_BK*128 B apart (LBO)
  p.a_sbo = 1024; p.a_lbo = a_mn ? TC_BK * 128 : 16; p.a_kstep = a_mn ? 16 * 128 : 32;
  p.b_sbo = 1024; p.b_lbo = b_mn ? TC_BK * 128 : 16; p.b_kstep = b_mn ? 16 * 128 : 32;
#ifdef VITK_DEV
  if (g_tc_debug[0] == 1) {  // debug variant: swap LBO/SBO roles of MN-major operands
    if (a_mn) { p.a_sbo = TC_BK * 128; p.a_lbo = 1024; }
    if (b_mn) { p.b_sbo = TC_BK * 128; p.b_lbo = 1024; }
  }
#endif
  const int grid = tc_decompose(p, pr.ep.mode == E_ACCUM, BN, CG, EK == EK_LEGACY ? 0 : tail_floats_of(pr),
                                EK == EK_STORE_F32 || EK == EK_RESIDUAL);
  p.tail_tickets = reinterpret_cast<int*>(pr.tail_scratch);
  p.tail_acc = pr.tail_scratch ? pr.tail_scratch + TC_TAIL_TICKETS : nullptr;
  p.ep = pr.ep;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid * CG, 1, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CG; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  count_launch();
  VITK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CG, EK>, map_a, map_b, map_c, map_d, p));
  return VITK_OK;
}

template <int BN, int CG>
static int launch_tc_kind(const GemmProblem& pr, int ek, cudaStream_t st) {
  switch (ek) {
    case EK_STORE_BF16: return launch_tc<BN, CG, EK_STORE_BF16>(pr, st);
    case EK_STORE_F32: return launch_tc<BN, CG, EK_STORE_F32>(pr, st);
    case EK_GELU: return launch_tc<BN, CG, EK_GELU>(pr, st);
    case EK_RESIDUAL: return launch_tc<BN, CG, EK_RESIDUAL>(pr, st);
    case EK_SCATTER: return launch_tc<BN, CG, EK_SCATTER>(pr, st);
    case EK_GELU_BWD: return launch_tc<BN, CG, EK_GELU_BWD>(pr, st);
    default: return launch_tc<BN, CG, EK_LEGACY>(pr, st);
  }
}

int gemm_tc(const GemmProblem& pr, cudaStream_t st) {
  VITK_CHECK_ARG(pr.I > 0 && pr.J > 0 && pr.R > 0 && pr.A && pr.B && pr.ep.out);
  if (pr.in_dtype != VITK_BF16) { set_error("gemm_tc: bf16 operands only"); return VITK_ERR_UNSUPPORTED; }
  if (pr.ep.mode != E_STORE && pr.ep.mode != E_ACCUM && pr.ep.mode != E_BIAS_RESIDUAL && pr.ep.mode != E_PATCH &&
      pr.ep.out_dtype != VITK_BF16) { set_error("gemm_tc: epilogue expects bf16 output"); return VITK_ERR_UNSUPPORTED; }
  if (pr.J % 128 != 0 || pr.ep.ldc % 8 != 0) { set_error("gemm_tc: J must be a multiple of 128 (got %d)", pr.J); return VITK_ERR_UNSUPPORTED; }
  const bool b_mn = pr.lb.s_row == 1 && pr.lb.s_col != 1;   // MN-major B: staged in 64-row atoms
  int cg = 0, bn = 0;
  const int ek = epilogue_kind(pr.ep);
  tc_pick_tile(pr.I, pr.J, pr.R, pr.ep.mode == E_ACCUM, b_mn, &bn, &cg, ek == EK_LEGACY ? 0 : tail_floats_of(pr),
               ek == EK_STORE_F32 || ek == EK_RESIDUAL);
  if (bn == 0 || pr.J % bn != 0) { set_error("gemm_tc: no BLOCK_N divides J=%d", pr.J); return VITK_ERR_UNSUPPORTED; }
  if (pr.ep.colsum && !(ek == EK_GELU_BWD || (ek == EK_STORE_BF16 && !pr.ep.bias) || pr.ep.mode == E_GELU_BWD)) {
    set_error("gemm_tc: fused column sums need the TMA epilogue of a bias-free bf16 store or of the GELU' multiply");
    return VITK_ERR_UNSUPPORTED;
  }
  if (cg == 2) {
    if (bn == 256) return launch_tc_kind<256, 2>(pr, ek, st);
    if (bn == 192) return launch_tc_kind<192, 2>(pr, ek, st);
    return launch_tc_kind<128, 2>(pr, ek, st);
  }
  if (bn == 256) return launch_tc_kind<256, 1>(pr, ek, st);
  if (bn == 192) return launch_tc_kind<192, 1>(pr, ek, st);
  return launch_tc_kind<128, 1>(pr, ek, st);
}

}  // namespace vitk

namespace vitk { void set_pdl(int on); }

// Host-only view of the tcgen05 GEMM's work decomposition (no launch, no device access beyond the SM count): tile shape,
// decomposition mode (0 whole-K tiles strided over the persistent clusters, 1 contiguous stream-K ranges, 2 sliced
// split-K) and, for one cluster, the (tile, first k-block, end k-block) items it walks -- produced by the same
// TcWorkIter the kernel's warps run.
extern "C" int vitk_gemm_plan(int I, int J, int R, int accumulate, int b_mn_major, int* block_n, int* cta_group, int* mode,
                              int* n_clusters, int* n_tiles_m, int* n_tiles_n, int* kb_total) {
  VITK_CHECK_ARG(I > 0 && J > 0 && R > 0 && J % 128 == 0);
  int bn = 0, cg = 0;
  const int64_t tail_floats = accumulate >= 2 ? vitk::plan_tail_floats(J) : 0;   // 2 / 3: no accumulation, tail scratch, bf16 / fp32 out
  vitk::tc_pick_tile(I, J, R, accumulate == 1, b_mn_major != 0, &bn, &cg, tail_floats, accumulate == 3);
  if (bn == 0 || J % bn != 0) { vitk::set_error("vitk_gemm_plan: no BLOCK_N divides J=%d", J); return VITK_ERR_UNSUPPORTED; }
  vitk::TcParams p{};
  p.I = I; p.J = J; p.R = R;
  const int grid = vitk::tc_decompose(p, accumulate == 1, bn, cg, tail_floats, accumulate == 3);
  if (block_n) *block_n = bn;
  if (cta_group) *cta_group = cg;
  if (mode) *mode = p.streamk;
  if (n_clusters) *n_clusters = grid;
  if (n_tiles_m) *n_tiles_m = p.n_tiles_m;
  if (n_tiles_n) *n_tiles_n = p.n_tiles_n;
  if (kb_total) *kb_total = p.kb_total;
  return VITK_OK;
}
// items: [max_items][3] ints (tile, kb0, kb1); returns the number of items of `cluster` (or a negative VITK_ERR_*)
extern "C" int vitk_gemm_plan_items(int I, int J, int R, int accumulate, int b_mn_major, int cluster, int* items, int max_items) {
  if (!(I > 0 && J > 0 && R > 0 && J % 128 == 0 && items && max_items > 0)) return -VITK_ERR_ARG;
  int bn = 0, cg = 0;
  const int64_t tail_floats = accumulate >= 2 ? vitk::plan_tail_floats(J) : 0;
  vitk::tc_pick_tile(I, J, R, accumulate == 1, b_mn_major != 0, &bn, &cg, tail_floats, accumulate == 3);
  if (bn == 0 || J % bn != 0) return -VITK_ERR_UNSUPPORTED;
  vitk::TcParams p{};
  p.I = I; p.J = J; p.R = R;
  const int grid = vitk::tc_decompose(p, accumulate == 1, bn, cg, tail_floats, accumulate == 3);
  if (cluster < 0 || cluster >= grid) return -VITK_ERR_ARG;
  vitk::TcWorkIter wi;
  wi.init(p, cluster, grid);
  vitk::TcWork w;
  int n = 0;
  while (wi.next(p, w)) {
    if (n < max_items) { items[3 * n] = w.tile; items[3 * n + 1] = w.kb0; items[3 * n + 2] = w.kb1; }
    ++n;
  }
  return n;
}
// Host-only view of the split tail (no launch) for a non-accumulating GEMM given a scratch of vitk_gemm_tail_scratch_floats(J)
// floats: *n_whole = tiles that stay whole-K items with the fused epilogue, *n_tail = tiles whose k-blocks are dealt out to
// all clusters (0: plain tile-strided launch), *tail_row0 = first matrix row of the first tail tile.
extern "C" int vitk_gemm_tail_plan(int I, int J, int R, int b_mn_major, int f32_out, int* n_whole, int* n_tail, int* tail_row0) {
  VITK_CHECK_ARG(I > 0 && J > 0 && R > 0 && J % 128 == 0 && n_whole && n_tail && tail_row0);
  int bn = 0, cg = 0;
  const int64_t tail_floats = vitk::plan_tail_floats(J);
  vitk::tc_pick_tile(I, J, R, false, b_mn_major != 0, &bn, &cg, tail_floats, f32_out != 0);
  if (bn == 0 || J % bn != 0) { vitk::set_error("vitk_gemm_tail_plan: no BLOCK_N divides J=%d", J); return VITK_ERR_UNSUPPORTED; }
  vitk::TcParams p{};
  p.I = I; p.J = J; p.R = R;
  vitk::tc_decompose(p, false, bn, cg, tail_floats, f32_out != 0);
  *n_whole = p.n_whole;
  *n_tail = p.n_tiles_m * p.n_tiles_n - p.n_whole;
  *tail_row0 = p.streamk == 3 ? p.tail_row0 : 0;
  return VITK_OK;
}
// floats of scratch that always suffice for a GEMM with J output columns: the tickets + 512 rows of partial sums
extern "C" size_t vitk_gemm_tail_scratch_floats(int J) { return J > 0 ? (size_t)vitk::plan_tail_floats(J) : 0; }
extern "C" int vitk_debug_set(int key, int value) {
  if (!vitk::knob_allowed(key)) {
    vitk::set_error("vitk_debug_set: key %d is not available in this build (development-only keys need libvitk_dev.so)", key);
    return VITK_ERR_ARG;
  }
  vitk::g_tc_debug[key] = value;
  if (key == 6) vitk::set_pdl(value ? 0 : 1);   // key 6: 1 = plain stream order (no programmatic dependent launch)
  return VITK_OK;
}
