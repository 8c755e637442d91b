// Microbenchmark: how many operand bytes per clock can one SM take in through TMA from an L2-resident buffer, and does
// cluster multicast raise that number?  (The tcgen05 GEMM mainloop is bound by this ingest rate, ~36 B/clk/SM measured.)
//   mode 0: unicast   -- every CTA loads its own 16 KB tile per iteration
//   mode 1: multicast -- 2-CTA clusters: each CTA loads HALF a tile (8 KB) and multicasts it to both CTAs, so every CTA
//                        still receives 16 KB per iteration but only 8 KB per CTA is requested from L2
//   mode 2: multicast over 4-CTA clusters (4 KB requested, 16 KB received per CTA and iteration)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_ingest tma_ingest.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

constexpr int STAGES = 8;
constexpr uint32_t TILE_BYTES = 128 * 128;   // 128 rows x 64 bf16

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ void arrive_remote(uint32_t a) { asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

template <int CL>
__global__ void __launch_bounds__(64, 1) ingest_kernel(const __grid_constant__ CUtensorMap map, int iters, int rows_total, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + STAGES * TILE_BYTES;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t rank = CL > 1 ? ctarank() : 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), CL); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (CL > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); } else __syncthreads();
  const int cluster = blockIdx.x / CL;
  const int n_tiles = rows_total / 128;
  unsigned long long t0 = clock64();
  if (threadIdx.x == 0) {            // producer
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(empty(stage), phase ^ 1);
      mbar_expect_tx(full(stage), TILE_BYTES);
      const int tile = (cluster * 131 + it * 17) % n_tiles;       // every cluster walks its own tile sequence
      const uint32_t dst = sbase + stage * TILE_BYTES;
      if (CL == 1) {
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&map)), "r"(full(stage)), "r"(0), "r"(tile * 128) : "memory");
      } else {
        constexpr int ROWS = 128 / CL;
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                     ::"r"(dst + rank * ROWS * 128), "l"(reinterpret_cast<uint64_t>(&map)), "r"(full(stage)), "r"(0), "r"(tile * 128 + (int)rank * ROWS),
                       "h"((uint16_t)((1u << CL) - 1)) : "memory");
      }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (threadIdx.x == 32) {    // consumer: releases the stage in every CTA of the cluster
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(full(stage), phase);
      for (int r = 0; r < CL; ++r) arrive_remote(mapa(empty(stage), r));
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  }
  __syncthreads();
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (CL > 1) { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CL> void run(const char* name, void* buf, int rows, EncodeTiledFn enc) {
  CUtensorMap map;
  cuuint64_t dims[2] = {64, (cuuint64_t)rows}, strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)(128 / CL)}, estr[2] = {1, 1};
  enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const size_t smem = 1024 + STAGES * TILE_BYTES + 256;
  cudaFuncSetAttribute(ingest_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  unsigned long long* cyc; cudaMalloc(&cyc, 148 * 8);
  const int grid = 148 / CL * CL, iters = 4000;
  cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = CL > 1 ? 1 : 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t err = cudaLaunchKernelEx(&cfg, ingest_kernel<CL>, map, iters, rows, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    if (err != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("%s: launch failed %s\n", name, cudaGetErrorString(err)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long h[148]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    if (rep == 1)
      printf("%-28s %7.3f ms  received %6.1f B/clk/SM (%5.2f TB/s chip), requested from L2 %6.1f B/clk/SM\n", name, ms,
             (double)iters * TILE_BYTES / avg, (double)grid * iters * TILE_BYTES / ms / 1e9, (double)iters * TILE_BYTES / CL / avg);
  }
}

int main() {
  void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)ptr;
  const int rows = 256 * 1024;   // 32 MB buffer: L2 resident
  void* buf; cudaMalloc(&buf, (size_t)rows * 128); cudaMemset(buf, 0, (size_t)rows * 128);
  run<1>("unicast (cluster 1)", buf, rows, enc);
  run<2>("multicast cluster 2", buf, rows, enc);
  run<4>("multicast cluster 4", buf, rows, enc);
  return 0;
}
