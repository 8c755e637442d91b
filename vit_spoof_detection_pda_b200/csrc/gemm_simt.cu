// SIMT FFMA GEMM engine: the arithmetic of VITK_PREC_FP32_VALIDATE (1e-4 parity mode needs true
// fp32 products -- bf16/TF32 tensor cores cannot give that) and a layout-general cross-check
// for the tcgen05 engine (same GemmProblem, bf16 operands, fp32 accumulate).
// 64x64x16 tiles, 256 threads, 4x4 register micro-tile; operands staged through shared memory.
#include "gemm.cuh"

namespace vitk {

constexpr int ST_BM = 64, ST_BN = 64, ST_BK = 16, ST_THREADS = 256, ST_PAD = 4;

template <typename T>
__device__ __forceinline__ void simt_load_tile(const T* __restrict__ base, const MatLayout& l, int row0, int nrows,
                                               int r0, int r_end, float (*dst)[ST_BM + ST_PAD], bool k_contig) {
#pragma unroll
  for (int e = 0; e < (ST_BM * ST_BK) / ST_THREADS; ++e) {
    const int idx = threadIdx.x + e * ST_THREADS;
    int rr, ii;
    if (k_contig) { rr = idx % ST_BK; ii = idx / ST_BK; } else { ii = idx % ST_BM; rr = idx / ST_BM; }
    const int row = row0 + ii, r = r0 + rr;
    float v = 0.f;
    if (row < nrows && r < r_end) {
      const int64_t o = l.at(row, r);
      if (o >= 0) v = to_f32(base[o]);     // o < 0: structurally zero element (CLS rows of the in-place patch matrix)
    }
    dst[rr][ii] = v;
  }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(ST_THREADS)
gemm_simt_kernel(GemmProblem p, int r_per_split) {
  pdl_sync_traced(TK_GEMM_SIMT);
  __shared__ float As[ST_BK][ST_BM + ST_PAD];
  __shared__ float Bs[ST_BK][ST_BN + ST_PAD];
  const int i0 = blockIdx.y * ST_BM, j0 = blockIdx.x * ST_BN;
  const int r_begin = blockIdx.z * r_per_split;
  const int r_end = min(p.R, r_begin + r_per_split);
  const TI* A = reinterpret_cast<const TI*>(p.A);
  const TI* B = reinterpret_cast<const TI*>(p.B);
  const bool a_kc = (p.la.s_col == 1), b_kc = (p.lb.s_col == 1);
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  for (int r0 = r_begin; r0 < r_end; r0 += ST_BK) {
    simt_load_tile<TI>(A, p.la, i0, p.I, r0, r_end, As, a_kc);
    simt_load_tile<TI>(B, p.lb, j0, p.J, r0, r_end, Bs, b_kc);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < ST_BK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {av.x, av.y, av.z, av.w};
      const float b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    const int i = i0 + ty * 4 + x;
    if (i >= p.I) continue;
#pragma unroll
    for (int y = 0; y < 4; ++y) {
      const int j = j0 + tx * 4 + y;
      if (j < p.J) epilogue_scalar<TO>(p.ep, i, j, acc[x][y]);
    }
  }
  trace_end(TK_GEMM_SIMT);
}

int gemm_simt(const GemmProblem& p, int splits, cudaStream_t st) {
  VITK_CHECK_ARG(p.I > 0 && p.J > 0 && p.R > 0 && p.A && p.B && p.ep.out);
  if (splits < 1) splits = 1;
  if (p.ep.mode != E_ACCUM) splits = 1;
  int r_per_split = (p.R + splits - 1) / splits;
  r_per_split = (r_per_split + ST_BK - 1) / ST_BK * ST_BK;
  splits = (p.R + r_per_split - 1) / r_per_split;
  dim3 grid((p.J + ST_BN - 1) / ST_BN, (p.I + ST_BM - 1) / ST_BM, splits);
  const bool in16 = p.in_dtype == VITK_BF16, out16 = p.ep.out_dtype == VITK_BF16;
  if (!in16 && !out16) VITK_LAUNCH((gemm_simt_kernel<float, float>), grid, ST_THREADS, 0, st, p, r_per_split);
  else if (in16 && out16) VITK_LAUNCH((gemm_simt_kernel<bf16, bf16>), grid, ST_THREADS, 0, st, p, r_per_split);
  else if (in16 && !out16) VITK_LAUNCH((gemm_simt_kernel<bf16, float>), grid, ST_THREADS, 0, st, p, r_per_split);
  else VITK_LAUNCH((gemm_simt_kernel<float, bf16>), grid, ST_THREADS, 0, st, p, r_per_split);
  return VITK_OK;
}

}  // namespace vitk
