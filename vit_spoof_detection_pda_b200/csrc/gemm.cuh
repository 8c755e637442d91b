// Internal GEMM problem description shared by the SIMT (fp32-validate) and tcgen05 engines.
//   C[i][j] = sum_r A(i,r) * B(j,r)      i < I, j < J, r < R   (both operands indexed (row, reduction))
// followed by one fused epilogue.  Operand storage is described by MatLayout so that the same
// kernels serve forward (X W^T), dgrad (dY W) and wgrad (dY^T X) of nn.Linear and the
// head-major [C/64][M][64] storage of q/k/v.
#pragma once
#include "common.cuh"

namespace vitk {

struct MatLayout {
  int64_t s_row;  // element stride of the row index (i or j)
  int64_t s_col;  // element stride of the reduction index r
  int32_t split;  // 0: none; 1: row index is split (row>>6)*s_blk + (row&63)*s_row; 2: same for r;
                  // 3 / 4: the patch matrix of an NCHW image read in place (no im2col copy): 3 = (row = token b*197+t,
                  // col = k = c*256+i*16+j), 4 = (row = k, col = token); CLS tokens (t = 0) have no patch: at() < 0 = zero
  int64_t s_blk;
  __host__ __device__ __forceinline__ int64_t at(int row, int col) const {
    if (split >= 3) {
      const int tok = split == 3 ? row : col, k = split == 3 ? col : row;
      const int b = tok / VITK_NTOK, t = tok % VITK_NTOK;
      if (t == 0) return -1;
      const int py = (t - 1) / 14, px = (t - 1) % 14, c = k >> 8, i = (k >> 4) & 15, j = k & 15;
      return (((int64_t)b * 3 + c) * VITK_IMG + py * 16 + i) * VITK_IMG + px * 16 + j;
    }
    int64_t a = (split == 1) ? (int64_t)(row >> 6) * s_blk + (int64_t)(row & 63) * s_row : (int64_t)row * s_row;
    a += (split == 2) ? (int64_t)(col >> 6) * s_blk + (int64_t)(col & 63) * s_col : (int64_t)col * s_col;
    return a;
  }
};

static inline MatLayout layout_rowmajor(int64_t ld) { return MatLayout{ld, 1, 0, 0}; }          // (row, r) -> row*ld + r
static inline MatLayout layout_transposed(int64_t ld) { return MatLayout{1, ld, 0, 0}; }        // (row, r) -> r*ld + row
// head-major X_hm(m, c) = (c>>6)*M*64 + m*64 + (c&63)
static inline MatLayout layout_headmajor_rows_m(int64_t M) { return MatLayout{64, 1, 2, M * 64}; }   // row = m, r = c
static inline MatLayout layout_headmajor_rows_c(int64_t M) { return MatLayout{1, 64, 1, M * 64}; }   // row = c, r = m
// the patch matrix read straight from the NCHW fp32 image (SIMT engine only; the tcgen05 path has its own TMA kernel)
static inline MatLayout layout_patches_rows_tok() { return MatLayout{0, 1, 3, 0}; }   // row = token, r = k   (k contiguous)
static inline MatLayout layout_patches_rows_k() { return MatLayout{1, 0, 4, 0}; }     // row = k, r = token

enum EpiMode {
  E_STORE = 0,          // out[i][j] = acc (+ bias[j])                       out_dtype, row-major ldc
  E_BIAS_GELU = 1,      // u = acc + bias[j]; out = gelu(u); aux = gelu'(u)  out_dtype (aux kept for backward)
  E_BIAS_RESIDUAL = 2,  // out = residual[i][j] + acc + bias[j]             fp32
  E_QKV_SCATTER = 3,    // out_hm(i, j) = acc + bias[j]                      out_dtype, head-major
  E_GELU_BWD = 4,       // out = acc * aux[i][j]   (aux = gelu'(u) saved by forward) out_dtype
  E_ACCUM = 5,          // atomicAdd(out_f32[i][j], acc)
  E_PATCH = 6           // rows are tokens (b*197+t): out_f32[i][j] = (t ? acc + bias[j] : cls[j]) + pos[t][j]  (aux = cls)
};

struct EpiParams {
  int32_t mode;
  int32_t out_dtype;      // VITK_F32 / VITK_BF16
  void* out;
  void* aux;              // gelu'(u) out (E_BIAS_GELU) / in (E_GELU_BWD), same dtype+layout as out; cls (E_PATCH)
  const float* bias;      // [J] or nullptr
  const float* residual;  // fp32 [I][ldc] (E_BIAS_RESIDUAL) / pos_embed (E_PATCH)
  int64_t ldc;            // row stride of out / aux / residual (row-major modes)
  int64_t hm_rows;        // M of the head-major output (E_QKV_SCATTER)
  float* colsum;          // optional (E_GELU_BWD, tcgen05 engine): += column sums of the output (bias gradient of fc1)
};

struct GemmProblem {
  int32_t I, J, R;
  const void* A;
  const void* B;
  MatLayout la, lb;
  int32_t in_dtype;  // VITK_F32 / VITK_BF16
  EpiParams ep;
  // optional scratch (all zero on entry, left all zero) for the split tail of the tcgen05 engine (gemm_tc.cu: tc_tail_plan);
  // nullptr = whole tiles only
  float* tail_scratch;
  int64_t tail_scratch_floats;
};

template <typename T> __device__ __forceinline__ void store_as(void* base, int64_t idx, float v) {
  reinterpret_cast<T*>(base)[idx] = from_f32<T>(v);
}
template <typename T> __device__ __forceinline__ float load_as(const void* base, int64_t idx) {
  return to_f32(reinterpret_cast<const T*>(base)[idx]);
}

// Scalar epilogue (SIMT engine and tile tails of the tcgen05 engine).
template <typename TO>
__device__ __forceinline__ void epilogue_scalar(const EpiParams& ep, int i, int j, float acc) {
  switch (ep.mode) {
    case E_STORE: {
      if (ep.bias) acc += ep.bias[j];
      store_as<TO>(ep.out, (int64_t)i * ep.ldc + j, acc);
    } break;
    case E_BIAS_GELU: {
      // the reference applies GELU to the 16-bit fc1 output under autocast: round first
      const float u = to_f32(from_f32<TO>(acc + ep.bias[j]));
      if (ep.aux) store_as<TO>(ep.aux, (int64_t)i * ep.ldc + j, gelu_erf_grad(u));
      store_as<TO>(ep.out, (int64_t)i * ep.ldc + j, gelu_erf(u));
    } break;
    case E_BIAS_RESIDUAL: {
      const int64_t o = (int64_t)i * ep.ldc + j;
      reinterpret_cast<float*>(ep.out)[o] = ep.residual[o] + (acc + ep.bias[j]);
    } break;
    case E_QKV_SCATTER: {
      const int64_t o = (int64_t)(j >> 6) * ep.hm_rows * 64 + (int64_t)i * 64 + (j & 63);
      store_as<TO>(ep.out, o, acc + ep.bias[j]);
    } break;
    case E_GELU_BWD: {
      const int64_t o = (int64_t)i * ep.ldc + j;
      store_as<TO>(ep.out, o, acc * load_as<TO>(ep.aux, o));
    } break;
    case E_ACCUM: {
      atomicAdd(reinterpret_cast<float*>(ep.out) + (int64_t)i * ep.ldc + j, acc);
    } break;
    case E_PATCH: {
      const int t = i % VITK_NTOK;
      const float pe = ep.residual[(int64_t)t * ep.ldc + j];
      const float v = (t == 0) ? reinterpret_cast<const float*>(ep.aux)[j] : acc + ep.bias[j];
      reinterpret_cast<float*>(ep.out)[(int64_t)i * ep.ldc + j] = v + pe;
    } break;
    default: break;
  }
}

int gemm_simt(const GemmProblem& p, int splits, cudaStream_t st);
// tcgen05/TMEM engine (bf16 operands only). Returns VITK_ERR_UNSUPPORTED for shapes it does not take.
int gemm_tc(const GemmProblem& p, cudaStream_t st);

}  // namespace vitk
