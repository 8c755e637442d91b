// C-ABI of the attention core + engine dispatch (fp32 -> FFMA kernel, bf16 -> tcgen05 kernel).
#include "common.cuh"

namespace vitk {
int attn_fwd_simt(const void* qkv, void* out, float* lse, int batch, int dtype, cudaStream_t st);
int attn_bwd_simt(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int batch,
                  int dtype, cudaStream_t st);
int attn_fwd_tc(const void* qkv, void* out, float* lse, int batch, cudaStream_t st);
int attn_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, float* dqkv_colsum, int batch,
                cudaStream_t st, int cs_sections);
int tune_knob(int key);
int colsum_headmajor(const void* x, int dtype, int M, int C, float* db, cudaStream_t st);

static bool use_tc(int dtype, int engine) { return dtype == VITK_BF16 && engine != VITK_ENGINE_SIMT; }

int attn_fwd_dispatch(const void* qkv, void* out, float* lse, int batch, int dtype, int engine, cudaStream_t st) {
  if (use_tc(dtype, engine)) return attn_fwd_tc(qkv, out, lse, batch, st);
  return attn_fwd_simt(qkv, out, lse, batch, dtype, st);
}
// dqkv_colsum (optional): fp32 [2304] += column sums of dqkv = the qkv bias gradient; fused into the tensor-core
// kernel, a separate reduction after the FFMA kernel.
//
// The model driver's bf16 path asks the attention kernel for the q section only (cs_sections = 1) when
// attn_bias_split_supported(): with attention dropout 0 (timm's attn_drop, the reference never sets it) every softmax row
// sums to one and every row of dS to zero, so
//     sum_keys dV[key, :] = sum_q (sum_key P[q, key]) dO[q, :] = sum_q dO[q, :]     -> the v section is the column sum of the
//                                                                                    proj dgrad output (fused in its epilogue)
//     sum_keys dK[key, :] = sum_q (sum_key dS[q, key]) Q[q, :] * scale = 0           -> the k section stays zero
// (softmax is invariant to a shift of all keys: the k bias has no gradient).  The attention kernel's epilogue warps,
// which gate the recycling of its dV / dK / dQ accumulators, then sum 4 instead of 12 slabs per item.
bool attn_bias_split_supported(int dtype, int engine) {
  return use_tc(dtype, engine) && tune_knob(12) != 1;   // knob 12 (development build only): all three sections in the kernel
}
int attn_bwd_dispatch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                      float* dqkv_colsum, int batch, int dtype, int engine, cudaStream_t st, int cs_sections) {
  if (use_tc(dtype, engine)) return attn_bwd_tc(qkv, out, dout, lse, dqkv, dqkv_colsum, batch, st, cs_sections);
  VITK_TRY(attn_bwd_simt(qkv, out, dout, lse, dqkv, batch, dtype, st));
  if (dqkv_colsum) VITK_TRY(colsum_headmajor(dqkv, dtype, batch * VITK_NTOK, 3 * VITK_DIM, dqkv_colsum, st));
  return VITK_OK;
}
}  // namespace vitk

using namespace vitk;

extern "C" int vitk_attn_fwd(const void* qkv, void* out, float* lse, int batch, int dtype, void* stream) {
  VITK_CHECK_ARG(qkv && out && batch > 0 && (dtype == VITK_F32 || dtype == VITK_BF16));
  return attn_fwd_dispatch(qkv, out, lse, batch, dtype, VITK_ENGINE_AUTO, (cudaStream_t)stream);
}
extern "C" int vitk_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                             float* dqkv_colsum, int batch, int dtype, void* stream) {
  VITK_CHECK_ARG(qkv && out && dout && lse && dqkv && batch > 0 && (dtype == VITK_F32 || dtype == VITK_BF16));
  return attn_bwd_dispatch(qkv, out, dout, lse, dqkv, dqkv_colsum, batch, dtype, VITK_ENGINE_AUTO, (cudaStream_t)stream, 7);
}
