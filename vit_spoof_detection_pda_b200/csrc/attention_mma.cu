// Tensor-core attention for bf16 storage (placeholder: routes to the FFMA kernel until the mma kernel lands).
#include "common.cuh"
namespace vitk {
int attn_fwd_simt(const void* qkv, void* out, float* lse, int batch, int dtype, cudaStream_t st);
int attn_bwd_simt(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int batch,
                  int dtype, cudaStream_t st);
int attn_fwd_mma(const void* qkv, void* out, float* lse, int batch, cudaStream_t st) {
  return attn_fwd_simt(qkv, out, lse, batch, VITK_BF16, st);
}
int attn_bwd_mma(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int batch,
                 cudaStream_t st) {
  return attn_bwd_simt(qkv, out, dout, lse, dqkv, batch, VITK_BF16, st);
}
}  // namespace vitk
