// Whole-model drivers: enqueue the ViT-B/16 + PAD head forward, and the backward in stages.
// The sequence restates timm VisionTransformer.forward as called by
// /root/reference/train_advanced.py:202-204 on top of the libvitk kernels.
#include "common.cuh"

namespace vitk {

int attn_fwd_dispatch(const void* qkv, void* out, float* lse, int batch, int dtype, int engine, cudaStream_t st);
int attn_bwd_dispatch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                      float* dqkv_colsum, int batch, int dtype, int engine, cudaStream_t st, int cs_sections);
bool attn_bias_split_supported(int dtype, int engine);
int colsum_headmajor(const void* x, int dtype, int M, int C, float* db, cudaStream_t st);   // linear.cu

// vitk_model.sm_budget applies to the kernels THIS call enqueues (thread-local, restored on return)
struct SmBudgetScope {
  int prev;
  bool on;
  explicit SmBudgetScope(int n) : prev(0), on(n > 0) { if (on) prev = set_sm_budget(n); }
  ~SmBudgetScope() { if (on) set_sm_budget(prev); }
};

constexpr int D = VITK_DIM, NT = VITK_NTOK, MLP = VITK_MLP, HID = VITK_HEAD_HIDDEN;

// Weight-gradient side stream.  The four wgrad GEMMs of a block only feed the optimizer, so they run on a second
// stream next to the dgrad / LayerNorm / attention chain: CTAs of whichever kernel is pending take the SMs that the
// other stream's kernel leaves idle in its ramp, its last partial wave and its tail (persistent grids of 148 CTAs are
// work-conserving this way).  Ordering is by events; the stage joins the side stream before it returns, so callers
// (autograd, the bucketed all-reduce) still see plain stream semantics.
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
};
static SideStream* side_stream() {
  static thread_local SideStream per_dev[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  SideStream& ss = per_dev[dev];
  if (!ss.s) {
    if (cudaStreamCreateWithFlags(&ss.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (auto& e : ss.ev)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  }
  return &ss;
}

struct BlockOffsets { int64_t n1w, n1b, qkvw, qkvb, projw, projb, n2w, n2b, fc1w, fc1b, fc2w, fc2b; };
struct ParamOffsets {
  int64_t cls, pos, pew, peb;
  BlockOffsets blk[64];
  int64_t normw, normb, hlnw, hlnb, hw1, hb1, hw2, hb2;
  int64_t total;
  int n_tensors;
};

static inline int64_t pad64(int64_t n) { return (n + 63) / 64 * 64; }

static int build_param_offsets(int depth, int num_classes, ParamOffsets* po, int64_t* offsets, int64_t* sizes, int maxn) {
  if (depth < 1 || depth > 64 || num_classes < 1) return -1;
  int64_t cur = 0;
  int idx = 0;
  auto add = [&](int64_t n) {
    const int64_t o = cur;
    if (offsets && idx < maxn) { offsets[idx] = o; sizes[idx] = n; }
    ++idx;
    cur += pad64(n);
    return o;
  };
  po->cls = add(D); po->pos = add((int64_t)NT * D); po->pew = add((int64_t)D * D); po->peb = add(D);
  for (int l = 0; l < depth; ++l) {
    BlockOffsets& b = po->blk[l];
    b.n1w = add(D); b.n1b = add(D);
    b.qkvw = add((int64_t)3 * D * D); b.qkvb = add(3 * D);
    b.projw = add((int64_t)D * D); b.projb = add(D);
    b.n2w = add(D); b.n2b = add(D);
    b.fc1w = add((int64_t)MLP * D); b.fc1b = add(MLP);
    b.fc2w = add((int64_t)D * MLP); b.fc2b = add(D);
  }
  po->normw = add(D); po->normb = add(D);
  po->hlnw = add(D); po->hlnb = add(D);
  po->hw1 = add((int64_t)HID * D); po->hb1 = add(HID);
  po->hw2 = add((int64_t)num_classes * HID); po->hb2 = add(num_classes);
  po->total = cur;
  po->n_tensors = idx;
  return 0;
}

// ---- workspace plan -------------------------------------------------------------------------
struct Plan {
  size_t total;
  // forward (per-layer stride applies only when activations are saved)
  size_t nchw, x_in, x_mid, ln1, ln2, qkv, ao, u, g, mean1, rstd1, mean2, rstd2, lse;
  size_t x_stride, act768_stride, qkv_stride, act3072_stride, stat_stride, lse_stride;
  size_t meanf, rstdf, feat, head_save;
  // backward transients
  size_t dx, dx16, du, dh, dqkv, dfeat, dxc;
  // scratch of the GEMMs' split tail (training, bf16; gemm_tc.cu tc_tail_plan): TAIL_FLOATS floats, kept all zero
  size_t tail;
};
constexpr size_t TAIL_FLOATS = (size_t)512 * D + 1024;   // = vitk_gemm_tail_scratch_floats(768): tickets + 512 rows

static void make_plan(int B, int depth, int precision, int training, int frozen, Plan* p) {
  const size_t T = precision == VITK_PREC_BF16 ? 2 : 4;
  const size_t M = (size_t)B * NT;
  const bool save = training && !frozen;
  size_t cur = 0;
  auto take = [&](size_t bytes) { const size_t o = cur; cur += align_up(bytes, 256); return o; };
  const size_t L = save ? (size_t)depth : 1;
  p->nchw = take((size_t)B * 3 * VITK_IMG * VITK_IMG * 4);   // fp32 NCHW image of a uint8-fed step (fp32-validate forward, weight gradient)
  p->x_stride = align_up(M * D * 4, 256);
  p->x_in = take(p->x_stride * (save ? depth + 1 : 2));   // eval: ping-pong
  p->x_mid = take(p->x_stride * L);
  p->act768_stride = align_up(M * D * T, 256);
  p->ln1 = take(p->act768_stride * L);
  p->ln2 = take(p->act768_stride * L);
  p->ao = take(p->act768_stride * L);
  p->qkv_stride = align_up(M * 3 * D * T, 256);
  p->qkv = take(p->qkv_stride * L);
  p->act3072_stride = align_up(M * MLP * T, 256);
  p->u = take(p->act3072_stride * L);
  p->g = take(p->act3072_stride * L);
  p->stat_stride = align_up(M * 4, 256);
  p->mean1 = take(p->stat_stride * L); p->rstd1 = take(p->stat_stride * L);
  p->mean2 = take(p->stat_stride * L); p->rstd2 = take(p->stat_stride * L);
  p->lse_stride = align_up(M * VITK_HEADS * 4, 256);
  p->lse = take(p->lse_stride * L);
  p->meanf = take((size_t)B * 4); p->rstdf = take((size_t)B * 4);
  p->feat = take((size_t)B * D * 4);
  p->head_save = take(vitk_head_save_floats(B) * 4);
  p->dfeat = take((size_t)B * D * 4);
  p->dxc = take((size_t)B * D * 4);
  if (save) {
    p->dx = take(M * D * 4);
    p->dx16 = take(M * D * T);
    p->du = take(M * MLP * T);
    p->dh = take(M * D * T);
    p->dqkv = take(M * 3 * D * T);
  } else {
    p->dx = p->dx16 = p->du = p->dh = p->dqkv = 0;
  }
  p->tail = (training && precision == VITK_PREC_BF16) ? take(TAIL_FLOATS * 4) : 0;
  p->total = cur;
}

// dx[m][:] = (m % 197 == 0) ? dxc[m / 197][:] : 0 ; act copy alongside
template <typename T>
__global__ void __launch_bounds__(256)
scatter_cls_grad_kernel(const float* __restrict__ dxc, float* __restrict__ dx, T* __restrict__ dx16, int64_t M,
                        float* __restrict__ colsum) {
  pdl_sync_traced(TK_SCATTER_CLS);
  const int64_t total = M * (D / 4);
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = idx / (D / 4);
    const int c4 = (int)(idx % (D / 4));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m % NT == 0) {
      v = *reinterpret_cast<const float4*>(dxc + (m / NT) * D + c4 * 4);
      if (colsum) {  // bias gradient of the last block's fc2 (its dY is this scattered gradient)
        atomicAdd(colsum + c4 * 4 + 0, v.x); atomicAdd(colsum + c4 * 4 + 1, v.y);
        atomicAdd(colsum + c4 * 4 + 2, v.z); atomicAdd(colsum + c4 * 4 + 3, v.w);
      }
    }
    *reinterpret_cast<float4*>(dx + m * D + c4 * 4) = v;
    if (dx16) {
      uint2 u;
      u.x = pack_bf16x2(v.x, v.y);
      u.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(dx16 + m * D + c4 * 4) = u;
    }
  }
  trace_end(TK_SCATTER_CLS);
}

struct Ctx {
  const vitk_model* m;
  ParamOffsets po;
  Plan pl;
  int dt;       // activation dtype
  size_t T;
  int M;
  bool save;
  char* ws;
  cudaStream_t st;
  float* tail;          // split-tail scratch (nullptr: eval / fp32-validate -- whole-K tiles only)
  size_t tail_floats;
  const void* W(int64_t off) const {
    return dt == VITK_BF16 ? (const void*)((const bf16*)m->params16 + off) : (const void*)(m->params + off);
  }
  const float* P(int64_t off) const { return m->params + off; }
  float* G(int64_t off) const { return m->grads + off; }
  char* at(size_t base, size_t stride, int l) const { return ws + base + (save ? stride * (size_t)l : 0); }
};

static int make_ctx(const vitk_model* m, void* stream, Ctx* c) {
  VITK_CHECK_ARG(m && m->batch > 0 && m->depth >= 1 && m->depth <= 64 && m->num_classes >= 1 && m->num_classes <= 8);
  VITK_CHECK_ARG(m->precision == VITK_PREC_FP32_VALIDATE || m->precision == VITK_PREC_BF16);
  VITK_CHECK_ARG(m->params && m->workspace && ((uintptr_t)m->workspace % 256) == 0);
  VITK_CHECK_ARG(m->precision != VITK_PREC_BF16 || m->params16);
  c->m = m;
  build_param_offsets(m->depth, m->num_classes, &c->po, nullptr, nullptr, 0);
  make_plan(m->batch, m->depth, m->precision, m->training, m->frozen_backbone, &c->pl);
  c->dt = m->precision == VITK_PREC_BF16 ? VITK_BF16 : VITK_F32;
  c->T = c->dt == VITK_BF16 ? 2 : 4;
  c->M = m->batch * NT;
  c->save = m->training && !m->frozen_backbone;
  c->ws = (char*)m->workspace;
  c->st = (cudaStream_t)stream;
  const bool split = m->training && m->precision == VITK_PREC_BF16;
  c->tail = split ? (float*)(c->ws + c->pl.tail) : nullptr;
  c->tail_floats = split ? TAIL_FLOATS : 0;
  return VITK_OK;
}

}  // namespace vitk

using namespace vitk;

extern "C" int64_t vitk_param_layout(int depth, int num_classes, int64_t* offsets, int64_t* sizes, int max_tensors) {
  ParamOffsets po;
  if (build_param_offsets(depth, num_classes, &po, offsets, sizes, max_tensors) != 0) {
    set_error("vitk_param_layout: bad depth/num_classes");
    return -1;
  }
  return po.total;
}

extern "C" size_t vitk_workspace_bytes(int batch, int depth, int precision, int training) {
  Plan p;
  // `training` bit 1 (value 2) marks the frozen-backbone plan (no encoder activations saved)
  make_plan(batch, depth, precision, training & 1, (training >> 1) & 1, &p);
  return p.total;
}

extern "C" int vitk_model_num_bwd_stages(int depth) { return depth + 2; }

extern "C" int vitk_model_fwd(const vitk_model* m, void* stream) {
  Ctx c;
  VITK_TRY(make_ctx(m, stream, &c));
  VITK_CHECK_ARG((m->images || m->images_u8) && m->logits);
  SmBudgetScope budget(m->sm_budget);
  const ParamOffsets& po = c.po;
  const Plan& pl = c.pl;
  const int M = c.M, dt = c.dt, eng = m->engine;
  void* st = stream;
  char* ws = c.ws;
  auto xin = [&](int l) { return (float*)(ws + pl.x_in + pl.x_stride * (size_t)(c.save ? l : (l & 1))); };
  auto w16 = [&](int64_t off) -> const void* { return dt == VITK_BF16 ? (const void*)((const bf16*)m->params16 + off) : nullptr; };

  // the split GEMMs hand the scratch back zeroed; clearing it once per call makes that independent of what the caller's
  // workspace held before (first use, a different plan in the same allocation)
  if (c.tail) VITK_CUDA(cudaMemsetAsync(c.tail, 0, c.tail_floats * 4, c.st));
  if (m->images_u8)   // uint8 HWC pixels: ToTensor + Normalize fused into the patch loader
    VITK_TRY(vitk_patch_embed_fwd_u8(m->images_u8, m->norm_mean, m->norm_std, c.P(po.pew), w16(po.pew), c.P(po.peb), c.P(po.cls),
                                     c.P(po.pos), xin(0), (float*)(ws + pl.nchw), m->batch, m->precision, eng, st));
  else
    VITK_TRY(vitk_patch_embed_fwd(m->images, c.P(po.pew), w16(po.pew), c.P(po.peb), c.P(po.cls), c.P(po.pos), xin(0), m->batch,
                                  m->precision, eng, st));
  for (int l = 0; l < m->depth; ++l) {
    const BlockOffsets& b = po.blk[l];
    float* x = xin(l);
    float* xmid = (float*)c.at(pl.x_mid, pl.x_stride, l);
    float* xout = xin(l + 1);
    void* ln1 = c.at(pl.ln1, pl.act768_stride, l);
    void* ln2 = c.at(pl.ln2, pl.act768_stride, l);
    void* qkv = c.at(pl.qkv, pl.qkv_stride, l);
    void* ao = c.at(pl.ao, pl.act768_stride, l);
    void* u = c.save ? (void*)c.at(pl.u, pl.act3072_stride, l) : nullptr;
    void* g = c.at(pl.g, pl.act3072_stride, l);
    float* mean1 = c.save ? (float*)c.at(pl.mean1, pl.stat_stride, l) : nullptr;
    float* rstd1 = c.save ? (float*)c.at(pl.rstd1, pl.stat_stride, l) : nullptr;
    float* mean2 = c.save ? (float*)c.at(pl.mean2, pl.stat_stride, l) : nullptr;
    float* rstd2 = c.save ? (float*)c.at(pl.rstd2, pl.stat_stride, l) : nullptr;
    float* lse = c.save ? (float*)c.at(pl.lse, pl.lse_stride, l) : nullptr;

    VITK_TRY(vitk_layernorm_fwd(x, D, c.P(b.n1w), c.P(b.n1b), ln1, dt, mean1, rstd1, M, 1e-6f, st));
    VITK_TRY(vitk_linear_fwd_ws(ln1, VITK_LAYOUT_ROWMAJOR, c.W(b.qkvw), c.P(b.qkvb), qkv, nullptr, M, 3 * D, D,
                                VITK_EPI_QKV_SCATTER, dt, eng, c.tail, c.tail_floats, st));
    VITK_TRY(attn_fwd_dispatch(qkv, ao, lse, m->batch, dt, eng, c.st));
    VITK_TRY(vitk_linear_fwd(ao, VITK_LAYOUT_ROWMAJOR, c.W(b.projw), c.P(b.projb), xmid, x, M, D, D,
                             VITK_EPI_BIAS_RESIDUAL, dt, eng, st));
    VITK_TRY(vitk_layernorm_fwd(xmid, D, c.P(b.n2w), c.P(b.n2b), ln2, dt, mean2, rstd2, M, 1e-6f, st));
    VITK_TRY(vitk_linear_fwd_ws(ln2, VITK_LAYOUT_ROWMAJOR, c.W(b.fc1w), c.P(b.fc1b), g, u, M, MLP, D,
                                VITK_EPI_BIAS_GELU, dt, eng, c.tail, c.tail_floats, st));
    VITK_TRY(vitk_linear_fwd_ws(g, VITK_LAYOUT_ROWMAJOR, c.W(b.fc2w), c.P(b.fc2b), xout, xmid, M, D, MLP,
                                VITK_EPI_BIAS_RESIDUAL, dt, eng, c.tail, c.tail_floats, st));
  }
  // final norm on the CLS rows only (global_pool='token': only x[:,0] is consumed), then the head
  float* xl = xin(m->depth);
  float* feat = (float*)(ws + pl.feat);
  VITK_TRY(vitk_layernorm_fwd(xl, (int64_t)NT * D, c.P(po.normw), c.P(po.normb), feat, VITK_F32,
                              m->training ? (float*)(ws + pl.meanf) : nullptr,
                              m->training ? (float*)(ws + pl.rstdf) : nullptr, m->batch, 1e-6f, st));
  VITK_TRY(vitk_head_fwd(feat, c.P(po.hlnw), c.P(po.hlnb), c.P(po.hw1), c.P(po.hb1), c.P(po.hw2), c.P(po.hb2), m->mask1,
                         m->mask2, m->logits, m->training ? (float*)(ws + pl.head_save) : nullptr, m->batch,
                         m->num_classes, st));
  return VITK_OK;
}

extern "C" int vitk_model_bwd_stage(const vitk_model* m, int stage, void* stream) {
  Ctx c;
  VITK_TRY(make_ctx(m, stream, &c));
  VITK_CHECK_ARG(m->training && m->grads && m->dlogits);
  VITK_CHECK_ARG(stage >= 0 && stage < m->depth + 2);
  SmBudgetScope budget(m->sm_budget);
  const ParamOffsets& po = c.po;
  const Plan& pl = c.pl;
  const int M = c.M, dt = c.dt, eng = m->engine;
  void* st = stream;
  char* ws = c.ws;
  float* dx = (float*)(ws + pl.dx);
  void* dxa = dt == VITK_BF16 ? (void*)(ws + pl.dx16) : (void*)dx;   // dx in the activation dtype
  void* dx16 = dt == VITK_BF16 ? (void*)(ws + pl.dx16) : nullptr;

  if (stage == 0) {
    if (c.tail) VITK_CUDA(cudaMemsetAsync(c.tail, 0, c.tail_floats * 4, c.st));
    float* dfeat = (float*)(ws + pl.dfeat);
    VITK_TRY(vitk_head_bwd(m->dlogits, (float*)(ws + pl.head_save), c.P(po.hlnw), c.P(po.hw1), c.P(po.hw2), m->mask1,
                           m->mask2, dfeat, c.G(po.hlnw), c.G(po.hlnb), c.G(po.hw1), c.G(po.hb1), c.G(po.hw2),
                           c.G(po.hb2), m->batch, m->num_classes, st));
    if (m->frozen_backbone) return VITK_OK;
    float* xl = (float*)(ws + pl.x_in + pl.x_stride * (size_t)m->depth);
    float* dxc = (float*)(ws + pl.dxc);
    VITK_TRY(vitk_layernorm_bwd(dfeat, VITK_F32, xl, (int64_t)NT * D, c.P(po.normw), (float*)(ws + pl.meanf),
                                (float*)(ws + pl.rstdf), nullptr, dxc, nullptr, c.G(po.normw), c.G(po.normb), nullptr,
                                m->batch, st));
    const int64_t total = (int64_t)M * (D / 4);
    const int grid = (int)((total + 255) / 256 < (int64_t)sm_count() * 8 ? (total + 255) / 256 : (int64_t)sm_count() * 8);
    VITK_LAUNCH((scatter_cls_grad_kernel<bf16>), grid, 256, 0, c.st, dxc, dx, (bf16*)dx16, M, c.G(po.blk[m->depth - 1].fc2b));
    return VITK_OK;
  }
  if (m->frozen_backbone) return VITK_OK;
  if (stage == m->depth + 1) {
    const float* img = m->images;
    if (m->images_u8) {     // uint8-fed step: the weight gradient reads the normalised fp32 image
      VITK_TRY(vitk_u8_to_nchw(m->images_u8, m->norm_mean, m->norm_std, (float*)(ws + pl.nchw), m->batch, st));
      img = (const float*)(ws + pl.nchw);
    }
    VITK_CHECK_ARG(img);
    return vitk_patch_embed_wgrad(dx, dx16, img, c.G(po.pew), c.G(po.peb), c.G(po.cls), c.G(po.pos), m->batch, m->precision, eng, st);
  }
  const int l = m->depth - stage;  // stage 1 -> last block
  const BlockOffsets& b = po.blk[l];
  float* x = (float*)(ws + pl.x_in + pl.x_stride * (size_t)l);
  float* xmid = (float*)c.at(pl.x_mid, pl.x_stride, l);
  void* ln1 = c.at(pl.ln1, pl.act768_stride, l);
  void* ln2 = c.at(pl.ln2, pl.act768_stride, l);
  void* qkv = c.at(pl.qkv, pl.qkv_stride, l);
  void* ao = c.at(pl.ao, pl.act768_stride, l);
  void* u = c.at(pl.u, pl.act3072_stride, l);
  void* g = c.at(pl.g, pl.act3072_stride, l);
  void* du = ws + pl.du;
  void* dh = ws + pl.dh;
  void* dqkv = ws + pl.dqkv;

  SideStream* ss = (dt == VITK_BF16 && !(m->flags & VITK_FLAG_WGRAD_INLINE)) ? side_stream() : nullptr;
  cudaStream_t ms = c.st;
  void* wst = ss ? (void*)ss->s : st;     // stream of the weight-gradient GEMMs
  auto after = [&](int e, cudaStream_t from, cudaStream_t to) -> int {   // `to` continues after everything enqueued on `from`
    if (!ss) return VITK_OK;
    VITK_CUDA(cudaEventRecord(ss->ev[e], from));
    VITK_CUDA(cudaStreamWaitEvent(to, ss->ev[e], 0));
    return VITK_OK;
  };
  // MLP:  x_out = x_mid + fc2(gelu(fc1(ln2(x_mid))))
  // (fc2 / proj bias gradients = column sums of the residual-stream gradient: produced by the kernel that wrote
  //  it -- the LayerNorm backward below, or the CLS scatter for the last block)
  VITK_TRY(after(0, ms, ss ? ss->s : ms));   // dx / dx16 of the previous stage are final
  VITK_TRY(vitk_linear_wgrad(dxa, VITK_LAYOUT_ROWMAJOR, g, c.G(b.fc2w), nullptr, M, D, MLP, dt, eng, wst));
  if (ss) VITK_CUDA(cudaEventRecord(ss->ev[1], ss->s));                  // fc2 wgrad has read dx16
  // fc1's bias gradient = column sums of du: fused into the epilogue of the GEMM that produces du
  VITK_TRY(vitk_linear_dgrad(dxa, VITK_LAYOUT_ROWMAJOR, c.W(b.fc2w), du, u, c.G(b.fc1b), M, D, MLP, dt, eng, st));
  VITK_TRY(after(2, ms, ss ? ss->s : ms));   // du ready
  VITK_TRY(vitk_linear_wgrad(du, VITK_LAYOUT_ROWMAJOR, ln2, c.G(b.fc1w), nullptr, M, MLP, D, dt, eng, wst));
  VITK_TRY(vitk_linear_dgrad_ws(du, VITK_LAYOUT_ROWMAJOR, c.W(b.fc1w), dh, nullptr, nullptr, M, MLP, D, dt, eng, c.tail,
                                c.tail_floats, st));
  if (ss) VITK_CUDA(cudaStreamWaitEvent(ms, ss->ev[1], 0));              // the LayerNorm backward overwrites dx16
  VITK_TRY(vitk_layernorm_bwd(dh, dt, xmid, D, c.P(b.n2w), (float*)c.at(pl.mean2, pl.stat_stride, l),
                              (float*)c.at(pl.rstd2, pl.stat_stride, l), dx, dx, dx16, c.G(b.n2w), c.G(b.n2b), c.G(b.projb), M, st));
  // attention:  x_mid = x + proj(attn(qkv(ln1(x))))
  VITK_TRY(after(3, ms, ss ? ss->s : ms));   // dx16 (gradient of x_mid) ready
  VITK_TRY(vitk_linear_wgrad(dxa, VITK_LAYOUT_ROWMAJOR, ao, c.G(b.projw), nullptr, M, D, D, dt, eng, wst));
  if (ss) VITK_CUDA(cudaEventRecord(ss->ev[4], ss->s));                  // proj wgrad has read dx16
  // the qkv bias gradient = column sums of dqkv.  bf16 path (attention.cu, attn_bias_split_supported): the v section is the
  // column sum of dh (softmax rows sum to one), taken in the epilogue of the GEMM that produces dh; the k section is zero
  // (rows of dS sum to zero); only the q section is summed by the attention backward's epilogue warps.  fp32 path: a
  // column-sum pass over dqkv.
  const bool split_bias = attn_bias_split_supported(dt, eng);
  VITK_TRY(vitk_linear_dgrad(dxa, VITK_LAYOUT_ROWMAJOR, c.W(b.projw), dh, nullptr, split_bias ? c.G(b.qkvb) + 2 * D : nullptr,
                             M, D, D, dt, eng, st));
  // (split_bias: the attention kernel sums nothing -- its epilogue warps gate the recycling of the dV / dK / dQ accumulators,
  //  and the red.global adds of 148 CTAs into the same 64 floats of a head serialise in L2: measured 5.7 us per item on the
  //  kernel's critical path.  The q section is a coalesced column-sum pass over dQ on the weight-gradient stream instead.)
  VITK_TRY(attn_bwd_dispatch(qkv, ao, dh, (float*)c.at(pl.lse, pl.lse_stride, l), dqkv, split_bias ? nullptr : c.G(b.qkvb), m->batch,
                             dt, eng, c.st, split_bias ? 0 : 7));
  VITK_TRY(after(5, ms, ss ? ss->s : ms));   // dqkv ready
  if (split_bias) VITK_TRY(colsum_headmajor(dqkv, dt, M, D, c.G(b.qkvb), (cudaStream_t)wst));   // q section: heads 0..11 of dqkv
  VITK_TRY(vitk_linear_wgrad(dqkv, VITK_LAYOUT_HEADMAJOR, ln1, c.G(b.qkvw), nullptr, M, 3 * D, D, dt, eng, wst));
  VITK_TRY(vitk_linear_dgrad_ws(dqkv, VITK_LAYOUT_HEADMAJOR, c.W(b.qkvw), dh, nullptr, nullptr, M, 3 * D, D, dt, eng, c.tail,
                                c.tail_floats, st));
  if (ss) VITK_CUDA(cudaStreamWaitEvent(ms, ss->ev[4], 0));              // the LayerNorm backward overwrites dx16
  VITK_TRY(vitk_layernorm_bwd(dh, dt, x, D, c.P(b.n1w), (float*)c.at(pl.mean1, pl.stat_stride, l),
                              (float*)c.at(pl.rstd1, pl.stat_stride, l), dx, dx, dx16, c.G(b.n1w), c.G(b.n1b),
                              l > 0 ? c.G(po.blk[l - 1].fc2b) : nullptr, M, st));
  // join: du / dqkv / dx16 are rewritten by the next stage, and the caller may reduce this stage's gradients
  if (ss) VITK_TRY(after(0, ss->s, ms));
  return VITK_OK;
}
