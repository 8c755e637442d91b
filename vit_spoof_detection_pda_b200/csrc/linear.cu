// nn.Linear forward / dgrad / wgrad: builds GemmProblems and dispatches them to the tcgen05 engine (bf16) or
// the SIMT engine (fp32-validate / cross-check).  (The patch-embedding GEMM has its own TMA-fed kernel: patch_embed.cu.)
// Reference call sites: timm Attention.qkv / .proj, Mlp.fc1 / .fc2 reached from
// /root/reference/train_advanced.py:203 (self.vit(x)); backward from loss.backward() at :330.
#include <atomic>
#include <mutex>
#include <vector>

#include "gemm.cuh"

namespace vitk {

// ---- optional per-launch GEMM timing (bench.py's live roofline measurement) ----------------------
// When enabled, every GEMM launch is bracketed by CUDA events on the launching stream; the list of
// (I, J, R, epilogue, engine, ms) is read back after a synchronize.  Off by default (no events recorded).
struct ProfRec { cudaEvent_t e0, e1; int I, J, R, mode, engine; };
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;

static int run_gemm_impl(const GemmProblem& p, int engine, int simt_splits, cudaStream_t st) {
  if (engine == VITK_ENGINE_TCGEN05) return gemm_tc(p, st);
  return gemm_simt(p, simt_splits, st);
}

static int run_gemm(const GemmProblem& p, int engine, int simt_splits, cudaStream_t st) {
  if (p.in_dtype == VITK_F32) engine = VITK_ENGINE_SIMT;  // fp32 products only exist on the FFMA pipe
  if (engine == VITK_ENGINE_AUTO) engine = VITK_ENGINE_TCGEN05;
  if (!g_prof_on.load()) return run_gemm_impl(p, engine, simt_splits, st);
  ProfRec r{};
  r.I = p.I; r.J = p.J; r.R = p.R; r.mode = p.ep.mode; r.engine = engine;
  VITK_CUDA(cudaEventCreate(&r.e0));
  VITK_CUDA(cudaEventCreate(&r.e1));
  VITK_CUDA(cudaEventRecord(r.e0, st));
  const int rc = run_gemm_impl(p, engine, simt_splits, st);
  VITK_CUDA(cudaEventRecord(r.e1, st));
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
  return rc;
}

// db[c] += sum_m dy(m, c)      (bias gradient; dy row-major [M][C] or head-major)
template <typename T>
__global__ void __launch_bounds__(256)
colsum_kernel(const T* __restrict__ dy, MatLayout l, int M, int C, int rows_per_block, float* __restrict__ db) {
  pdl_sync_traced(TK_COLSUM);
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float s = 0.f;
  if (c < C)
    for (int m = m0 + ty; m < m1; m += 8) s += to_f32(dy[l.at(m, c)]);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < C) {
#pragma unroll
    for (int w = 1; w < 8; ++w) s += red[w][tx];
    atomicAdd(db + c, s);
  }
  trace_end(TK_COLSUM);
}

// bf16 fast path: every thread owns 8 consecutive columns (one 16-byte load per row), CHUNKS threads cover a
// row slab of 8*CHUNKS columns, 256/CHUNKS rows are in flight per iteration -> fully coalesced, HBM-bound.
// Matrix is [M][ld] with the slab starting at column blockIdx.x*8*CHUNKS; blockIdx.z selects a [M][64] head
// block (head-major storage, block stride M*64).
template <int CHUNKS>
__global__ void __launch_bounds__(256)
colsum_bf16_kernel(const bf16* __restrict__ dy, int64_t ld, int64_t blk_stride, int M, int rows_per_block,
                   float* __restrict__ db, int db_blk_stride) {
  pdl_sync_traced(TK_COLSUM);
  constexpr int RL = 256 / CHUNKS;
  __shared__ float red[RL][CHUNKS * 8 + 1];
  const int ch = threadIdx.x % CHUNKS, rl = threadIdx.x / CHUNKS;
  const bf16* base = dy + (int64_t)blockIdx.z * blk_stride + (int64_t)blockIdx.x * (8 * CHUNKS) + ch * 8;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int m = m0 + rl; m < m1; m += RL) {
    const uint4 u = *reinterpret_cast<const uint4*>(base + (int64_t)m * ld);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y; acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][ch * 8 + e] = acc[e];
  __syncthreads();
  for (int c = threadIdx.x; c < CHUNKS * 8; c += 256) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < RL; ++r) s += red[r][c];
    atomicAdd(db + blockIdx.z * db_blk_stride + blockIdx.x * (8 * CHUNKS) + c, s);
  }
  trace_end(TK_COLSUM);
}

static int colsum(const void* dy, int dtype, MatLayout l, int M, int C, float* db, cudaStream_t st) {
  if (dtype == VITK_BF16 && C % 64 == 0) {
    const int sms = sm_count();
    if (l.split == 2) {  // head-major: C/64 blocks of [M][64]
      const int nblk = C / 64;
      int rows_per_block = (M + (4 * sms / nblk)) / (4 * sms / nblk + 1);
      if (rows_per_block < 64) rows_per_block = 64;
      dim3 grid(1, (M + rows_per_block - 1) / rows_per_block, nblk);
      VITK_LAUNCH((colsum_bf16_kernel<8>), grid, 256, 0, st, (const bf16*)dy, 64, l.s_blk, M, rows_per_block, db, 64);
      return VITK_OK;
    }
    if (l.split == 0 && l.s_col == 1 && C % 256 == 0) {
      const int slabs = C / 256;
      int rows_per_block = (M + (4 * sms / slabs)) / (4 * sms / slabs + 1);
      if (rows_per_block < 64) rows_per_block = 64;
      dim3 grid(slabs, (M + rows_per_block - 1) / rows_per_block, 1);
      VITK_LAUNCH((colsum_bf16_kernel<32>), grid, 256, 0, st, (const bf16*)dy, l.s_row, 0, M, rows_per_block, db, 0);
      return VITK_OK;
    }
  }
  const int rows_per_block = 512;
  dim3 grid((C + 31) / 32, (M + rows_per_block - 1) / rows_per_block);
  if (dtype == VITK_BF16) VITK_LAUNCH((colsum_kernel<bf16>), grid, 256, 0, st, (const bf16*)dy, l, M, C, rows_per_block, db);
  else VITK_LAUNCH((colsum_kernel<float>), grid, 256, 0, st, (const float*)dy, l, M, C, rows_per_block, db);
  return VITK_OK;
}

// column sums of a head-major [C/64][M][64] matrix (used by the attention backward's FFMA path for the qkv bias grad)
int colsum_headmajor(const void* x, int dtype, int M, int C, float* db, cudaStream_t st) {
  return colsum(x, dtype, layout_headmajor_rows_m(M), M, C, db, st);
}

}  // namespace vitk

using namespace vitk;

extern "C" int vitk_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
  return g_prof_on.exchange(on);
}

// Synchronises the recorded events and copies up to `max` records: info[5*i..] = I, J, R, epilogue mode, engine.
extern "C" int vitk_prof_read(float* ms, int* info, int max) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int n = 0;
  for (auto& r : g_prof) {
    if (n >= max) break;
    if (cudaEventSynchronize(r.e1) != cudaSuccess) break;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) break;
    ms[n] = t;
    info[5 * n + 0] = r.I; info[5 * n + 1] = r.J; info[5 * n + 2] = r.R; info[5 * n + 3] = r.mode; info[5 * n + 4] = r.engine;
    ++n;
  }
  return n;
}

extern "C" int vitk_linear_fwd(const void* x, int x_layout, const void* w, const float* bias, void* y, void* aux,
                               int M, int N, int K, int epilogue, int dtype, int engine, void* stream) {
  return vitk_linear_fwd_ws(x, x_layout, w, bias, y, aux, M, N, K, epilogue, dtype, engine, nullptr, 0, stream);
}

extern "C" int vitk_linear_fwd_ws(const void* x, int x_layout, const void* w, const float* bias, void* y, void* aux,
                                  int M, int N, int K, int epilogue, int dtype, int engine, float* scratch,
                                  size_t scratch_floats, void* stream) {
  VITK_CHECK_ARG(x && w && bias && y && M > 0 && N > 0 && K > 0);
  VITK_CHECK_ARG(dtype == VITK_F32 || dtype == VITK_BF16);
  GemmProblem p{};
  p.I = M; p.J = N; p.R = K;
  p.A = x; p.B = w; p.in_dtype = dtype;
  p.la = x_layout == VITK_LAYOUT_HEADMAJOR ? layout_headmajor_rows_m(M) : layout_rowmajor(K);
  p.lb = layout_rowmajor(K);
  p.ep.out = y; p.ep.bias = bias; p.ep.ldc = N; p.ep.out_dtype = dtype;
  switch (epilogue) {
    case VITK_EPI_BIAS: p.ep.mode = E_STORE; break;
    case VITK_EPI_BIAS_GELU: p.ep.mode = E_BIAS_GELU; p.ep.aux = aux; break;  // aux == NULL: eval, gelu'(u) not kept
    case VITK_EPI_BIAS_RESIDUAL:
      VITK_CHECK_ARG(aux);
      p.ep.mode = E_BIAS_RESIDUAL; p.ep.residual = (const float*)aux; p.ep.out_dtype = VITK_F32;
      break;
    case VITK_EPI_QKV_SCATTER: VITK_CHECK_ARG(N % 64 == 0); p.ep.mode = E_QKV_SCATTER; p.ep.hm_rows = M; break;
    default: VITK_CHECK_ARG(!"bad epilogue");
  }
  p.tail_scratch = scratch; p.tail_scratch_floats = scratch ? (int64_t)scratch_floats : 0;
  return run_gemm(p, engine, 1, (cudaStream_t)stream);
}

extern "C" int vitk_linear_dgrad(const void* dy, int dy_layout, const void* w, void* dx, const void* gelu_grad,
                                 float* dx_colsum, int M, int N, int K, int dtype, int engine, void* stream) {
  return vitk_linear_dgrad_ws(dy, dy_layout, w, dx, gelu_grad, dx_colsum, M, N, K, dtype, engine, nullptr, 0, stream);
}

extern "C" int vitk_linear_dgrad_ws(const void* dy, int dy_layout, const void* w, void* dx, const void* gelu_grad,
                                    float* dx_colsum, int M, int N, int K, int dtype, int engine, float* scratch,
                                    size_t scratch_floats, void* stream) {
  VITK_CHECK_ARG(dy && w && dx && M > 0 && N > 0 && K > 0);
  VITK_CHECK_ARG(dtype == VITK_F32 || dtype == VITK_BF16);
  VITK_CHECK_ARG(dy_layout != VITK_LAYOUT_HEADMAJOR || N % 64 == 0);
  GemmProblem p{};
  p.I = M; p.J = K; p.R = N;
  p.A = dy; p.B = w; p.in_dtype = dtype;
  p.la = dy_layout == VITK_LAYOUT_HEADMAJOR ? layout_headmajor_rows_m(M) : layout_rowmajor(N);
  p.lb = layout_transposed(K);  // B(j=k, r=n) = w[n*K + k]
  p.ep.out = dx; p.ep.ldc = K; p.ep.out_dtype = dtype;
  p.ep.mode = gelu_grad ? E_GELU_BWD : E_STORE;
  p.ep.aux = const_cast<void*>(gelu_grad);
  // column sums of dX (with gelu_grad: the bias gradient of the Linear in front of the GELU; without: e.g. the v section
  // of the qkv bias gradient from the proj dgrad, attention.cu): fused into the tcgen05 epilogue; the SIMT engine
  // (fp32-validate) runs the reduction kernel on the finished output instead
  int eng = engine;
  if (dtype == VITK_F32) eng = VITK_ENGINE_SIMT;
  if (eng == VITK_ENGINE_AUTO) eng = VITK_ENGINE_TCGEN05;
  p.ep.colsum = (eng == VITK_ENGINE_TCGEN05) ? dx_colsum : nullptr;
  p.tail_scratch = scratch; p.tail_scratch_floats = scratch ? (int64_t)scratch_floats : 0;
  VITK_TRY(run_gemm(p, eng, 1, (cudaStream_t)stream));
  if (dx_colsum && eng != VITK_ENGINE_TCGEN05)
    VITK_TRY(colsum(dx, dtype, layout_rowmajor(K), M, K, dx_colsum, (cudaStream_t)stream));
  return VITK_OK;
}

extern "C" int vitk_linear_wgrad(const void* dy, int dy_layout, const void* x, float* dw, float* db, int M, int N,
                                 int K, int dtype, int engine, void* stream) {
  VITK_CHECK_ARG(dy && x && dw && M > 0 && N > 0 && K > 0);
  VITK_CHECK_ARG(dtype == VITK_F32 || dtype == VITK_BF16);
  VITK_CHECK_ARG(dy_layout != VITK_LAYOUT_HEADMAJOR || N % 64 == 0);
  cudaStream_t st = (cudaStream_t)stream;
  GemmProblem p{};
  p.I = N; p.J = K; p.R = M;
  p.A = dy; p.B = x; p.in_dtype = dtype;
  p.la = dy_layout == VITK_LAYOUT_HEADMAJOR ? layout_headmajor_rows_c(M) : layout_transposed(N);
  p.lb = layout_transposed(K);
  p.ep.mode = E_ACCUM; p.ep.out = dw; p.ep.ldc = K; p.ep.out_dtype = VITK_F32;
  // split the token reduction so small weight tiles still fill the machine
  const int tiles = ((N + 63) / 64) * ((K + 63) / 64);
  int splits = (4 * sm_count() + tiles - 1) / tiles;
  if (splits > (M + 255) / 256) splits = (M + 255) / 256;
  VITK_TRY(run_gemm(p, engine, splits, st));
  if (db) {
    const MatLayout l = dy_layout == VITK_LAYOUT_HEADMAJOR ? layout_headmajor_rows_m(M) : layout_rowmajor(N);
    VITK_TRY(colsum(dy, dtype, l, M, N, db, st));
  }
  return VITK_OK;
}
