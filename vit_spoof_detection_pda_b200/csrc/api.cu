// libvitk: version, error reporting, device query.
#include <stdarg.h>
#include <atomic>
#include <mutex>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace vitk {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(); }
static thread_local int t_sm_budget = 0;
static int hw_sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev == cached_dev && cached) return cached;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached_dev = dev;
  cached = n;
  return n;
}
// SMs the persistent kernels size their grids for: all of them, or the budget of the calling thread / call
// (vitk_model.sm_budget, vitk_set_sm_budget) while a communication kernel (NCCL all-reduce) owns some SMs -- a
// persistent CTA that cannot become resident would stall its statically assigned tiles until the collective
// finishes.  Thread-local: no process-wide mutable state on the data path.
int sm_count() {
  const int hw = hw_sm_count(), b = t_sm_budget;
  return (b > 0 && b < hw) ? b : hw;
}
int set_sm_budget(int n) { const int prev = t_sm_budget; t_sm_budget = n; return prev; }
// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to a function on ONE device: set it once per (function, device)
int set_max_dyn_smem_once(const void* fn, int bytes) {
  struct Done { const void* fn; int dev; };
  static Done done[256];
  static int n_done = 0;
  static std::mutex mu;
  int dev = 0;
  VITK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  for (int i = 0; i < n_done; ++i)
    if (done[i].fn == fn && done[i].dev == dev) return VITK_OK;
  VITK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  if (n_done < 256) done[n_done++] = Done{fn, dev};
  return VITK_OK;
}
#ifdef VITK_DEV
// device-side tracer: one setter per translation unit (each has its own copy of the device pointer)
static void (*g_trace_setters[32])(unsigned long long*);
static int g_n_trace_setters = 0;
void trace_register(void (*setter)(unsigned long long*)) {
  if (g_n_trace_setters < 32) g_trace_setters[g_n_trace_setters++] = setter;
}
static void trace_set_all(unsigned long long* p) {
  for (int i = 0; i < g_n_trace_setters; ++i) g_trace_setters[i](p);
}
#endif
// ---- tensor-map cache: open addressing over FNV-1a of the key bytes
struct TmapEntry { TmapKey key; unsigned char map[128]; bool used; };
static constexpr int TMAP_SLOTS = 8192;
static TmapEntry* g_tmap = nullptr;
static int g_tmap_count = 0;
static std::mutex g_tmap_mu;
static uint64_t tmap_hash(const TmapKey& k) {
  const unsigned char* b = reinterpret_cast<const unsigned char*>(&k);
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < sizeof(TmapKey); ++i) { h ^= b[i]; h *= 1099511628211ull; }
  return h;
}
bool tmap_cache_get(const TmapKey& key, void* map128) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (!g_tmap) return false;
  for (uint64_t i = tmap_hash(key) % TMAP_SLOTS, n = 0; n < TMAP_SLOTS; i = (i + 1) % TMAP_SLOTS, ++n) {
    const TmapEntry& e = g_tmap[i];
    if (!e.used) return false;
    if (memcmp(&e.key, &key, sizeof(TmapKey)) == 0) { memcpy(map128, e.map, 128); return true; }
  }
  return false;
}
void tmap_cache_put(const TmapKey& key, const void* map128) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (!g_tmap) g_tmap = static_cast<TmapEntry*>(calloc(TMAP_SLOTS, sizeof(TmapEntry)));
  if (!g_tmap) return;
  if (g_tmap_count >= TMAP_SLOTS / 2) { memset(g_tmap, 0, sizeof(TmapEntry) * TMAP_SLOTS); g_tmap_count = 0; }   // shapes changed a lot: start over
  for (uint64_t i = tmap_hash(key) % TMAP_SLOTS;; i = (i + 1) % TMAP_SLOTS) {
    TmapEntry& e = g_tmap[i];
    if (!e.used) { e.key = key; memcpy(e.map, map128, 128); e.used = true; ++g_tmap_count; return; }
    if (memcmp(&e.key, &key, sizeof(TmapKey)) == 0) return;
  }
}
static std::atomic<int> g_pdl{1};
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }
void set_pdl(int on) { g_pdl.store(on); }
}  // namespace vitk

extern "C" {
int vitk_version(void) { return VITK_VERSION; }
int vitk_set_sm_budget(int n) { return vitk::set_sm_budget(n); }
int vitk_is_dev_build(void) {
#ifdef VITK_DEV
  return 1;
#else
  return 0;
#endif
}
// Device-side tracer (development build only): `buf` = device buffer of `bytes` bytes, cleared here; every CTA of every
// libvitk kernel appends three 64-bit words per mark (see common.cuh) until vitk_trace_stop().  Both calls synchronise the device.
int vitk_trace_start(void* buf, size_t bytes) {
#ifdef VITK_DEV
  // layout (64-bit words): [0] marks written, [1] mark capacity, [2] word offset of the detail area, [3] unused,
  // [4 ..] marks of 3 words, then the detail area: 32 kernels x 24 phases x 64 slots
  const size_t detail_words = 32 * 24 * 64;
  VITK_CHECK_ARG(buf && bytes >= (detail_words + 64) * 8);
  VITK_CUDA(cudaDeviceSynchronize());
  const size_t words = bytes / 8;
  const unsigned long long cap = (words - 4 - detail_words) / 3;
  const unsigned long long hdr[4] = {0ull, cap, 4 + 3 * cap, 0ull};
  VITK_CUDA(cudaMemcpy(buf, hdr, sizeof(hdr), cudaMemcpyHostToDevice));
  VITK_CUDA(cudaMemset(reinterpret_cast<unsigned long long*>(buf) + hdr[2], 0, detail_words * 8));
  vitk::trace_set_all(reinterpret_cast<unsigned long long*>(buf));
  VITK_CUDA(cudaDeviceSynchronize());
  return VITK_OK;
#else
  (void)buf; (void)bytes;
  vitk::set_error("vitk_trace_start: the tracer exists only in the development build (libvitk_dev.so)");
  return VITK_ERR_UNSUPPORTED;
#endif
}
int vitk_trace_stop(void) {
#ifdef VITK_DEV
  VITK_CUDA(cudaDeviceSynchronize());
  vitk::trace_set_all(nullptr);
  VITK_CUDA(cudaDeviceSynchronize());
#endif
  return VITK_OK;
}
long long vitk_launch_count(void) { return vitk::launches(); }
const char* vitk_last_error_string(void) { return vitk::g_err; }
int vitk_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  VITK_CUDA(cudaGetDevice(&dev));
  if (sm_count) VITK_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) VITK_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) VITK_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return VITK_OK;
}
}
