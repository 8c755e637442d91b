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
static std::atomic<int> g_sm_budget{0};
static int hw_sm_count() {
  static int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached = n;
  return n;
}
// SMs the persistent kernels size their grids for: all of them, or the budget set while a communication
// kernel (NCCL all-reduce) owns some SMs -- a persistent CTA that cannot become resident would stall its
// statically assigned tiles until the collective finishes.
int sm_count() {
  const int hw = hw_sm_count(), b = g_sm_budget.load(std::memory_order_relaxed);
  return (b > 0 && b < hw) ? b : hw;
}
int set_sm_budget(int n) { return g_sm_budget.exchange(n); }
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
