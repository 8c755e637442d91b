// Evaluation post-processing on the device: the decision-threshold sweep of
// find_optimal_threshold (/root/reference/train_advanced.py:239-275) and the confusion counts of calculate_metrics
// (/root/reference/test.py:241-243) without shipping per-image scores to the host.
//
// Thresholds ascend (np.linspace), so "score >= thresholds[s]" holds exactly for s < k, k = number of thresholds <= score.
// One pass bins every sample into hist[label == 1][k] (k in 0..steps, per-block shared-memory histogram, one 64-bit global
// add per bin and block); a second, single-block kernel turns the histogram into (tp, fp, tn, fn) per threshold by prefix
// sums.  The comparison runs in float64 like numpy's `probs >= thresh` (float32 array vs float64 scalar): counts are
// bit-exact integers, so every sklearn metric derived from them on the host is bit-exact too.  The histogram accumulates
// (+=) across calls -- one call per evaluation batch, one read-back per epoch.  NaN scores predict spoof for every
// threshold, as in numpy.
#include "common.cuh"

namespace vitk {

constexpr int TS_MAX_STEPS = 255;

__global__ void __launch_bounds__(256)
threshold_hist_kernel(const float* __restrict__ probs, const int64_t* __restrict__ labels, const double* __restrict__ thresholds,
                      int n, int steps, unsigned long long* __restrict__ hist) {
  pdl_sync_traced(TK_EVAL);
  __shared__ unsigned int sh[2 * (TS_MAX_STEPS + 1)];
  __shared__ double th[TS_MAX_STEPS];
  for (int i = threadIdx.x; i < 2 * (steps + 1); i += blockDim.x) sh[i] = 0u;
  for (int i = threadIdx.x; i < steps; i += blockDim.x) th[i] = thresholds[i];
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double p = (double)probs[i];
    // k = |{s : th[s] <= p}| by binary search over the ascending thresholds (NaN: every comparison false -> k = 0)
    int lo = 0, hi = steps;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (p >= th[mid]) lo = mid + 1; else hi = mid;
    }
    atomicAdd(&sh[(labels[i] == 1 ? (steps + 1) : 0) + lo], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * (steps + 1); i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
  trace_end(TK_EVAL);
}

// counts[s] = (tp, fp, tn, fn) at thresholds[s]:  predicted live <=> k > s
__global__ void __launch_bounds__(256)
threshold_counts_kernel(const unsigned long long* __restrict__ hist, int steps, long long* __restrict__ counts) {
  pdl_sync_traced(TK_EVAL);
  const unsigned long long* h0 = hist;               // label spoof (0)
  const unsigned long long* h1 = hist + steps + 1;   // label live (1)
  for (int s = threadIdx.x; s < steps; s += blockDim.x) {
    unsigned long long tp = 0, fp = 0, tn = 0, fn = 0;
    for (int k = 0; k <= steps; ++k) {
      if (k > s) { tp += h1[k]; fp += h0[k]; } else { fn += h1[k]; tn += h0[k]; }
    }
    counts[4 * s + 0] = (long long)tp; counts[4 * s + 1] = (long long)fp;
    counts[4 * s + 2] = (long long)tn; counts[4 * s + 3] = (long long)fn;
  }
  trace_end(TK_EVAL);
}

}  // namespace vitk

using namespace vitk;

extern "C" int vitk_threshold_hist(const float* probs, const int64_t* labels, const double* thresholds, int n, int steps,
                                   unsigned long long* hist, void* stream) {
  VITK_CHECK_ARG(probs && labels && thresholds && hist && n >= 0 && steps >= 1 && steps <= TS_MAX_STEPS);
  if (n == 0) return VITK_OK;
  const int sms = sm_count();
  int grid = (n + 255) / 256;
  if (grid > 2 * sms) grid = 2 * sms;
  VITK_LAUNCH((threshold_hist_kernel), grid, 256, 0, (cudaStream_t)stream, probs, labels, thresholds, n, steps, hist);
  return VITK_OK;
}

extern "C" int vitk_threshold_counts(const unsigned long long* hist, int steps, long long* counts, void* stream) {
  VITK_CHECK_ARG(hist && counts && steps >= 1 && steps <= TS_MAX_STEPS);
  VITK_LAUNCH((threshold_counts_kernel), 1, 256, 0, (cudaStream_t)stream, hist, steps, counts);
  return VITK_OK;
}
