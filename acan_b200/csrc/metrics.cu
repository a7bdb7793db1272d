// metrics.cu -- the eight depth-error measures of the reference's evaluation loop in one pass.
//
// Replaces utils/eval_utils.py:16-48 (compute_errors), called once per validation / test batch
// (depthest_trainer.py:185-190, 222-226): boolean-mask indexing (a host synchronisation) followed by ~25
// elementwise / reduction kernels over the masked copies.  Here: one grid-stride pass over gt and pred (8 B per pixel, HBM-bound),
// nine fp32 partial sums per block, a fixed-order double-precision finalisation.  Semantics kept literally:
//   mask = gt > 0;  thresh = max(gt/pred, pred/gt);  a_k = mean(thresh < 1.25^k);  rmse = sqrt(mean((gt-pred)^2));
//   rmse_log = sqrt(mean((slog gt - slog pred)^2));  log10 = mean|slog gt - slog pred|   (the reference's "log10" is a NATURAL log,
//   eval_utils.py:31);  abs_rel = mean(|gt-pred|/gt);  sq_rel = mean((gt-pred)^2/gt);  slog x = log(clamp(x, 1e-6, 1e6));
//   every mean is over the masked pixels of the whole batch and is multiplied by batch_size (eval_utils.py:37-46).
#include "common.cuh"
#include <cmath>

namespace acan {
namespace {

constexpr int kMetThreads = 256;
constexpr int kMetSums = 9;          // count, a1, a2, a3, sq, sqlog, abslog, absrel, sqrel

__device__ __forceinline__ void metric_terms(float g, float p, float (&s)[kMetSums]) {
  if (!(g > 0.f)) return;
  const float t = fmaxf(g / p, p / g);
  const float d = g - p, l = logf(fminf(fmaxf(g, 1e-6f), 1e6f)) - logf(fminf(fmaxf(p, 1e-6f), 1e6f));
  s[0] += 1.f;
  s[1] += (t < 1.25f) ? 1.f : 0.f;
  s[2] += (t < 1.25f * 1.25f) ? 1.f : 0.f;
  s[3] += (t < 1.25f * 1.25f * 1.25f) ? 1.f : 0.f;
  s[4] += d * d;
  s[5] += l * l;
  s[6] += fabsf(l);
  s[7] += fabsf(d) / g;
  s[8] += d * d / g;
}

__global__ void __launch_bounds__(kMetThreads) metrics_partial_kernel(const float* __restrict__ gt, const float* __restrict__ pred, float* __restrict__ partial,
                                                                       int64_t n) {
  __shared__ float red[kMetThreads / 32][kMetSums];
  float s[kMetSums];
#pragma unroll
  for (int k = 0; k < kMetSums; ++k) s[k] = 0.f;
  const int64_t n4 = ((reinterpret_cast<uintptr_t>(gt) | reinterpret_cast<uintptr_t>(pred)) % 16 == 0) ? n / 4 : 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(kMetThreads) + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * kMetThreads) {
    const float4 g = ldg_stream(reinterpret_cast<const float4*>(gt) + i), p = ldg_stream(reinterpret_cast<const float4*>(pred) + i);
    metric_terms(g.x, p.x, s); metric_terms(g.y, p.y, s); metric_terms(g.z, p.z, s); metric_terms(g.w, p.w, s);
  }
  for (int64_t i = 4 * n4 + blockIdx.x * static_cast<int64_t>(kMetThreads) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * kMetThreads)
    metric_terms(__ldg(gt + i), __ldg(pred + i), s);
#pragma unroll
  for (int k = 0; k < kMetSums; ++k) s[k] = warp_sum(s[k]);
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < kMetSums; ++k) red[threadIdx.x >> 5][k] = s[k];
  }
  __syncthreads();
  if (threadIdx.x < kMetSums) {
    float t = 0.f;
    for (int w = 0; w < kMetThreads / 32; ++w) t += red[w][threadIdx.x];
    partial[static_cast<size_t>(blockIdx.x) * kMetSums + threadIdx.x] = t;
  }
}

// out[8] in the reference's measure_list order: a1, a2, a3, rmse, rmse_log, log10, abs_rel, sq_rel
__global__ void metrics_final_kernel(const float* __restrict__ partial, int blocks, float batch, float* __restrict__ out) {
  __shared__ double tot[kMetSums];
  if (threadIdx.x < kMetSums) {
    double t = 0.0;
    for (int b = 0; b < blocks; ++b) t += static_cast<double>(partial[static_cast<size_t>(b) * kMetSums + threadIdx.x]);
    tot[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double c = tot[0];                      // 0 valid pixels: 0/0 = NaN, as torch's mean over an empty selection
    out[0] = static_cast<float>(tot[1] / c * batch);
    out[1] = static_cast<float>(tot[2] / c * batch);
    out[2] = static_cast<float>(tot[3] / c * batch);
    out[3] = static_cast<float>(sqrt(tot[4] / c) * batch);
    out[4] = static_cast<float>(sqrt(tot[5] / c) * batch);
    out[5] = static_cast<float>(tot[6] / c * batch);
    out[6] = static_cast<float>(tot[7] / c * batch);
    out[7] = static_cast<float>(tot[8] / c * batch);
  }
}

}  // namespace
}  // namespace acan

extern "C" int acan_depth_metrics_ws_floats(void) { return acan::kNumSMs * 8 * acan::kMetSums; }

extern "C" int acan_depth_metrics(const float* gt, const float* pred, float* out8, float* ws, int64_t n, int batch_size, void* stream) {
  ACAN_CHECK_ARG(gt && pred && out8 && ws, "acan_depth_metrics: null pointer");
  ACAN_CHECK_ARG(n > 0 && batch_size > 0, "acan_depth_metrics: n=%lld batch_size=%d must be positive", (long long)n, batch_size);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t blocks = (n / 4 + acan::kMetThreads - 1) / acan::kMetThreads;
  if (blocks > acan::kNumSMs * 8) blocks = acan::kNumSMs * 8;
  if (blocks < 1) blocks = 1;
  acan::metrics_partial_kernel<<<static_cast<int>(blocks), acan::kMetThreads, 0, st>>>(gt, pred, ws, n);
  if (int e = acan::launch_status("acan_depth_metrics/partial")) return e;
  acan::metrics_final_kernel<<<1, 32, 0, st>>>(ws, static_cast<int>(blocks), static_cast<float>(batch_size), out8);
  return acan::launch_status("acan_depth_metrics/final");
}
