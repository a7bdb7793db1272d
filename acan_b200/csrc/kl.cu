// kl.cu -- attention loss on a MATERIALISED sim_map (HBM-bound form) and the ground-truth affinity.
//
// Replaces AttentionLoss2d.get_similarity / get_gt_sim_map / forward and _BaseKLDivergence.forward
// (reference losses.py:12,18-33,40-62).  The reference makes ~14 elementwise / softmax passes over
// [B,N,N]; here the predicted affinity is read exactly once (4*N*N bytes / image) and the target
//   W_ij = min(a_i/a_j, a_j/a_i) / Z_i,   Z_i = sum_j min(a_i/a_j, a_j/a_i),   rows / columns with a = 0 are zero
// (closed form of softmax_j(-|log a_i - log a_j|) with NaN -> 0) is generated on the
// fly from the B*N stride-8 bin indices.  The fused, tensor-core form that never materialises sim_map
// lives in attention.cu (acan_attention_loss_fused).
//
// One CTA per (b, i) row; row sums by warp shuffle + smem; the final mean is a single-CTA,
// fixed-order reduction, so the loss is bit-reproducible run to run.
#include "common.cuh"

namespace acan {

// bins[b, n] = dis_label[b, 0, 8*(n / w), 8*(n % w)]   -- F.interpolate(mode='nearest') to (H//8, W//8)
__global__ void nearest_bins_kernel(const float* __restrict__ label, float* __restrict__ bins, int B, int H, int W) {
  const int h = H / 8, w = W / 8, N = h * w;
  const int64_t total = (int64_t)B * N;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(q / N), n = (int)(q - (int64_t)b * N);
    bins[q] = label[((int64_t)b * H + 8 * (n / w)) * W + 8 * (n % w)];
  }
}

__device__ __forceinline__ float block_sum(float v, float* red) {   // all threads get the total; blockDim multiple of 32
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

__device__ __forceinline__ float affinity_ratio(float ai, float aj) {
  const float lo = fminf(ai, aj), hi = fmaxf(ai, aj);
  return lo > 0.f ? lo / hi : 0.f;
}

namespace {
constexpr int kThreads = 256;

// MODE 0: write W row.  MODE 1: KL row loss (+ optional gradient w.r.t. sim_map).
template <int MODE>
__global__ void __launch_bounds__(kThreads) gt_row_kernel(const float* __restrict__ bins, const float* __restrict__ sim,
                                                           float* __restrict__ out, float* __restrict__ grad, float grad_scale, int N) {
  __shared__ float red[kThreads / 32];
  const int64_t row = blockIdx.x;                // b*N + i
  const int b = (int)(row / N);
  const float* a = bins + (int64_t)b * N;
  const float ai = a[row - (int64_t)b * N];
  float z = 0.f;
  for (int j = threadIdx.x; j < N; j += kThreads) z += affinity_ratio(ai, a[j]);
  z = block_sum(z, red);
  const float inv_z = z > 0.f ? 1.0f / z : 0.f;
  if constexpr (MODE == 0) {
    float* o = out + row * N;
    for (int j = threadIdx.x; j < N; j += kThreads) o[j] = affinity_ratio(ai, a[j]) * inv_z;
  } else {
    const float* q = sim + row * N;
    float* g = grad ? grad + row * N : nullptr;
    float acc = 0.f;
    for (int j = threadIdx.x; j < N; j += kThreads) {
      const float w = affinity_ratio(ai, a[j]) * inv_z;
      const float qv = __ldg(q + j);
      float gv = 0.f;
      if (w > 0.f) {
        const float ratio = w / qv;
        acc += w * logf(fminf(fmaxf(ratio, 1e-8f), 1e8f));
        if (ratio > 1e-8f && ratio < 1e8f) gv = -ratio * grad_scale;   // d/dq [w log(w/q)] = -w/q
      }
      if (g) g[j] = gv;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[row] = acc;
  }
}
}  // namespace

// loss = scale * sum(vals[0..n)) in a fixed order (one CTA)
__global__ void __launch_bounds__(1024) ordered_sum_kernel(const float* __restrict__ vals, int64_t n, float scale, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += 1024) s += vals[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *out = s * scale;
}

}  // namespace acan

using namespace acan;

static int check_hw(const char* who, int B, int H, int W) {
  if (B <= 0 || H < 8 || W < 8) return fail(ACAN_ERR_BAD_ARG, "%s: B=%d H=%d W=%d invalid (need H, W >= 8)", who, B, H, W);
  return 0;
}

extern "C" int acan_gt_sim_map(const float* dis_label, float* gt, float* bins_ws, int B, int H, int W, void* stream) {
  ACAN_CHECK_ARG(dis_label && gt && bins_ws, "acan_gt_sim_map: null pointer");
  if (int e = check_hw("acan_gt_sim_map", B, H, W)) return e;
  const int N = (H / 8) * (W / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  nearest_bins_kernel<<<(B * N + 255) / 256, 256, 0, st>>>(dis_label, bins_ws, B, H, W);
  if (int e = launch_status("acan_gt_sim_map/bins")) return e;
  gt_row_kernel<0><<<B * N, kThreads, 0, st>>>(bins_ws, nullptr, gt, nullptr, 0.f, N);
  return launch_status("acan_gt_sim_map");
}

extern "C" int acan_attention_loss(const float* sim_map, const float* dis_label, float* loss_out, float* ws,
                                   float* grad_sim, float grad_scale, int B, int H, int W, void* stream) {
  ACAN_CHECK_ARG(sim_map && dis_label && loss_out && ws, "acan_attention_loss: null pointer");
  if (int e = check_hw("acan_attention_loss", B, H, W)) return e;
  const int N = (H / 8) * (W / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(PROF_KL, st);
  float* row_loss = ws;
  float* bins = ws + (int64_t)B * N;
  nearest_bins_kernel<<<(B * N + 255) / 256, 256, 0, st>>>(dis_label, bins, B, H, W);
  if (int e = launch_status("acan_attention_loss/bins")) return e;
  const float inv_rows = 1.0f / ((float)B * (float)N);
  gt_row_kernel<1><<<B * N, kThreads, 0, st>>>(bins, sim_map, row_loss, grad_sim, grad_scale * inv_rows, N);
  if (int e = launch_status("acan_attention_loss/rows")) return e;
  ordered_sum_kernel<<<1, 1024, 0, st>>>(row_loss, (int64_t)B * N, inv_rows, loss_out);
  return launch_status("acan_attention_loss/sum");
}
