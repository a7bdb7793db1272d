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

__global__ void nearest_logbins_kernel(const float* __restrict__ label, float* __restrict__ logbin, int B, int H, int W) {
  const int h = H / 8, w = W / 8, N = h * w;
  const int64_t total = (int64_t)B * N;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(q / N), n = (int)(q - (int64_t)b * N);
    const float a = label[((int64_t)b * H + 8 * (n / w)) * W + 8 * (n % w)];
    logbin[q] = a > 0.f ? logf(a) : -INFINITY;
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

// write one row of the target affinity W
__global__ void __launch_bounds__(kThreads) gt_row_kernel(const float* __restrict__ bins, float* __restrict__ out, int N) {
  __shared__ float red[kThreads / 32];
  const int64_t row = blockIdx.x;                // b*N + i
  const int b = (int)(row / N);
  const float* a = bins + (int64_t)b * N;
  const float ai = a[row - (int64_t)b * N];
  float z = 0.f;
  for (int j = threadIdx.x; j < N; j += kThreads) z += affinity_ratio(ai, a[j]);
  z = block_sum(z, red);
  const float inv_z = z > 0.f ? 1.0f / z : 0.f;
  float* o = out + row * N;
  for (int j = threadIdx.x; j < N; j += kThreads) o[j] = affinity_ratio(ai, a[j]) * inv_z;
}

// KL row loss (+ optional gradient w.r.t. sim_map) in the log domain: with la = ln(bin) (-inf for bin 0)
//   ln W_ij = -|la_i - la_j| - ln Z_i,   Z_i = sum_j exp(-|la_i - la_j|),   term = W (clamp(ln W - ln Q, +-ln 1e8))
// one ex2 + one lg2 per element instead of two divisions and a logf: the pass stays HBM-bound (4*N*N bytes / image).
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f, kLnClamp = 18.420680744f;

__global__ void __launch_bounds__(kThreads) kl_row_kernel(const float* __restrict__ logbin, const float* __restrict__ sim,
                                                           float* __restrict__ row_loss, float* __restrict__ grad, float grad_scale, int N) {
  __shared__ float red[kThreads / 32];
  const int64_t row = blockIdx.x;
  const int b = (int)(row / N);
  const float* la = logbin + (int64_t)b * N;
  const float la_i = la[row - (int64_t)b * N];
  const float* q = sim + row * N;
  float* g = grad ? grad + row * N : nullptr;
  if (!(la_i > -INFINITY)) {                     // bin 0: all-zero target row -> zero loss, zero gradient
    if (g) for (int j = threadIdx.x; j < N; j += kThreads) g[j] = 0.f;
    if (threadIdx.x == 0) row_loss[row] = 0.f;
    return;
  }
  float z = 0.f;
  for (int j = threadIdx.x; j < N; j += kThreads) z += ex2_approx(-fabsf(la_i - la[j]) * kLog2e);   // exp(-inf) = 0 for bin 0
  z = block_sum(z, red);
  const float log2_z = lg2_approx(z);            // z >= 1 (the j = i term)
  float acc = 0.f;
  for (int j = threadIdx.x; j < N; j += kThreads) {
    const float l2w = -fabsf(la_i - la[j]) * kLog2e - log2_z;        // log2 W_ij
    const float w = ex2_approx(l2w);
    const float qv = __ldg(q + j);
    float gv = 0.f;
    if (w > 0.f) {
      const float lr = (l2w - lg2_approx(qv)) * kLn2;                // ln(W/Q); +inf when Q = 0
      acc += w * fminf(fmaxf(lr, -kLnClamp), kLnClamp);
      if (lr > -kLnClamp && lr < kLnClamp) gv = -__fdividef(w, qv) * grad_scale;   // d/dq [w ln(w/q)] = -w/q inside the clamp window
    }
    if (g) g[j] = gv;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) row_loss[row] = acc;
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
  // the kernels read label pixel (8 i, 8 j): F.interpolate(mode='nearest') to (H//8, W//8) picks exactly that one only when H and W
  // are multiples of 8 (otherwise its source index is floor(i * H / (H//8))); the network itself needs multiples of 8 (README.md:52)
  if (B <= 0 || H < 8 || W < 8 || (H % 8) != 0 || (W % 8) != 0)
    return fail(ACAN_ERR_BAD_ARG, "%s: B=%d H=%d W=%d invalid (need H, W positive multiples of 8)", who, B, H, W);
  return 0;
}

extern "C" int acan_gt_sim_map(const float* dis_label, float* gt, float* bins_ws, int B, int H, int W, void* stream) {
  ACAN_CHECK_ARG(dis_label && gt && bins_ws, "acan_gt_sim_map: null pointer");
  if (int e = check_hw("acan_gt_sim_map", B, H, W)) return e;
  const int N = (H / 8) * (W / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  nearest_bins_kernel<<<(B * N + 255) / 256, 256, 0, st>>>(dis_label, bins_ws, B, H, W);
  if (int e = launch_status("acan_gt_sim_map/bins")) return e;
  gt_row_kernel<<<B * N, kThreads, 0, st>>>(bins_ws, gt, N);
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
  float* logbin = ws + (int64_t)B * N;
  nearest_logbins_kernel<<<(B * N + 255) / 256, 256, 0, st>>>(dis_label, logbin, B, H, W);
  if (int e = launch_status("acan_attention_loss/bins")) return e;
  const float inv_rows = 1.0f / ((float)B * (float)N);
  kl_row_kernel<<<B * N, kThreads, 0, st>>>(logbin, sim_map, row_loss, grad_sim, grad_scale * inv_rows, N);
  if (int e = launch_status("acan_attention_loss/rows")) return e;
  ordered_sum_kernel<<<1, 1024, 0, st>>>(row_loss, (int64_t)B * N, inv_rows, loss_out);
  return launch_status("acan_attention_loss/sum");
}
