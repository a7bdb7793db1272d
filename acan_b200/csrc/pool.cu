// pool.cu -- image-level pooling branch of the context aggregation module.
//
// Replaces SADecoder.image_context (reference sadecoder.py:97-106):
//   AvgPool2d(h,w) -> Dropout2d(0.5) -> Linear(2048,512)+ReLU -> Linear(512,512)+ReLU -> Upsample(h,w)
// The upsample of a 1x1 map is a spatial broadcast; it is never written -- the merge step
// consumes g[B,512] as a per-image bias (acan_b200/modules.py).
//
// Kernel 1 (HBM-bound, B*C*N*4 bytes): one warp per (b,c) row of N contiguous floats, 16 B loads,
//           shuffle reduction, optional keep-mask * 2.  Kernel 2/3: one warp per output neuron, the
//           weight row is read once (coalesced 16 B loads) and dotted with all B inputs.
#include "common.cuh"

namespace acan {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__global__ void __launch_bounds__(kThreads) row_mean_kernel(const float* __restrict__ x, const float* __restrict__ keep,
                                                             float* __restrict__ mean, int64_t rows, int N) {
  const int lane = threadIdx.x & 31;
  const float inv_n = 1.0f / (float)N;
  for (int64_t r = blockIdx.x * (int64_t)kWarps + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * kWarps) {
    const float* row = x + r * N;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if ((N & 3) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
      const float4* r4 = reinterpret_cast<const float4*>(row);
      const int n4 = N >> 2;
      int i = lane;
      for (; i + 96 < n4; i += 128) {       // 4 independent 16 B loads in flight per lane
        const float4 a = ldg_stream(r4 + i), b = ldg_stream(r4 + i + 32), c = ldg_stream(r4 + i + 64), d = ldg_stream(r4 + i + 96);
        s0 += (a.x + a.y) + (a.z + a.w); s1 += (b.x + b.y) + (b.z + b.w);
        s2 += (c.x + c.y) + (c.z + c.w); s3 += (d.x + d.y) + (d.z + d.w);
      }
      for (; i < n4; i += 32) { const float4 a = ldg_stream(r4 + i); s0 += (a.x + a.y) + (a.z + a.w); }
    } else {
      for (int i = lane; i < N; i += 32) s0 += row[i];
    }
    float s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) {
      float m = s * inv_n;
      if (keep != nullptr) m *= keep[r] * 2.0f;   // Dropout2d(0.5): kept channels scaled by 1/(1-p)
      mean[r] = m;
    }
  }
}

// out[b,o] = relu(bias[o] + sum_i W[o,i] * in[b,i]);  one warp per o
__global__ void __launch_bounds__(kThreads) linear_relu_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                                const float* __restrict__ bias, float* __restrict__ out,
                                                                int B, int I, int O) {
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (o >= O) return;
  const float* w = W + (int64_t)o * I;
  const float bo = bias[o];
  for (int b = 0; b < B; ++b) {
    const float* v = in + (int64_t)b * I;
    float acc = 0.f;
    if ((I & 3) == 0) {
      const float4* w4 = reinterpret_cast<const float4*>(w);
      const float4* v4 = reinterpret_cast<const float4*>(v);
      for (int i = lane; i < (I >> 2); i += 32) {
        const float4 a = __ldg(w4 + i), c = __ldg(v4 + i);
        acc += a.x * c.x + a.y * c.y + a.z * c.z + a.w * c.w;
      }
    } else {
      for (int i = lane; i < I; i += 32) acc += w[i] * v[i];
    }
    acc = warp_sum(acc);
    if (lane == 0) out[(int64_t)b * O + o] = fmaxf(acc + bo, 0.f);
  }
}

}  // namespace
}  // namespace acan

using namespace acan;

extern "C" int acan_pool_branch_fwd(const float* x, const float* keep_mask, const float* w1, const float* b1,
                                    const float* w2, const float* b2, float* mean_out, float* hidden_out, float* g_out,
                                    int B, int C, int N, int H1, int H2, void* stream) {
  ACAN_CHECK_ARG(x && w1 && b1 && w2 && b2 && mean_out && hidden_out && g_out, "acan_pool_branch_fwd: null pointer");
  ACAN_CHECK_ARG(B > 0 && C > 0 && N > 0 && H1 > 0 && H2 > 0, "acan_pool_branch_fwd: sizes must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(PROF_POOL, st);
  const int64_t rows = (int64_t)B * C;
  int64_t blocks = (rows + kWarps - 1) / kWarps;
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  row_mean_kernel<<<(int)blocks, kThreads, 0, st>>>(x, keep_mask, mean_out, rows, N);
  if (int e = launch_status("acan_pool_branch_fwd/mean")) return e;
  linear_relu_kernel<<<(H1 + kWarps - 1) / kWarps, kThreads, 0, st>>>(mean_out, w1, b1, hidden_out, B, C, H1);
  if (int e = launch_status("acan_pool_branch_fwd/linear1")) return e;
  linear_relu_kernel<<<(H2 + kWarps - 1) / kWarps, kThreads, 0, st>>>(hidden_out, w2, b2, g_out, B, H1, H2);
  return launch_status("acan_pool_branch_fwd/linear2");
}
