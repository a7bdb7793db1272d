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
#include <atomic>
#include <cuda_bf16.h>

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

// out[b,o] = relu(bias[o] + sum_i W[o,i] * in[b,i]).  A CTA of 8 warps stages the input rows of up to 16 images in shared
// memory ONCE (every output neuron needs all of them: re-reading them per neuron made the first version L2-bound at
// 64 MB of redundant traffic), then each warp streams one weight row from HBM into registers (NI x 16 B per lane, all
// loads in flight) and dots it with the staged rows.
constexpr int kLinWarps = 8;
constexpr int kLinImages = 16;

template <int NI>   // I == NI * 128
__global__ void __launch_bounds__(kLinWarps * 32) linear_relu_smem_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                                         const float* __restrict__ bias, float* __restrict__ out, int B, int O) {
  extern __shared__ float4 xin[];                       // [images][I / 4]
  constexpr int I4 = NI * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int o = blockIdx.x * kLinWarps + warp;
  float4 w[NI];
  if (o < O) {
    const float4* w4 = reinterpret_cast<const float4*>(W + (int64_t)o * (I4 * 4));
#pragma unroll
    for (int k = 0; k < NI; ++k) w[k] = ldg_stream(w4 + lane + 32 * k);
  }
  const float bo = o < O ? bias[o] : 0.f;
  for (int b0 = 0; b0 < B; b0 += kLinImages) {
    const int nb = min(kLinImages, B - b0);
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(in + (int64_t)b0 * (I4 * 4));
    for (int i = threadIdx.x; i < nb * I4; i += kLinWarps * 32) xin[i] = __ldg(src + i);
    __syncthreads();
    if (o < O) {
      for (int b = 0; b < nb; ++b) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < NI; ++k) {
          const float4 v = xin[b * I4 + lane + 32 * k];
          acc += w[k].x * v.x + w[k].y * v.y + w[k].z * v.z + w[k].w * v.w;
        }
        acc = warp_sum(acc);
        if (lane == 0) out[(int64_t)(b0 + b) * O + o] = fmaxf(acc + bo, 0.f);
      }
    }
  }
}

// generic sizes: one warp per o, strided loop
__global__ void __launch_bounds__(128) linear_relu_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                          const float* __restrict__ bias, float* __restrict__ out, int B, int I, int O) {
  const int lane = threadIdx.x & 31;
  const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= O) return;
  const float* w = W + (int64_t)o * I;
  const float bo = bias[o];
  for (int b = 0; b < B; ++b) {
    const float* v = in + (int64_t)b * I;
    float acc = 0.f;
    for (int i = lane; i < I; i += 32) acc += w[i] * v[i];
    acc = warp_sum(acc);
    if (lane == 0) out[(int64_t)b * O + o] = fmaxf(acc + bo, 0.f);
  }
}

template <int NI>
int launch_linear_smem(const float* in, const float* W, const float* bias, float* out, int B, int O, cudaStream_t st, const char* what) {
  const size_t smem = static_cast<size_t>(B < kLinImages ? B : kLinImages) * NI * 128 * sizeof(float);
  static std::atomic<uint64_t> attr_done{0};                 // the opt-in is per device: one bit per device ordinal
  int dev = 0;
  ACAN_CUDA(cudaGetDevice(&dev));
  const uint64_t bit = 1ull << (dev & 63);
  if (!(attr_done.load(std::memory_order_acquire) & bit)) {
    ACAN_CUDA(cudaFuncSetAttribute(linear_relu_smem_kernel<NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLinImages * NI * 128 * 4));
    attr_done.fetch_or(bit, std::memory_order_release);
  }
  linear_relu_smem_kernel<NI><<<(O + kLinWarps - 1) / kLinWarps, kLinWarps * 32, smem, st>>>(in, W, bias, out, B, O);
  return launch_status(what);
}

int launch_linear(const float* in, const float* W, const float* bias, float* out, int B, int I, int O, cudaStream_t st, const char* what) {
  const bool aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(W)) & 15) == 0;
  if (I == 2048 && aligned) return launch_linear_smem<16>(in, W, bias, out, B, O, st, what);
  if (I == 1024 && aligned) return launch_linear_smem<8>(in, W, bias, out, B, O, st, what);
  if (I == 512 && aligned) return launch_linear_smem<4>(in, W, bias, out, B, O, st, what);
  linear_relu_kernel<<<(O + 3) / 4, 128, 0, st>>>(in, W, bias, out, B, I, O);
  return launch_status(what);
}

// ---- channels-last half-precision input: x [B, N, C] fp16 / bf16 (what a channels_last autocast pipeline holds).
// Column sums: a thread owns 8 adjacent channels (one 16 B load per row), a CTA a slab of kSlabRows rows; slab partials
// go to scratch and are reduced in a fixed order (bit-reproducible, no atomics).  Algorithmic bytes B*N*C*2.
constexpr int kSlabRows = 32;

template <bool BF16>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (BF16) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    } else {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      const float2 t = __half22float2(h);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
}

template <bool BF16>
__global__ void __launch_bounds__(256, 4) slab_colsum_kernel(const uint4* __restrict__ x, float* __restrict__ partial, int N, int C8, int slabs) {
  const int c8 = blockIdx.x * 256 + threadIdx.x;
  const int slab = blockIdx.y, b = blockIdx.z;
  if (c8 >= C8) return;
  const int r0 = slab * kSlabRows, r1 = min(N, r0 + kSlabRows);
  const uint4* src = x + (static_cast<int64_t>(b) * N + r0) * C8 + c8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int r = r0; r < r1; r += 8) {                  // 8 independent 16 B loads in flight (volatile asm: ptxas keeps them batched);
    uint4 u[8];                                       // rows beyond the slab re-read row r and are zeroed
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      u[k] = ldg_nc_u4(src + static_cast<int64_t>((r + k < r1 ? r + k : r) - r0) * C8);
      if (r + k >= r1) u[k] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float f[8];
      unpack8<BF16>(u[k], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
  }
  float4* dst = reinterpret_cast<float4*>(partial + ((static_cast<int64_t>(b) * slabs + slab) * C8 + c8) * 8);
  dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

__global__ void __launch_bounds__(256) slab_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ keep,
                                                           float* __restrict__ mean, int C, int slabs, int N, int64_t total) {
  for (int64_t q = blockIdx.x * 256LL + threadIdx.x; q < total; q += static_cast<int64_t>(gridDim.x) * 256) {
    const int64_t b = q / C, c = q - b * C;
    float s = 0.f;
    for (int t = 0; t < slabs; ++t) s += partial[(b * slabs + t) * C + c];
    float m = s / static_cast<float>(N);
    if (keep != nullptr) m *= keep[q] * 2.0f;
    mean[q] = m;
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
  if (int e = launch_linear(mean_out, w1, b1, hidden_out, B, C, H1, st, "acan_pool_branch_fwd/linear1")) return e;
  return launch_linear(hidden_out, w2, b2, g_out, B, H1, H2, st, "acan_pool_branch_fwd/linear2");
}

extern "C" size_t acan_pool_nhwc_scratch_floats(int B, int C, int N) {
  if (B <= 0 || C <= 0 || N <= 0) return 0;
  return static_cast<size_t>(B) * ((N + kSlabRows - 1) / kSlabRows) * C;
}

extern "C" int acan_pool_branch_fwd_nhwc(const void* x, int x_dtype, const float* keep_mask, const float* w1, const float* b1,
                                         const float* w2, const float* b2, float* scratch, float* mean_out, float* hidden_out, float* g_out,
                                         int B, int C, int N, int H1, int H2, void* stream) {
  ACAN_CHECK_ARG(x && w1 && b1 && w2 && b2 && scratch && mean_out && hidden_out && g_out, "acan_pool_branch_fwd_nhwc: null pointer");
  ACAN_CHECK_ARG(B > 0 && C > 0 && (C % 8) == 0 && N > 0 && H1 > 0 && H2 > 0, "acan_pool_branch_fwd_nhwc: sizes must be positive, C a multiple of 8");
  ACAN_CHECK_ARG(x_dtype == 1 || x_dtype == 2, "acan_pool_branch_fwd_nhwc: x_dtype must be 1 (fp16) or 2 (bf16)");
  ACAN_CHECK_ARG(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(scratch) % 16 == 0, "acan_pool_branch_fwd_nhwc: 16-byte alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(PROF_POOL, st);
  const int slabs = (N + kSlabRows - 1) / kSlabRows, C8 = C / 8;
  const dim3 grid((C8 + 255) / 256, slabs, B);
  if (x_dtype == 2) slab_colsum_kernel<true><<<grid, 256, 0, st>>>(static_cast<const uint4*>(x), scratch, N, C8, slabs);
  else              slab_colsum_kernel<false><<<grid, 256, 0, st>>>(static_cast<const uint4*>(x), scratch, N, C8, slabs);
  if (int e = launch_status("acan_pool_branch_fwd_nhwc/colsum")) return e;
  const int64_t total = static_cast<int64_t>(B) * C;
  slab_reduce_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, st>>>(scratch, keep_mask, mean_out, C, slabs, N, total);
  if (int e = launch_status("acan_pool_branch_fwd_nhwc/reduce")) return e;
  if (int e = launch_linear(mean_out, w1, b1, hidden_out, B, C, H1, st, "acan_pool_branch_fwd_nhwc/linear1")) return e;
  return launch_linear(hidden_out, w2, b2, g_out, B, H1, H2, st, "acan_pool_branch_fwd_nhwc/linear2");
}


// ---------------------------------------------------------------- merge epilogue (sadecoder.py:107-111, 120-121)
// The pooled branch g is spatially constant, so its share of the merge convolution ReLU(conv1x1(cat(context, g))) is a per-image
// bias  W[:, dv:] g + b  added to the GEMM over the context channels.  This kernel applies that [B, C] bias and the ReLU in place
// on the channels-last half-precision GEMM output: one read + one write of y (the framework's broadcast add + ReLU took two
// passes, the add through a non-vectorised kernel: 154 us -> one HBM pass at C2).
namespace acan {
namespace {

template <bool BF16>
__global__ void __launch_bounds__(256) bias_relu_nhwc_kernel(uint4* __restrict__ y, const float* __restrict__ bias, int64_t n8_per_image, int C) {
  const int c8 = C >> 3;
  const float* bb = bias + static_cast<size_t>(blockIdx.y) * C;
  uint4* yb = y + static_cast<size_t>(blockIdx.y) * n8_per_image;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n8_per_image; i += static_cast<int64_t>(gridDim.x) * 256) {
    const int c = static_cast<int>(i % c8) * 8;
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bb + c)), b1 = __ldg(reinterpret_cast<const float4*>(bb + c) + 1);
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    const uint4 u = yb[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float lo, hi;
      if (BF16) {
        lo = __uint_as_float(w[e] << 16);
        hi = __uint_as_float(w[e] & 0xffff0000u);
      } else {
        const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
        lo = t.x;
        hi = t.y;
      }
      lo = fmaxf(lo + bv[2 * e], 0.f);
      hi = fmaxf(hi + bv[2 * e + 1], 0.f);
      if (BF16) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        o[e] = *reinterpret_cast<const uint32_t*>(&h);
      } else {
        const __half2 h = __floats2half2_rn(lo, hi);
        o[e] = *reinterpret_cast<const uint32_t*>(&h);
      }
    }
    yb[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace
}  // namespace acan

extern "C" int acan_bias_relu_nhwc(void* y, int dtype, const float* bias, int B, int N, int C, void* stream) {
  ACAN_CHECK_ARG(y && bias, "acan_bias_relu_nhwc: null pointer");
  ACAN_CHECK_ARG(B > 0 && B <= 65535 && N > 0 && C > 0 && (C % 8) == 0, "acan_bias_relu_nhwc: B=%d N=%d C=%d invalid (C %% 8 == 0)", B, N, C);
  ACAN_CHECK_ARG(dtype == 1 || dtype == 2, "acan_bias_relu_nhwc: dtype %d (1 = fp16, 2 = bf16)", dtype);
  ACAN_CHECK_ARG(reinterpret_cast<uintptr_t>(y) % 16 == 0 && reinterpret_cast<uintptr_t>(bias) % 16 == 0, "acan_bias_relu_nhwc: 16-byte alignment");
  const int64_t n8 = static_cast<int64_t>(N) * (C / 8);
  int64_t gx = (n8 + 255) / 256;
  const int64_t cap = (static_cast<int64_t>(acan::kNumSMs) * 16 + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  const dim3 grid(static_cast<unsigned>(gx), B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == 2) acan::bias_relu_nhwc_kernel<true><<<grid, 256, 0, st>>>(static_cast<uint4*>(y), bias, n8, C);
  else            acan::bias_relu_nhwc_kernel<false><<<grid, 256, 0, st>>>(static_cast<uint4*>(y), bias, n8, C);
  return acan::launch_status("acan_bias_relu_nhwc");
}


// ---------------------------------------------------------------- stem max pooling (model.py:170, 218: nn.MaxPool2d(3, 2, 1))
// 3x3 / stride 2 / pad 1 max pooling of the stem's channels-last half-precision output, inference only (no indices): a thread owns
// 8 channels of one output pixel and reads its <= 9 taps as 16 B vectors (neighbouring outputs share taps through L1 / L2).
// HBM-bound: 46 MB read + 11.5 MB written at configs[1]; the framework's NHWC kernel takes 113 us there.
namespace acan {
namespace {

template <bool BF16>
__global__ void __launch_bounds__(256) maxpool3x3s2_nhwc_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int H, int W, int Ho, int Wo, int c8,
                                                                 int64_t total) {
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * 256) {
    const int cg = static_cast<int>(i % c8);
    int64_t t = i / c8;
    const int wo = static_cast<int>(t % Wo); t /= Wo;
    const int ho = static_cast<int>(t % Ho);
    const int64_t b = t / Ho;
    const int h0 = 2 * ho - 1, w0 = 2 * wo - 1;
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int hh = h0 + dy;
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        const int ww = w0 + dx;
        if (ww < 0 || ww >= W) continue;
        const uint4 u = __ldg(x + ((b * H + hh) * W + ww) * c8 + cg);
        const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float lo, hi;
          if (BF16) {
            lo = __uint_as_float(wv[e] << 16);
            hi = __uint_as_float(wv[e] & 0xffff0000u);
          } else {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&wv[e]));
            lo = f.x;
            hi = f.y;
          }
          m[2 * e] = fmaxf(m[2 * e], lo);              // (NaN inputs are dropped by fmaxf; the framework propagates them)
          m[2 * e + 1] = fmaxf(m[2 * e + 1], hi);
        }
      }
    }
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (BF16) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(m[2 * e], m[2 * e + 1]);
        o[e] = *reinterpret_cast<const uint32_t*>(&h);
      } else {
        const __half2 h = __floats2half2_rn(m[2 * e], m[2 * e + 1]);
        o[e] = *reinterpret_cast<const uint32_t*>(&h);
      }
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace
}  // namespace acan

extern "C" int acan_maxpool3x3s2_nhwc(const void* x, void* y, int dtype, int B, int H, int W, int C, void* stream) {
  ACAN_CHECK_ARG(x && y, "acan_maxpool3x3s2_nhwc: null pointer");
  ACAN_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && (C % 8) == 0, "acan_maxpool3x3s2_nhwc: B=%d H=%d W=%d C=%d invalid (C %% 8 == 0)", B, H, W, C);
  ACAN_CHECK_ARG(dtype == 1 || dtype == 2, "acan_maxpool3x3s2_nhwc: dtype %d (1 = fp16, 2 = bf16)", dtype);
  ACAN_CHECK_ARG(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0, "acan_maxpool3x3s2_nhwc: 16-byte alignment");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1, c8 = C / 8;
  const int64_t total = static_cast<int64_t>(B) * Ho * Wo * c8;
  int64_t blocks = (total + 255) / 256;
  if (blocks > acan::kNumSMs * 32) blocks = acan::kNumSMs * 32;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == 2) acan::maxpool3x3s2_nhwc_kernel<true><<<static_cast<int>(blocks), 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<uint4*>(y), H, W, Ho, Wo, c8, total);
  else            acan::maxpool3x3s2_nhwc_kernel<false><<<static_cast<int>(blocks), 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<uint4*>(y), H, W, Ho, Wo, c8, total);
  return acan::launch_status("acan_maxpool3x3s2_nhwc");
}
