// optim.cu -- SGD (momentum, weight decay) over ONE flat fp32 parameter buffer in a single pass.
//
// The reference trains with torch.optim.SGD over three parameter groups (creator/optimizer.py:8-10, creator/params.py:14-30: backbone
// at lr, new weights at 10 lr, new biases at 20 lr).  The framework's multi-tensor SGD makes four passes (weight decay out of place,
// momentum scale, momentum add, parameter update: 11 x 370 MB of traffic for ResNet-101, 1.1 ms of a 15 ms step measured on B200) and the
// training driver then refreshes its 16-bit shadow weights in a fifth.  With parameters, gradients and momentum buffers bound to flat
// buffers (acan_b200/training.py) the whole update is one streaming kernel: read p, g, m; write p, m and the bf16 shadow -- 5.5 x 4 bytes
// per element, HBM-bound.  Arithmetic is torch's, operation for operation (torch/optim/sgd.py, nesterov = False, dampening = 0):
//     g' = g + wd * p;   m = mu * m + g'   (a zero-initialised m makes the first step m = g');   p = p - lr * m
// Hyper-parameters live in a small DEVICE table, so a scheduler changing lr needs no re-capture of a CUDA graph around this kernel.
#include "common.cuh"
#include <cuda_bf16.h>

namespace acan {
namespace {

constexpr int kSgdThreads = 256, kSgdChunk = 2048;     // elements per block: 2 float4 per thread

__device__ __forceinline__ void sgd_one(float& p, float g, float& m, float lr, float mu, float wd) {
  const float gw = fmaf(wd, p, g);
  m = __fadd_rn(__fmul_rn(mu, m), gw);
  p = fmaf(-lr, m, p);
}

__global__ void __launch_bounds__(kSgdThreads) sgd_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                                __nv_bfloat16* __restrict__ shadow, const signed char* __restrict__ chunk_group,
                                                                const long long* __restrict__ seg_end, const int* __restrict__ seg_group, int n_seg,
                                                                const float4* __restrict__ hyper, long long first_chunk, long long n) {
  const long long chunk = first_chunk + blockIdx.x;
  const long long base = chunk * kSgdChunk;
  const int cg = chunk_group[chunk];
  if (cg >= 0) {
    const float4 h = __ldg(hyper + cg);                 // (lr, momentum, weight_decay, -)
#pragma unroll
    for (int r = 0; r < kSgdChunk / (kSgdThreads * 4); ++r) {
      const long long i = base + (static_cast<long long>(r) * kSgdThreads + threadIdx.x) * 4;
      if (i + 3 < n) {
        float4 pv = *reinterpret_cast<const float4*>(p + i);
        const float4 gv = ldg_stream(reinterpret_cast<const float4*>(g + i));
        float4 mv = *reinterpret_cast<const float4*>(m + i);
        sgd_one(pv.x, gv.x, mv.x, h.x, h.y, h.z); sgd_one(pv.y, gv.y, mv.y, h.x, h.y, h.z);
        sgd_one(pv.z, gv.z, mv.z, h.x, h.y, h.z); sgd_one(pv.w, gv.w, mv.w, h.x, h.y, h.z);
        *reinterpret_cast<float4*>(p + i) = pv;
        *reinterpret_cast<float4*>(m + i) = mv;
        if (shadow) {
          const __nv_bfloat162 a = __floats2bfloat162_rn(pv.x, pv.y), b = __floats2bfloat162_rn(pv.z, pv.w);
          *reinterpret_cast<uint2*>(shadow + i) = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
        }
      } else {
        for (long long j = i; j < n && j < i + 4; ++j) {
          float pv = p[j], mv = m[j];
          sgd_one(pv, g[j], mv, h.x, h.y, h.z);
          p[j] = pv; m[j] = mv;
          if (shadow) shadow[j] = __float2bfloat16_rn(pv);
        }
      }
    }
  } else {                                               // the chunk spans parameter groups: look every element's segment up
    for (int r = threadIdx.x; r < kSgdChunk; r += kSgdThreads) {
      const long long j = base + r;
      if (j >= n) break;
      int lo = 0, hi = n_seg - 1;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(seg_end + mid) > j) hi = mid; else lo = mid + 1; }
      const float4 h = __ldg(hyper + __ldg(seg_group + lo));
      float pv = p[j], mv = m[j];
      sgd_one(pv, g[j], mv, h.x, h.y, h.z);
      p[j] = pv; m[j] = mv;
      if (shadow) shadow[j] = __float2bfloat16_rn(pv);
    }
  }
}


// ---------------------------------------------------------------- 16-bit gradients -> their fp32 slices of the flat buffer, one launch
// Under autocast the weight gradients of the convolutions come out of cuDNN in 16 bit, one tensor per layer; the optimizer (and the
// all-reduce) want them in fp32 in the flat buffer.  torch._foreach_copy_ with mixed dtypes falls back to one cast kernel per tensor
// (121 launches, 0.73 ms of a 13 ms ResNet-101 step for 550 MB of traffic).  Here up to kCastBatch (source, destination, count)
// triples travel in the kernel's parameter space and every block converts one kCastChunk-element piece of one tensor.
constexpr int kCastBatch = 96, kCastChunk = 4096, kCastThreads = 256;
struct CastBatch {
  const void* src[kCastBatch];
  float* dst[kCastBatch];
  unsigned first_chunk[kCastBatch + 1];     // prefix sums of ceil(count / kCastChunk)
  unsigned count[kCastBatch];
  int n;
};

template <bool BF16>
__device__ __forceinline__ float to_f32(unsigned short h) {
  if (BF16) return __uint_as_float(static_cast<uint32_t>(h) << 16);
  return __half2float(__ushort_as_half(h));
}

template <bool BF16>
__global__ void __launch_bounds__(kCastThreads) cast_multi_kernel(const __grid_constant__ CastBatch b) {
  int lo = 0, hi = b.n;                                   // entry e with first_chunk[e] <= blockIdx.x < first_chunk[e + 1]
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (b.first_chunk[mid] <= blockIdx.x) lo = mid; else hi = mid; }
  const unsigned short* src = static_cast<const unsigned short*>(b.src[lo]);
  float* dst = b.dst[lo];
  const unsigned n = b.count[lo];
  const unsigned base = (blockIdx.x - b.first_chunk[lo]) * kCastChunk;
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
#pragma unroll
  for (int r = 0; r < kCastChunk / (kCastThreads * 8); ++r) {
    const unsigned i = base + (r * kCastThreads + threadIdx.x) * 8;
    if (vec && i + 7 < n) {
      const uint4 u = ldg_nc_u4(reinterpret_cast<const uint4*>(src + i));
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      float f[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { f[2 * e] = to_f32<BF16>(static_cast<unsigned short>(w[e] & 0xffffu)); f[2 * e + 1] = to_f32<BF16>(static_cast<unsigned short>(w[e] >> 16)); }
      *reinterpret_cast<float4*>(dst + i) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(dst + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
    } else {
      for (unsigned j = i; j < n && j < i + 8; ++j) dst[j] = to_f32<BF16>(src[j]);
    }
  }
}
}  // namespace
}  // namespace acan

using namespace acan;

extern "C" int acan_sgd_flat_chunk(void) { return kSgdChunk; }

extern "C" int acan_sgd_flat(float* param, const float* grad, float* momentum_buf, void* shadow_bf16, const signed char* chunk_group,
                             const long long* seg_end, const int* seg_group, int n_seg, const float* hyper,
                             long long first_chunk, long long n_chunks, long long n, void* stream) {
  ACAN_CHECK_ARG(param && grad && momentum_buf && chunk_group && seg_end && seg_group && hyper, "acan_sgd_flat: null pointer");
  ACAN_CHECK_ARG(n > 0 && n_seg > 0 && first_chunk >= 0 && n_chunks > 0 && (first_chunk + n_chunks - 1) * kSgdChunk < n, "acan_sgd_flat: bad range");
  ACAN_CHECK_ARG(reinterpret_cast<uintptr_t>(param) % 16 == 0 && reinterpret_cast<uintptr_t>(grad) % 16 == 0 && reinterpret_cast<uintptr_t>(momentum_buf) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(shadow_bf16) % 8 == 0 && reinterpret_cast<uintptr_t>(hyper) % 16 == 0, "acan_sgd_flat: buffers must be 16-byte aligned");
  ACAN_CHECK_ARG(n_chunks < 0x7fffffffLL, "acan_sgd_flat: too many chunks for one launch");
  sgd_flat_kernel<<<static_cast<unsigned>(n_chunks), kSgdThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, momentum_buf, static_cast<__nv_bfloat16*>(shadow_bf16), chunk_group, seg_end, seg_group, n_seg,
      reinterpret_cast<const float4*>(hyper), first_chunk, n);
  return launch_status("acan_sgd_flat");
}

extern "C" int acan_cast_to_f32_multi(const void* const* src, float* const* dst, const long long* count, int n_tensors, int src_dtype, void* stream) {
  ACAN_CHECK_ARG(n_tensors >= 0 && (n_tensors == 0 || (src && dst && count)), "acan_cast_to_f32_multi: null table");
  ACAN_CHECK_ARG(src_dtype == ACAN_DT_F16 || src_dtype == ACAN_DT_BF16, "acan_cast_to_f32_multi: src_dtype %d (ACAN_DT_F16 or ACAN_DT_BF16)", src_dtype);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int t0 = 0; t0 < n_tensors; t0 += kCastBatch) {
    CastBatch b;
    b.n = 0;
    unsigned chunks = 0;
    for (int t = t0; t < n_tensors && b.n < kCastBatch; ++t) {
      ACAN_CHECK_ARG(count[t] >= 0 && count[t] < (1ll << 32) - kCastChunk, "acan_cast_to_f32_multi: tensor %d has %lld elements", t, count[t]);
      if (count[t] == 0) continue;
      ACAN_CHECK_ARG(src[t] && dst[t] && reinterpret_cast<uintptr_t>(src[t]) % 2 == 0 && reinterpret_cast<uintptr_t>(dst[t]) % 4 == 0,
                     "acan_cast_to_f32_multi: tensor %d: null or misaligned pointer", t);
      b.src[b.n] = src[t]; b.dst[b.n] = dst[t]; b.count[b.n] = static_cast<unsigned>(count[t]);
      b.first_chunk[b.n] = chunks;
      chunks += static_cast<unsigned>((count[t] + kCastChunk - 1) / kCastChunk);
      ++b.n;
    }
    if (b.n == 0) continue;
    b.first_chunk[b.n] = chunks;
    if (src_dtype == ACAN_DT_BF16) cast_multi_kernel<true><<<chunks, kCastThreads, 0, st>>>(b);
    else                           cast_multi_kernel<false><<<chunks, kCastThreads, 0, st>>>(b);
    if (int e = launch_status("cast_multi_kernel")) return e;
  }
  return 0;
}
