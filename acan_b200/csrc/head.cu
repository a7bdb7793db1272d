// head.cu -- ordinal / cross-entropy depth-bin head as single-pass HBM-streaming kernels.
//
// Replaces BaseClassificationModel_.decode_ord / soft|hard_ordinal_regression /
// soft|hard_cross_entropy / inference (reference model.py:40-99) and continuous2discrete
// (model.py:19-23).  The reference runs ~5 ATen passes for decode_ord (incl. a full-size exp(y)
// temporary) and ~10 more for the inference; here every variant reads its input exactly once:
//   prob   -> depth : (K+1)*HW*4  bytes / image   (324*HW at K = 80)
//   logits -> depth : (2K+1)*HW*4 bytes / image   (644*HW), decode_ord fused, prob never stored
//   logits -> prob  : 3K*HW*4     bytes / image   (960*HW)
// Layout [B, C, HW]: a thread owns 4 consecutive pixels (one 16 B load per channel), a warp a
// 512 B contiguous run per channel, and walks the C channels with the loads unrolled 8 deep so
// ~8 x 16 B per thread are in flight.  Grid = a multiple of the 148 SMs, grid-stride over quads.
//
// Bit-exactness of the hard indices: p > 0.5 is evaluated literally as e1/(e0+e1) > 0.5f with
// full-precision expf and IEEE division (no fast-math), arg-max keeps the first maximum and treats
// NaN as the maximum, as torch.argmax does.
#include "common.cuh"
#include <cmath>

namespace acan {
namespace {

constexpr int kThreads = 256;
constexpr int kBlocksPerSM = 8;

struct BinScale {      // d(k) = exp(k / (K-1) * c1 + c0), evaluated in fp32 like the reference's tensor expression
  float inv_km1_den;   // K - 1 as float (division, not multiplication by a reciprocal, to round as torch does)
  float c1;            // (float) log(d_max / d_min)
  float c0;            // (float) log(d_min)
};

__device__ __forceinline__ float bin_to_depth(float k, const BinScale& s) {
  return expf(k / s.inv_km1_den * s.c1 + s.c0);
}

__device__ __forceinline__ float ord_prob(float y0, float y1) {
  const float e0 = expf(y0), e1 = expf(y1);
  return e1 / (e0 + e1);
}

// ---------------------------------------------------------------- decode_ord fwd / bwd (vector path)
__global__ void __launch_bounds__(kThreads) decode_ord_fwd_kernel(const float4* __restrict__ y, float4* __restrict__ p,
                                                                   int K, int64_t hw4, int64_t total) {
  // total = B*K*hw4 quads; quad q -> (bk = q / hw4, x = q % hw4); logits rows 2*bk and 2*bk+1
  for (int64_t q = blockIdx.x * (int64_t)kThreads + threadIdx.x; q < total; q += (int64_t)gridDim.x * kThreads) {
    const int64_t bk = q / hw4, x = q - bk * hw4;
    const float4 a = ldg_stream(y + (2 * bk) * hw4 + x);
    const float4 b = ldg_stream(y + (2 * bk + 1) * hw4 + x);
    float4 r;
    r.x = ord_prob(a.x, b.x); r.y = ord_prob(a.y, b.y); r.z = ord_prob(a.z, b.z); r.w = ord_prob(a.w, b.w);
    p[q] = r;   // prob is re-read by the loss / inference right after: leave it cacheable
  }
}

__global__ void __launch_bounds__(kThreads) decode_ord_fwd_scalar(const float* __restrict__ y, float* __restrict__ p,
                                                                   int64_t hw, int64_t total) {
  for (int64_t q = blockIdx.x * (int64_t)kThreads + threadIdx.x; q < total; q += (int64_t)gridDim.x * kThreads) {
    const int64_t bk = q / hw, x = q - bk * hw;
    p[q] = ord_prob(y[(2 * bk) * hw + x], y[(2 * bk + 1) * hw + x]);
  }
}

__global__ void __launch_bounds__(kThreads) decode_ord_bwd_kernel(const float* __restrict__ p, const float* __restrict__ g,
                                                                   float* __restrict__ gy, int64_t hw, int64_t total) {
  for (int64_t q = blockIdx.x * (int64_t)kThreads + threadIdx.x; q < total; q += (int64_t)gridDim.x * kThreads) {
    const int64_t bk = q / hw, x = q - bk * hw;
    const float pv = p[q];
    const float d = g[q] * pv * (1.0f - pv);
    gy[(2 * bk) * hw + x] = -d;
    gy[(2 * bk + 1) * hw + x] = d;
  }
}

// ---------------------------------------------------------------- inference
// per-pixel accumulator for each head kind; feed() is called once per bin in increasing k.
template <int KIND> struct Acc;

template <> struct Acc<ACAN_HEAD_OR_SOFT> {
  float s = 0.f;
  __device__ __forceinline__ void feed(float p, int) { s += p; }
  __device__ __forceinline__ float finish(const BinScale& bs, int& idx) const {
    const float i = floorf(s), f = s - i;
    const float d0 = bin_to_depth(i, bs), d1 = bin_to_depth(i + 1.f, bs), d2 = bin_to_depth(i + 2.f, bs);
    idx = (int)i;
    return (d0 + d1) / 2.f * (1.f - f) + (d1 + d2) / 2.f * f;
  }
};
template <> struct Acc<ACAN_HEAD_OR_HARD> {
  int n = 0;
  __device__ __forceinline__ void feed(float p, int) { n += (p > 0.5f) ? 1 : 0; }
  __device__ __forceinline__ float finish(const BinScale& bs, int& idx) const {
    idx = n;
    return (bin_to_depth((float)n, bs) + bin_to_depth((float)n + 1.f, bs)) / 2.f;
  }
};
template <> struct Acc<ACAN_HEAD_CE_HARD> {
  float best = 0.f; int arg = -1;
  __device__ __forceinline__ void feed(float v, int k) {
    // first maximum wins; NaN is the maximum and sticks (torch.argmax semantics)
    if (arg < 0 || v > best || (v != v && best == best)) { best = v; arg = k; }
  }
  __device__ __forceinline__ float finish(const BinScale& bs, int& idx) const {
    idx = arg;
    return bin_to_depth((float)arg, bs);
  }
};
template <> struct Acc<ACAN_HEAD_CE_SOFT> {   // online softmax expectation of the log-depth bin centres
  float m = -INFINITY, z = 0.f, a = 0.f; int arg = 0;
  // raise the running maximum to gm (>= every score of the coming group): ONE rescale per group of channels
  __device__ __forceinline__ void lift(float gm, int garg) {
    if (gm > m) { const float c = ex2_approx((m - gm) * 1.4426950408889634f); z *= c; a *= c; m = gm; arg = garg; }
  }
  __device__ __forceinline__ void feed_w(float v, int, float w) {
    const float e = ex2_approx((v - m) * 1.4426950408889634f);
    z += e; a = fmaf(e, w, a);
  }
  __device__ __forceinline__ float finish(const BinScale&, int& idx) const {
    idx = arg;
    return expf(a / z);
  }
};

template <int KIND, bool FROM_LOGITS>
__global__ void __launch_bounds__(kThreads) head_kernel(const float4* __restrict__ in, float4* __restrict__ depth,
                                                        int4* __restrict__ index, int K, int64_t hw4, int64_t total, BinScale bs) {
  constexpr int U = 8;
  const int C = FROM_LOGITS ? 2 * K : K;
  for (int64_t q = blockIdx.x * (int64_t)kThreads + threadIdx.x; q < total; q += (int64_t)gridDim.x * kThreads) {
    const int64_t b = q / hw4, x = q - b * hw4;
    const float4* src = in + b * (int64_t)C * hw4 + x;
    Acc<KIND> acc[4];
    for (int c0 = 0; c0 < C; c0 += U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (c0 + u < C) v[u] = ldg_stream(src + (int64_t)(c0 + u) * hw4);
      if constexpr (FROM_LOGITS) {
#pragma unroll
        for (int u = 0; u < U; u += 2) {
          if (c0 + u < C) {
            const int k = (c0 + u) >> 1;
            acc[0].feed(ord_prob(v[u].x, v[u + 1].x), k);
            acc[1].feed(ord_prob(v[u].y, v[u + 1].y), k);
            acc[2].feed(ord_prob(v[u].z, v[u + 1].z), k);
            acc[3].feed(ord_prob(v[u].w, v[u + 1].w), k);
          }
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (c0 + u < C) {
            const int k = c0 + u;
            if constexpr (KIND == ACAN_HEAD_CE_SOFT) {
              if (u == 0) {   // group maximum first (first index on ties), one rescale, then U independent ex2
                float gx = v[0].x, gy = v[0].y, gz = v[0].z, gw = v[0].w; int ax = c0, ay = c0, az = c0, aw = c0;
#pragma unroll
                for (int t = 1; t < U; ++t)
                  if (c0 + t < C) {
                    if (v[t].x > gx) { gx = v[t].x; ax = c0 + t; }
                    if (v[t].y > gy) { gy = v[t].y; ay = c0 + t; }
                    if (v[t].z > gz) { gz = v[t].z; az = c0 + t; }
                    if (v[t].w > gw) { gw = v[t].w; aw = c0 + t; }
                  }
                acc[0].lift(gx, ax); acc[1].lift(gy, ay); acc[2].lift(gz, az); acc[3].lift(gw, aw);
              }
              // weight_k = k * log(dmax/dmin) / (K-1) + log(dmin)   (model.py:54-55, fp32 tensor expression)
              const float w = (float)k * bs.c1 / bs.inv_km1_den + bs.c0;
              acc[0].feed_w(v[u].x, k, w); acc[1].feed_w(v[u].y, k, w);
              acc[2].feed_w(v[u].z, k, w); acc[3].feed_w(v[u].w, k, w);
            } else {
              acc[0].feed(v[u].x, k); acc[1].feed(v[u].y, k); acc[2].feed(v[u].z, k); acc[3].feed(v[u].w, k);
            }
          }
        }
      }
    }
    float4 d; int4 i;
    d.x = acc[0].finish(bs, i.x); d.y = acc[1].finish(bs, i.y); d.z = acc[2].finish(bs, i.z); d.w = acc[3].finish(bs, i.w);
    depth[q] = d;
    if (index != nullptr) index[q] = i;
  }
}

// scalar twin for HW % 4 != 0 (never the case for the stride-8 network, kept for the ABI's generality)
template <int KIND, bool FROM_LOGITS>
__global__ void __launch_bounds__(kThreads) head_kernel_scalar(const float* __restrict__ in, float* __restrict__ depth,
                                                               int* __restrict__ index, int K, int64_t hw, int64_t total, BinScale bs) {
  const int C = FROM_LOGITS ? 2 * K : K;
  for (int64_t q = blockIdx.x * (int64_t)kThreads + threadIdx.x; q < total; q += (int64_t)gridDim.x * kThreads) {
    const int64_t b = q / hw, x = q - b * hw;
    const float* src = in + b * (int64_t)C * hw + x;
    Acc<KIND> acc;
    for (int k = 0; k < K; ++k) {
      if constexpr (FROM_LOGITS) {
        acc.feed(ord_prob(src[(int64_t)(2 * k) * hw], src[(int64_t)(2 * k + 1) * hw]), k);
      } else if constexpr (KIND == ACAN_HEAD_CE_SOFT) {
        const float v = src[(int64_t)k * hw];
        acc.lift(v, k);
        acc.feed_w(v, k, (float)k * bs.c1 / bs.inv_km1_den + bs.c0);
      } else {
        acc.feed(src[(int64_t)k * hw], k);
      }
    }
    int i;
    depth[q] = acc.finish(bs, i);
    if (index != nullptr) index[q] = i;
  }
}

__global__ void __launch_bounds__(kThreads) c2d_kernel(const float* __restrict__ d, float* __restrict__ k, int64_t n,
                                                        float dmin, float dmax, float inv_log_range, float km1) {
  for (int64_t q = blockIdx.x * (int64_t)kThreads + threadIdx.x; q < n; q += (int64_t)gridDim.x * kThreads) {
    const float v = d[q];
    // torch: round(log(depth / d_min) / np.log(d_max/d_min) * (n_c - 1)); rintf = round-half-even like torch.round
    const float r = rintf(logf(v / dmin) / inv_log_range * km1);
    k[q] = (v > dmin && v < dmax) ? r : 0.f;
  }
}

// ---------------------------------------------------------------- fused classifier tail
// x8 bilinear upsampling (align_corners=True, the arithmetic of ATen's upsample_bilinear2d: src = dst*(in-1)/(out-1),
// fp32) + decode_ord + inference in one kernel that reads the stride-8 logits z [B,h,w,C] (C = 2K for OR, K for CE; 0.9 MB
// per NYU image) instead of the upsampled [B,C,8h,8w] tensor (58 MB per image written, converted, read twice).
// A CTA produces an 8 x 32 output tile; the <= 3 x 6 source pixels it touches are staged in shared memory with the
// channel dimension padded by 2 floats so that the 4 bilinear taps of a warp hit distinct banks.
constexpr int kTailRows = 8, kTailCols = 32, kTailSrcR = 3, kTailSrcC = 6;

template <int KIND>
__global__ void __launch_bounds__(kTailRows * kTailCols)
tail_kernel(const float* __restrict__ z, float* __restrict__ depth, float* __restrict__ prob, int* __restrict__ index,
            int h, int w, int K, BinScale bs) {
  extern __shared__ float patch[];                       // [kTailSrcR][kTailSrcC][C + 2]
  constexpr bool OR = (KIND == ACAN_HEAD_OR_SOFT || KIND == ACAN_HEAD_OR_HARD);
  const int C = OR ? 2 * K : K, CP = C + 2;
  const int H = 8 * h, W = 8 * w;
  const int b = blockIdx.z, Y0 = blockIdx.y * kTailRows, X0 = blockIdx.x * kTailCols;
  const float rh = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float rw = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const int r0 = (int)(rh * (float)Y0), c0 = (int)(rw * (float)X0);
  const float* zb = z + (int64_t)b * h * w * C;
  for (int i = threadIdx.x; i < kTailSrcR * kTailSrcC * C; i += kTailRows * kTailCols) {
    const int c = i % C, rc = i / C, cc = rc % kTailSrcC, rr = rc / kTailSrcC;
    const int sr = min(r0 + rr, h - 1), sc = min(c0 + cc, w - 1);
    patch[(rr * kTailSrcC + cc) * CP + c] = __ldg(zb + ((int64_t)sr * w + sc) * C + c);
  }
  __syncthreads();
  const int Y = Y0 + (threadIdx.x >> 5), X = X0 + (threadIdx.x & 31);
  if (X >= W || Y >= H) return;
  const float h1r = rh * (float)Y, w1r = rw * (float)X;
  const int h1 = (int)h1r, w1 = (int)w1r;
  const int h1p = (h1 < h - 1) ? 1 : 0, w1p = (w1 < w - 1) ? 1 : 0;
  const float h1l = h1r - (float)h1, h0l = 1.f - h1l, w1l = w1r - (float)w1, w0l = 1.f - w1l;
  const float* p00 = patch + ((h1 - r0) * kTailSrcC + (w1 - c0)) * CP;
  const float* p01 = p00 + w1p * CP;
  const float* p10 = p00 + h1p * kTailSrcC * CP;
  const float* p11 = p10 + w1p * CP;
  const int64_t HW = (int64_t)H * W, pix = (int64_t)Y * W + X;
  Acc<KIND> acc;
  for (int k = 0; k < K; ++k) {
    if constexpr (OR) {
      const float2 a = *reinterpret_cast<const float2*>(p00 + 2 * k), bq = *reinterpret_cast<const float2*>(p01 + 2 * k);
      const float2 c = *reinterpret_cast<const float2*>(p10 + 2 * k), d = *reinterpret_cast<const float2*>(p11 + 2 * k);
      const float y0 = h0l * (w0l * a.x + w1l * bq.x) + h1l * (w0l * c.x + w1l * d.x);
      const float y1 = h0l * (w0l * a.y + w1l * bq.y) + h1l * (w0l * c.y + w1l * d.y);
      // soft inference with no probability output: p = 1 / (1 + exp(y0 - y1)) with the fast intrinsics (~1e-6 relative; the depth
      // is a float output with a 1e-3 bar).  Hard inference and the materialised 'y' keep the reference's literal e1 / (e0 + e1):
      // the bin index is bit-exact and p > 0.5 is not equivalent to y1 > y0 at rounding ties.
      float pk;
      if (KIND == ACAN_HEAD_OR_SOFT && prob == nullptr) pk = __fdividef(1.0f, 1.0f + __expf(y0 - y1));
      else pk = ord_prob(y0, y1);
      acc.feed(pk, k);
      if (prob != nullptr) prob[((int64_t)b * K + k) * HW + pix] = pk;
    } else {
      const float v = h0l * (w0l * p00[k] + w1l * p01[k]) + h1l * (w0l * p10[k] + w1l * p11[k]);
      if constexpr (KIND == ACAN_HEAD_CE_SOFT) {
        acc.lift(v, k);
        acc.feed_w(v, k, (float)k * bs.c1 / bs.inv_km1_den + bs.c0);
      } else {
        acc.feed(v, k);
      }
      if (prob != nullptr) prob[((int64_t)b * K + k) * HW + pix] = v;
    }
  }
  int idx;
  depth[(int64_t)b * HW + pix] = acc.finish(bs, idx);
  if (index != nullptr) index[(int64_t)b * HW + pix] = idx;
}

// Soft ordinal inference alone (depth only: no probability / index output) -- the eval call pattern of the benched step.  The first
// version (tail_kernel: one thread per output pixel, four float2 shared-memory loads and twelve FMAs per (pixel, bin)) ran at 155 us for
// configs[1] with three pipes -- LSU, FMA, MUFU -- at ~54 us each and little overlap.  Here only the MUFU work is left per output:
//   * p = 1 / (1 + exp(y0 - y1)) needs the DIFFERENCE of the pair only, and interpolation is linear: the patch holds
//     D = (z0 - z1) log2(e) per (source pixel, bin) -- one float instead of two;
//   * a thread owns one output COLUMN of an 8-row band: per bin it interpolates horizontally ONCE per source row (<= 3 rows: 6 loads)
//     and then walks the 8 output rows with one FMA each (v = R + hl (R' - R)), one ex2, one rcp, one add.
// Per output: 0.75 loads, ~2 FMA-pipe ops, 2 MUFU ops: MUFU-bound (16 / clock / SM), ~2.4x faster than the per-pixel kernel.
// Rounding differs from ATen's upsample formula at the 1e-7 level (same as the fast sigmoid it already used): float output, 1e-5 test bar;
// hard inference and the materialised 'y' keep tail_kernel and the reference's literal arithmetic (bit-exact indices).
constexpr int kSoftCols = 128, kSoftRows = 8, kSoftSrcR = 3, kSoftSrcC = 20;

template <int SW>   // rows [0, SW) of the band interpolate between source rows 0 / 1 of the patch, rows [SW, 8) between rows 1 / 2
__device__ __forceinline__ void soft_band(const float* __restrict__ pa, const float* __restrict__ pb, int rowp, int KP, int K, float wl, const float (&hl)[kSoftRows],
                                          float (&acc)[kSoftRows]) {
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    const float a0 = pa[k], b0 = pb[k];
    const float a1 = pa[rowp + k], b1 = pb[rowp + k];
    const float r0 = fmaf(wl, b0 - a0, a0), r1 = fmaf(wl, b1 - a1, a1);
    float r2 = r1;
    if (SW < kSoftRows) { const float a2 = pa[2 * rowp + k], b2 = pb[2 * rowp + k]; r2 = fmaf(wl, b2 - a2, a2); }
    const float d01 = r1 - r0, d12 = r2 - r1;
#pragma unroll
    for (int j = 0; j < kSoftRows; ++j) {
      const float v = (j < SW) ? fmaf(hl[j], d01, r0) : fmaf(hl[j], d12, r1);
      acc[j] += __fdividef(1.0f, 1.0f + ex2_approx(v));
    }
  }
}

__global__ void __launch_bounds__(kSoftCols)
tail_soft_kernel(const float* __restrict__ z, float* __restrict__ depth, int h, int w, int K, BinScale bs) {
  extern __shared__ float patch[];                       // [kSoftSrcR][kSoftSrcC][K + 1]: D = (z0 - z1) * log2(e)
  const int KP = K + 1, C = 2 * K;
  const int H = 8 * h, W = 8 * w;
  const int b = blockIdx.z, Y0 = blockIdx.y * kSoftRows, X0 = blockIdx.x * kSoftCols;
  const float rh = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float rw = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const int r0 = (int)(rh * (float)Y0), c0 = (int)(rw * (float)X0);
  const float* zb = z + (int64_t)b * h * w * C;
  for (int i = threadIdx.x; i < kSoftSrcR * kSoftSrcC * K; i += kSoftCols) {
    const int k = i % K, rc = i / K, cc = rc % kSoftSrcC, rr = rc / kSoftSrcC;
    const int sr = min(r0 + rr, h - 1), sc = min(c0 + cc, w - 1);
    const float2 v = __ldg(reinterpret_cast<const float2*>(zb + ((int64_t)sr * w + sc) * C) + k);
    patch[(rr * kSoftSrcC + cc) * KP + k] = (v.x - v.y) * 1.4426950408889634f;
  }
  __syncthreads();
  const int X = X0 + threadIdx.x;
  if (X >= W) return;
  const float w1r = rw * (float)X;
  const int w1 = (int)w1r;
  const int w1p = (w1 < w - 1) ? 1 : 0;
  const float wl = w1r - (float)w1;
  float hl[kSoftRows];
  int sw = 0;                                             // block-uniform: number of rows of the band whose upper source row is r0
#pragma unroll
  for (int j = 0; j < kSoftRows; ++j) {
    const float h1r = rh * (float)min(Y0 + j, H - 1);
    const int h1 = (int)h1r;
    hl[j] = h1r - (float)h1;
    sw += (h1 == r0) ? 1 : 0;
  }
  const float* pa = patch + (w1 - c0) * KP;
  const float* pb = pa + w1p * KP;
  const int rowp = kSoftSrcC * KP;
  float acc[kSoftRows];
#pragma unroll
  for (int j = 0; j < kSoftRows; ++j) acc[j] = 0.f;
  switch (sw) {
    case 1: soft_band<1>(pa, pb, rowp, KP, K, wl, hl, acc); break;
    case 2: soft_band<2>(pa, pb, rowp, KP, K, wl, hl, acc); break;
    case 3: soft_band<3>(pa, pb, rowp, KP, K, wl, hl, acc); break;
    case 4: soft_band<4>(pa, pb, rowp, KP, K, wl, hl, acc); break;
    case 5: soft_band<5>(pa, pb, rowp, KP, K, wl, hl, acc); break;
    case 6: soft_band<6>(pa, pb, rowp, KP, K, wl, hl, acc); break;
    case 7: soft_band<7>(pa, pb, rowp, KP, K, wl, hl, acc); break;
    default: soft_band<8>(pa, pb, rowp, KP, K, wl, hl, acc); break;
  }
  const int64_t HW = (int64_t)H * W;
#pragma unroll
  for (int j = 0; j < kSoftRows; ++j) {
    if (Y0 + j < H) {
      Acc<ACAN_HEAD_OR_SOFT> a;
      a.s = acc[j];
      int idx;
      depth[(int64_t)b * HW + (int64_t)(Y0 + j) * W + X] = a.finish(bs, idx);
    }
  }
}

template <int KIND>
int launch_tail(const float* z, float* depth, float* prob, int32_t* index, int B, int h, int w, int K, BinScale bs, cudaStream_t st) {
  if constexpr (KIND == ACAN_HEAD_OR_SOFT) {
    // the band of 8 output rows starting at Y0 = 8 by touches source rows floor(rh Y0) .. +2 and a 128-column tile at most 18 source columns
    if (prob == nullptr && index == nullptr && (size_t)kSoftSrcR * kSoftSrcC * (K + 1) * sizeof(float) <= 48 * 1024) {
      const dim3 grid((8 * w + kSoftCols - 1) / kSoftCols, h, B);
      tail_soft_kernel<<<grid, kSoftCols, (size_t)kSoftSrcR * kSoftSrcC * (K + 1) * sizeof(float), st>>>(z, depth, h, w, K, bs);
      return launch_status("acan_tail_fwd (soft)");
    }
  }
  const int C = (KIND == ACAN_HEAD_OR_SOFT || KIND == ACAN_HEAD_OR_HARD) ? 2 * K : K;
  const size_t smem = (size_t)kTailSrcR * kTailSrcC * (C + 2) * sizeof(float);
  const dim3 grid((8 * w + kTailCols - 1) / kTailCols, h, B);     // 8 output rows per source row
  tail_kernel<KIND><<<grid, kTailRows * kTailCols, smem, st>>>(z, depth, prob, index, h, w, K, bs);
  return launch_status("acan_tail_fwd");
}

inline int grid_for(int64_t work_items) {
  const int64_t want = (work_items + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)kNumSMs * kBlocksPerSM;
  return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

template <int KIND, bool FL>
int launch_head(const float* in, float* depth, int32_t* index, int B, int K, int64_t HW, BinScale bs, cudaStream_t st) {
  if (HW % 4 == 0 && (reinterpret_cast<uintptr_t>(in) % 16 == 0) && (reinterpret_cast<uintptr_t>(depth) % 16 == 0) &&
      (index == nullptr || reinterpret_cast<uintptr_t>(index) % 16 == 0)) {
    const int64_t hw4 = HW / 4, total = (int64_t)B * hw4;
    head_kernel<KIND, FL><<<grid_for(total), kThreads, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(depth),
                                                                reinterpret_cast<int4*>(index), K, hw4, total, bs);
  } else {
    const int64_t total = (int64_t)B * HW;
    head_kernel_scalar<KIND, FL><<<grid_for(total), kThreads, 0, st>>>(in, depth, index, K, HW, total, bs);
  }
  return launch_status("acan_head_fwd");
}

}  // namespace
}  // namespace acan

using namespace acan;

extern "C" int acan_decode_ord_fwd(const float* logits, float* prob, int B, int K, int64_t HW, void* stream) {
  ACAN_CHECK_ARG(logits && prob, "acan_decode_ord_fwd: null pointer");
  ACAN_CHECK_ARG(B > 0 && K > 0 && HW > 0, "acan_decode_ord_fwd: B=%d K=%d HW=%lld must be positive", B, K, (long long)HW);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (HW % 4 == 0 && reinterpret_cast<uintptr_t>(logits) % 16 == 0 && reinterpret_cast<uintptr_t>(prob) % 16 == 0) {
    const int64_t hw4 = HW / 4, total = (int64_t)B * K * hw4;
    decode_ord_fwd_kernel<<<grid_for(total), kThreads, 0, st>>>(reinterpret_cast<const float4*>(logits), reinterpret_cast<float4*>(prob), K, hw4, total);
  } else {
    const int64_t total = (int64_t)B * K * HW;
    decode_ord_fwd_scalar<<<grid_for(total), kThreads, 0, st>>>(logits, prob, HW, total);
  }
  return launch_status("acan_decode_ord_fwd");
}

extern "C" int acan_decode_ord_bwd(const float* prob, const float* grad_prob, float* grad_logits, int B, int K, int64_t HW, void* stream) {
  ACAN_CHECK_ARG(prob && grad_prob && grad_logits, "acan_decode_ord_bwd: null pointer");
  ACAN_CHECK_ARG(B > 0 && K > 0 && HW > 0, "acan_decode_ord_bwd: sizes must be positive");
  const int64_t total = (int64_t)B * K * HW;
  decode_ord_bwd_kernel<<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(prob, grad_prob, grad_logits, HW, total);
  return launch_status("acan_decode_ord_bwd");
}

extern "C" int acan_head_fwd(const float* in, float* depth, int32_t* index, int B, int K, int64_t HW,
                             double d_min, double d_max, int kind, int from_logits, void* stream) {
  ACAN_CHECK_ARG(in && depth, "acan_head_fwd: null pointer");
  ACAN_CHECK_ARG(B > 0 && K > 1 && HW > 0, "acan_head_fwd: B=%d K=%d HW=%lld invalid", B, K, (long long)HW);
  ACAN_CHECK_ARG(d_min > 0 && d_max > d_min, "acan_head_fwd: need 0 < d_min < d_max");
  ACAN_CHECK_ARG(kind >= 0 && kind <= 3, "acan_head_fwd: kind %d unknown", kind);
  ACAN_CHECK_ARG(!(from_logits && kind >= ACAN_HEAD_CE_SOFT), "acan_head_fwd: from_logits applies to the OR kinds only");
  BinScale bs{(float)(K - 1), (float)std::log(d_max / d_min), (float)std::log(d_min)};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(PROF_HEAD, st);
  switch (kind * 2 + (from_logits ? 1 : 0)) {
    case ACAN_HEAD_OR_SOFT * 2 + 0: return launch_head<ACAN_HEAD_OR_SOFT, false>(in, depth, index, B, K, HW, bs, st);
    case ACAN_HEAD_OR_SOFT * 2 + 1: return launch_head<ACAN_HEAD_OR_SOFT, true>(in, depth, index, B, K, HW, bs, st);
    case ACAN_HEAD_OR_HARD * 2 + 0: return launch_head<ACAN_HEAD_OR_HARD, false>(in, depth, index, B, K, HW, bs, st);
    case ACAN_HEAD_OR_HARD * 2 + 1: return launch_head<ACAN_HEAD_OR_HARD, true>(in, depth, index, B, K, HW, bs, st);
    case ACAN_HEAD_CE_SOFT * 2 + 0: return launch_head<ACAN_HEAD_CE_SOFT, false>(in, depth, index, B, K, HW, bs, st);
    default:                        return launch_head<ACAN_HEAD_CE_HARD, false>(in, depth, index, B, K, HW, bs, st);
  }
}

extern "C" int acan_continuous2discrete(const float* depth, float* bins, int64_t n, double d_min, double d_max, int K, void* stream) {
  ACAN_CHECK_ARG(depth && bins && n > 0 && K > 1 && d_min > 0 && d_max > d_min, "acan_continuous2discrete: bad argument");
  c2d_kernel<<<grid_for(n), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(depth, bins, n, (float)d_min, (float)d_max,
                                                                              (float)std::log(d_max / d_min), (float)(K - 1));
  return launch_status("acan_continuous2discrete");
}

extern "C" int acan_tail_fwd(const float* z_nhwc, float* depth, float* prob, int32_t* index, int B, int h, int w, int K,
                             double d_min, double d_max, int kind, void* stream) {
  ACAN_CHECK_ARG(z_nhwc && depth, "acan_tail_fwd: null pointer");
  ACAN_CHECK_ARG(B > 0 && h > 0 && w > 0 && K > 1 && K <= 2048, "acan_tail_fwd: B=%d h=%d w=%d K=%d invalid", B, h, w, K);
  ACAN_CHECK_ARG(d_min > 0 && d_max > d_min, "acan_tail_fwd: need 0 < d_min < d_max");
  ACAN_CHECK_ARG(kind >= 0 && kind <= 3, "acan_tail_fwd: kind %d unknown", kind);
  ACAN_CHECK_ARG(B <= 65535 && h <= 65535, "acan_tail_fwd: grid too large");
  BinScale bs{(float)(K - 1), (float)std::log(d_max / d_min), (float)std::log(d_min)};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope ps(PROF_HEAD, st);
  switch (kind) {
    case ACAN_HEAD_OR_SOFT: return launch_tail<ACAN_HEAD_OR_SOFT>(z_nhwc, depth, prob, index, B, h, w, K, bs, st);
    case ACAN_HEAD_OR_HARD: return launch_tail<ACAN_HEAD_OR_HARD>(z_nhwc, depth, prob, index, B, h, w, K, bs, st);
    case ACAN_HEAD_CE_SOFT: return launch_tail<ACAN_HEAD_CE_SOFT>(z_nhwc, depth, prob, index, B, h, w, K, bs, st);
    default:                return launch_tail<ACAN_HEAD_CE_HARD>(z_nhwc, depth, prob, index, B, h, w, K, bs, st);
  }
}

// ---------------------------------------------------------------- ordinal-regression loss
// OrdinalRegression2d with the reference's live configuration (losses.py:103-158: reduction='sum', no class weights,
// ignore_index) fused with decode_ord (model.py:40-45): per valid pixel
//     loss = - sum_{k <  l} log clamp(p_k, 1e-8, 1e8) - sum_{k >= l} log clamp(1 - p_k, 1e-8, 1e8),   p_k = e1 / (e0 + e1)
// averaged over the pixels whose label != ignore_index.  Forward reads the pair logits once (640*HW B / image at K = 80)
// instead of decode_ord + ~10 elementwise passes over [B,K,H,W]; backward re-reads them and writes d/dlogits directly
// (d/dy1 = -(1-p) below the label, +p at / above it, inside the clamp windows; d/dy0 = -d/dy1), so the probability
// tensor and its gradient never exist.
namespace acan {
namespace {

__device__ __forceinline__ float ord_term(float p, bool below) {
  const float q = below ? p : 1.0f - p;
  return -logf(fminf(fmaxf(q, 1e-8f), 1e8f));
}

__global__ void __launch_bounds__(kThreads) ord_loss_fwd_kernel(const float4* __restrict__ y, const int64_t* __restrict__ label,
                                                                 float2* __restrict__ partial, int K, int64_t hw4, int64_t total, int ignore_index) {
  __shared__ float red_s[kThreads / 32], red_c[kThreads / 32];
  float sum = 0.f, cnt = 0.f;
  for (int64_t q = blockIdx.x * (int64_t)kThreads + threadIdx.x; q < total; q += (int64_t)gridDim.x * kThreads) {
    const int64_t b = q / hw4, x = q - b * hw4;
    const float4* src = y + b * (int64_t)(2 * K) * hw4 + x;
    const longlong2 la = *reinterpret_cast<const longlong2*>(label + (b * hw4 + x) * 4);
    const longlong2 lb = *reinterpret_cast<const longlong2*>(label + (b * hw4 + x) * 4 + 2);
    const int l[4] = {(int)la.x, (int)la.y, (int)lb.x, (int)lb.y};
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < K; k += 4) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (k + (u >> 1) < K) v[u] = ldg_stream(src + (int64_t)(2 * k + u) * hw4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (k + u < K) {
          const int kk = k + u;
          acc[0] += ord_term(ord_prob(v[2 * u].x, v[2 * u + 1].x), kk < l[0]);
          acc[1] += ord_term(ord_prob(v[2 * u].y, v[2 * u + 1].y), kk < l[1]);
          acc[2] += ord_term(ord_prob(v[2 * u].z, v[2 * u + 1].z), kk < l[2]);
          acc[3] += ord_term(ord_prob(v[2 * u].w, v[2 * u + 1].w), kk < l[3]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (l[i] != ignore_index) { sum += acc[i]; cnt += 1.f; }
  }
  sum = warp_sum(sum); cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) { red_s[threadIdx.x >> 5] = sum; red_c[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f, c = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) { s += red_s[i]; c += red_c[i]; }
    partial[blockIdx.x] = make_float2(s, c);
  }
}

__global__ void ord_loss_final_kernel(const float2* __restrict__ partial, int n, float* __restrict__ loss, float* __restrict__ count) {
  // one warp, fixed order
  float s = 0.f, c = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) { s += partial[i].x; c += partial[i].y; }
  s = warp_sum(s); c = warp_sum(c);
  if (threadIdx.x == 0) { *loss = s / c; *count = c; }      // mean over an empty set is NaN, as torch's .mean() gives
}

__global__ void __launch_bounds__(kThreads) ord_loss_bwd_kernel(const float4* __restrict__ y, const int64_t* __restrict__ label,
                                                                 const float* __restrict__ count, const float* __restrict__ grad_out,
                                                                 float4* __restrict__ gy, int K, int64_t hw4, int64_t total, int ignore_index) {
  const float g = __ldg(grad_out) / __ldg(count);
  // total = B*K*hw4 quads of (pair k, 4 pixels)
  for (int64_t q = blockIdx.x * (int64_t)kThreads + threadIdx.x; q < total; q += (int64_t)gridDim.x * kThreads) {
    const int64_t bk = q / hw4, x = q - bk * hw4;
    const int64_t b = bk / K;
    const int k = (int)(bk - b * K);
    const float4 a = ldg_stream(y + (2 * bk) * hw4 + x), c = ldg_stream(y + (2 * bk + 1) * hw4 + x);
    const longlong2 la = *reinterpret_cast<const longlong2*>(label + (b * hw4 + x) * 4);
    const longlong2 lb = *reinterpret_cast<const longlong2*>(label + (b * hw4 + x) * 4 + 2);
    auto one = [&](float y0, float y1, int l) {
      if (l == ignore_index) return 0.f;
      const float p = ord_prob(y0, y1);
      if (k < l) return (p > 1e-8f && p < 1e8f) ? -(1.0f - p) * g : 0.f;            // d/dy1 of -log p
      const float qv = 1.0f - p;
      return (qv > 1e-8f && qv < 1e8f) ? p * g : 0.f;                               // d/dy1 of -log (1 - p)
    };
    float4 d;
    d.x = one(a.x, c.x, (int)la.x); d.y = one(a.y, c.y, (int)la.y); d.z = one(a.z, c.z, (int)lb.x); d.w = one(a.w, c.w, (int)lb.y);
    gy[(2 * bk + 1) * hw4 + x] = d;
    gy[(2 * bk) * hw4 + x] = make_float4(-d.x, -d.y, -d.z, -d.w);
  }
}

}  // namespace
}  // namespace acan

// ---------------------------------------------------------------- fused classifier tail + ordinal loss (training; rows f-1 x f-2)
// loss = OrdinalRegression2d(decode_ord(upsample_x8(z)), label) straight from the stride-8 pair logits z [B,h,w,2K]
// (model.py:123-125, 228-230, 40-45; losses.py:103-158).  The full-resolution logits [B,2K,H,W] (58 MB / NYU image), their
// gradient and both upsampling passes of the unfused step never touch HBM: forward reads 0.9 MB / image + the labels,
// backward writes the 0.9 MB gradient of z.
//   forward : tail_kernel's tiling (8 x 32 output pixels per CTA, <= 3 x 6 source pixels staged in shared memory).
//   backward: dL/dz[sr,sc,2k+1] = sum over the <= 18 x 18 output pixels whose bilinear taps touch (sr,sc) of
//             wy wx dL/dy1(Y,X,k); a thread owns one (source pixel, k), recomputes the interpolated pair from its 3 x 3
//             source neighbourhood held in registers and gathers in a fixed order (deterministic, no atomics);
//             dL/dz[..,2k] = -dL/dz[..,2k+1].
namespace acan {
namespace {

__global__ void __launch_bounds__(kTailRows * kTailCols)
tail_loss_fwd_kernel(const float* __restrict__ z, const int64_t* __restrict__ label, float2* __restrict__ partial,
                     int h, int w, int K, int ignore_index) {
  extern __shared__ float patch[];                       // [kTailSrcR][kTailSrcC][2K + 2]
  __shared__ float red_s[kTailRows], red_c[kTailRows];
  const int C = 2 * K, CP = C + 2;
  const int H = 8 * h, W = 8 * w;
  const int b = blockIdx.z, Y0 = blockIdx.y * kTailRows, X0 = blockIdx.x * kTailCols;
  const float rh = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
  const float rw = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
  const int r0 = (int)(rh * (float)Y0), c0 = (int)(rw * (float)X0);
  const float* zb = z + (int64_t)b * h * w * C;
  for (int i = threadIdx.x; i < kTailSrcR * kTailSrcC * C; i += kTailRows * kTailCols) {
    const int c = i % C, rc = i / C, cc = rc % kTailSrcC, rr = rc / kTailSrcC;
    const int sr = min(r0 + rr, h - 1), sc = min(c0 + cc, w - 1);
    patch[(rr * kTailSrcC + cc) * CP + c] = __ldg(zb + ((int64_t)sr * w + sc) * C + c);
  }
  __syncthreads();
  const int Y = Y0 + (threadIdx.x >> 5), X = X0 + (threadIdx.x & 31);
  float sum = 0.f, cnt = 0.f;
  if (X < W && Y < H) {
    const int l = (int)label[((int64_t)b * H + Y) * W + X];
    if (l != ignore_index) {
      const float h1r = rh * (float)Y, w1r = rw * (float)X;
      const int h1 = (int)h1r, w1 = (int)w1r;
      const int h1p = (h1 < h - 1) ? 1 : 0, w1p = (w1 < w - 1) ? 1 : 0;
      const float h1l = h1r - (float)h1, h0l = 1.f - h1l, w1l = w1r - (float)w1, w0l = 1.f - w1l;
      const float* p00 = patch + ((h1 - r0) * kTailSrcC + (w1 - c0)) * CP;
      const float* p01 = p00 + w1p * CP;
      const float* p10 = p00 + h1p * kTailSrcC * CP;
      const float* p11 = p10 + w1p * CP;
      for (int k = 0; k < K; ++k) {
        const float2 a = *reinterpret_cast<const float2*>(p00 + 2 * k), bq = *reinterpret_cast<const float2*>(p01 + 2 * k);
        const float2 c = *reinterpret_cast<const float2*>(p10 + 2 * k), d = *reinterpret_cast<const float2*>(p11 + 2 * k);
        const float y0 = h0l * (w0l * a.x + w1l * bq.x) + h1l * (w0l * c.x + w1l * d.x);
        const float y1 = h0l * (w0l * a.y + w1l * bq.y) + h1l * (w0l * c.y + w1l * d.y);
        sum += ord_term(ord_prob(y0, y1), k < l);
      }
      cnt = 1.f;
    }
  }
  sum = warp_sum(sum); cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) { red_s[threadIdx.x >> 5] = sum; red_c[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f, c = 0.f;
    for (int i = 0; i < kTailRows; ++i) { s += red_s[i]; c += red_c[i]; }
    partial[((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = make_float2(s, c);
  }
}

// Backward of the fused tail loss.  The output pixels whose upper-left tap is source pixel (sr, sc) form that pixel's CELL (~8 x 8
// outputs: floor(rh Y) == sr, floor(rw X) == sc with the forward's own arithmetic).  One thread per (cell, bin k) walks its cell once,
// evaluates dL/dy1 of every output pixel ONCE and accumulates it towards the cell's four corners (the version with one thread per
// SOURCE pixel gathered from the four cells around it and so evaluated every exp / reciprocal four times: 0.50 ms of a 13.7 ms
// training step).  Corners are then combined without any order-dependent arithmetic: the cell to the left hands its right-hand
// corners over through shared memory (lanes are consecutive cells; a block recomputes one halo cell), and the two cell ROWS that meet
// in a source row each add once into the zero-initialised gradient -- two contributions, and fl(a + b) = fl(b + a).
// p is evaluated as 1 / (1 + exp(y0 - y1)) with the fast intrinsics on the interpolated DIFFERENCE of the pair: the gradient needs
// ~1e-6, not the bit pattern of the reference's e1 / (e0 + e1) that the forward / inference kernels reproduce.
constexpr int kTlbCells = kThreads - 1;       // cells a block owns; thread 0 recomputes the cell in front of them

__global__ void __launch_bounds__(kThreads)
tail_loss_bwd_kernel(const float* __restrict__ z, const int64_t* __restrict__ label, const float* __restrict__ count,
                     const float* __restrict__ grad_out, float* __restrict__ gz, int B, int h, int w, int K, int ignore_index) {
  __shared__ float s01[kThreads], s11[kThreads];
  const int k = blockIdx.y;
  const int64_t n_cells = (int64_t)B * h * w;
  const int64_t cell = (int64_t)blockIdx.x * kTlbCells + (int)threadIdx.x - 1;
  const bool valid = cell >= 0 && cell < n_cells;
  const int C = 2 * K, H = 8 * h, W = 8 * w;
  float c00 = 0.f, c01 = 0.f, c10 = 0.f, c11 = 0.f;
  int sc = 0, sr = 0, b = 0;
  if (valid) {
    sc = (int)(cell % w); sr = (int)((cell / w) % h); b = (int)(cell / ((int64_t)w * h));
    const float rh = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
    const float rw = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const bool last_r = sr == h - 1, last_c = sc == w - 1;       // both taps of the last row / column coincide
    const int r1 = last_r ? sr : sr + 1, c1 = last_c ? sc : sc + 1;
    auto pair_diff = [&](int rr, int cc) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(z + (((int64_t)b * h + rr) * w + cc) * C) + k);
      return v.x - v.y;
    };
    const float d00 = pair_diff(sr, sc), d01 = pair_diff(sr, c1), d10 = pair_diff(r1, sc), d11 = pair_diff(r1, c1);
    // candidate output rows / columns around the cell; membership is decided by the forward's own arithmetic
    const int Ylo = (h > 1) ? max(0, (int)((float)sr / rh) - 1) : 0, Yhi = (h > 1) ? min(H - 1, (int)((float)(sr + 1) / rh) + 1) : H - 1;
    const int Xlo = (w > 1) ? max(0, (int)((float)sc / rw) - 1) : 0, Xhi = (w > 1) ? min(W - 1, (int)((float)(sc + 1) / rw) + 1) : W - 1;
    const int64_t* lab_b = label + (int64_t)b * H * W;
    for (int Y = Ylo; Y <= Yhi; ++Y) {
      const float h1r = rh * (float)Y;
      const int h1 = (int)h1r;
      if (h1 != sr) continue;
      const float h1l = h1r - (float)h1, h0l = 1.f - h1l;
      const float vl = h0l * d00 + h1l * d10, vr = h0l * d01 + h1l * d11;
      const int64_t* lab_y = lab_b + (int64_t)Y * W;
      float racc0 = 0.f, racc1 = 0.f;
      for (int X = Xlo; X <= Xhi; ++X) {
        const float w1r = rw * (float)X;
        const int w1 = (int)w1r;
        if (w1 != sc) continue;
        const int l = (int)__ldg(lab_y + X);
        if (l == ignore_index) continue;
        const float w1l = w1r - (float)w1, w0l = 1.f - w1l;
        const float dyl = w0l * vl + w1l * vr;                        // y0 - y1 at (Y, X)
        const float p = __fdividef(1.0f, 1.0f + __expf(dyl)), q = 1.0f - p;
        const float dy1 = (k < l) ? ((p > 1e-8f) ? -q : 0.f) : ((q > 1e-8f) ? p : 0.f);      // clamp windows of log(p) / log(1 - p)
        racc0 = fmaf(last_c ? w0l + w1l : w0l, dy1, racc0);
        racc1 = fmaf(last_c ? 0.f : w1l, dy1, racc1);
      }
      const float wy0 = last_r ? h0l + h1l : h0l, wy1 = last_r ? 0.f : h1l;
      c00 = fmaf(wy0, racc0, c00); c01 = fmaf(wy0, racc1, c01);
      c10 = fmaf(wy1, racc0, c10); c11 = fmaf(wy1, racc1, c11);
    }
  }
  s01[threadIdx.x] = c01;
  s11[threadIdx.x] = c11;
  __syncthreads();
  if (!valid || threadIdx.x == 0) return;
  // the cell to the left (lane - 1) owns the right-hand corners that land in this column; the last cell of the previous row has none
  const float g = __ldg(grad_out) / __ldg(count);
  const float top = (c00 + s01[threadIdx.x - 1]) * g, bot = (c10 + s11[threadIdx.x - 1]) * g;
  float* dst = gz + (((int64_t)b * h + sr) * w + sc) * C + 2 * k;
  atomicAdd(dst, -top);
  atomicAdd(dst + 1, top);
  if (sr < h - 1) {
    dst += (int64_t)w * C;
    atomicAdd(dst, -bot);
    atomicAdd(dst + 1, bot);
  }
}

}  // namespace
}  // namespace acan

extern "C" size_t acan_tail_loss_ws_floats(int B, int h, int w) {
  if (B <= 0 || h <= 0 || w <= 0) return 0;
  return 2 * (size_t)B * h * ((8 * (size_t)w + acan::kTailCols - 1) / acan::kTailCols) + 8;
}

extern "C" int acan_tail_loss_fwd(const float* z_nhwc, const int64_t* label, float* loss_out, float* count_out, float* ws,
                                  int B, int h, int w, int K, int ignore_index, void* stream) {
  ACAN_CHECK_ARG(z_nhwc && label && loss_out && count_out && ws, "acan_tail_loss_fwd: null pointer");
  ACAN_CHECK_ARG(B > 0 && h > 0 && w > 0 && K > 0 && K <= 2048, "acan_tail_loss_fwd: B=%d h=%d w=%d K=%d invalid", B, h, w, K);
  ACAN_CHECK_ARG(B <= 65535 && h <= 65535, "acan_tail_loss_fwd: grid too large");
  ACAN_CHECK_ARG(reinterpret_cast<uintptr_t>(z_nhwc) % 8 == 0 && reinterpret_cast<uintptr_t>(ws) % 8 == 0, "acan_tail_loss_fwd: alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid((8 * w + kTailCols - 1) / kTailCols, h, B);
  const size_t smem = (size_t)kTailSrcR * kTailSrcC * (2 * K + 2) * sizeof(float);
  tail_loss_fwd_kernel<<<grid, kTailRows * kTailCols, smem, st>>>(z_nhwc, label, reinterpret_cast<float2*>(ws), h, w, K, ignore_index);
  if (int e = launch_status("acan_tail_loss_fwd")) return e;
  ord_loss_final_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const float2*>(ws), (int)(grid.x * grid.y * grid.z), loss_out, count_out);
  return launch_status("acan_tail_loss_fwd/final");
}

extern "C" int acan_tail_loss_bwd(const float* z_nhwc, const int64_t* label, const float* count, const float* grad_out, float* grad_z_nhwc,
                                  int B, int h, int w, int K, int ignore_index, void* stream) {
  ACAN_CHECK_ARG(z_nhwc && label && count && grad_out && grad_z_nhwc, "acan_tail_loss_bwd: null pointer");
  ACAN_CHECK_ARG(B > 0 && h > 0 && w > 0 && K > 0, "acan_tail_loss_bwd: sizes must be positive");
  ACAN_CHECK_ARG(reinterpret_cast<uintptr_t>(z_nhwc) % 8 == 0 && reinterpret_cast<uintptr_t>(grad_z_nhwc) % 8 == 0, "acan_tail_loss_bwd: alignment");
  ACAN_CHECK_ARG(K <= 65535, "acan_tail_loss_bwd: K=%d too large", K);
  const int64_t n_cells = (int64_t)B * h * w;
  const int64_t blocks = (n_cells + kTlbCells - 1) / kTlbCells;
  ACAN_CHECK_ARG(blocks <= 0x7fffffffLL, "acan_tail_loss_bwd: problem too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ACAN_CUDA(cudaMemsetAsync(grad_z_nhwc, 0, static_cast<size_t>(n_cells) * 2 * K * sizeof(float), st));      // every element takes <= 2 additions
  tail_loss_bwd_kernel<<<dim3((unsigned)blocks, (unsigned)K), kThreads, 0, st>>>(z_nhwc, label, count, grad_out, grad_z_nhwc, B, h, w, K, ignore_index);
  return launch_status("acan_tail_loss_bwd");
}

// ---------------------------------------------------------------- cross-entropy loss of the CE classifier
// CrossEntropy2d with reduction='sum', no class weights (losses.py:103-141, 164-196): per valid pixel, with p = softmax_k(y) and the
// smoothed target s_k = (1 - eps) [k == l] + eps prior_k [k != l]  (prior: uniform 1/(K-1), or gaussian N(k - l; sigma = K/10)),
//     loss = - sum_k ( s_k log clamp(p_k, 1e-8, 1e8) + (1 - s_k) log clamp(1 - p_k, 1e-8, 1e8) )
// -- a per-class binary cross-entropy on the softmax outputs, not the usual -log p_l -- averaged over the pixels whose label != ignore.
// Forward reads the scores [B,K,HW] once for the online (max, sum-exp) and once more (L2) for the terms; backward re-reads them and
// writes dL/dy_k = p_k (a_k - sum_c a_c p_c) / count,  a_k = dloss/dp_k inside the clamp windows.  The softmax output, the one-hot /
// smoothed label tensors and their gradients ([B,H,W,K] each in the reference) never exist.
namespace acan {
namespace {

struct CePrior { float eps; int kind; float inv_km1; float sigma; };     // kind 0: uniform, 1: gaussian

__device__ __forceinline__ float ce_target(int k, int l, const CePrior& pr) {
  if (k == l) return 1.0f - pr.eps;
  if (pr.eps == 0.f) return 0.f;
  if (pr.kind == 0) return pr.eps * pr.inv_km1;
  const float d = (float)(k - l);
  return pr.eps * expf(-d * d / (2.0f * pr.sigma * pr.sigma)) * rsqrtf(6.283185307179586f * pr.sigma * pr.sigma);
}

// one thread per pixel (consecutive threads = consecutive pixels: coalesced 4 B loads per class plane)
__global__ void __launch_bounds__(kThreads) ce_loss_fwd_kernel(const float* __restrict__ y, const int64_t* __restrict__ label, float2* __restrict__ partial,
                                                                int K, int64_t HW, int64_t total, int ignore_index, CePrior pr) {
  __shared__ float red_s[kThreads / 32], red_c[kThreads / 32];
  float sum = 0.f, cnt = 0.f;
  for (int64_t q = blockIdx.x * (int64_t)kThreads + threadIdx.x; q < total; q += (int64_t)gridDim.x * kThreads) {
    const int l = (int)label[q];
    if (l == ignore_index) continue;
    const int64_t b = q / HW, x = q - b * HW;
    const float* src = y + b * (int64_t)K * HW + x;
    float m = -INFINITY, z = 0.f;
    for (int k = 0; k < K; ++k) {
      const float v = __ldg(src + (int64_t)k * HW);
      if (v > m) { z = z * expf(m - v) + 1.0f; m = v; } else { z += expf(v - m); }
    }
    const float inv_z = 1.0f / z;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      const float p = expf(__ldg(src + (int64_t)k * HW) - m) * inv_z;
      const float s_k = ce_target(k, l, pr);
      acc -= s_k * logf(fminf(fmaxf(p, 1e-8f), 1e8f)) + (1.0f - s_k) * logf(fminf(fmaxf(1.0f - p, 1e-8f), 1e8f));
    }
    sum += acc;
    cnt += 1.f;
  }
  sum = warp_sum(sum); cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) { red_s[threadIdx.x >> 5] = sum; red_c[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s_ = 0.f, c_ = 0.f;
    for (int i = 0; i < kThreads / 32; ++i) { s_ += red_s[i]; c_ += red_c[i]; }
    partial[blockIdx.x] = make_float2(s_, c_);
  }
}

__global__ void __launch_bounds__(kThreads) ce_loss_bwd_kernel(const float* __restrict__ y, const int64_t* __restrict__ label, const float* __restrict__ count,
                                                                const float* __restrict__ grad_out, float* __restrict__ gy, int K, int64_t HW, int64_t total,
                                                                int ignore_index, CePrior pr) {
  const float g = __ldg(grad_out) / __ldg(count);
  for (int64_t q = blockIdx.x * (int64_t)kThreads + threadIdx.x; q < total; q += (int64_t)gridDim.x * kThreads) {
    const int64_t b = q / HW, x = q - b * HW;
    const float* src = y + b * (int64_t)K * HW + x;
    float* dst = gy + b * (int64_t)K * HW + x;
    const int l = (int)label[q];
    if (l == ignore_index) {
      for (int k = 0; k < K; ++k) dst[(int64_t)k * HW] = 0.f;
      continue;
    }
    float m = -INFINITY, z = 0.f;
    for (int k = 0; k < K; ++k) {
      const float v = __ldg(src + (int64_t)k * HW);
      if (v > m) { z = z * expf(m - v) + 1.0f; m = v; } else { z += expf(v - m); }
    }
    const float inv_z = 1.0f / z;
    auto slope = [&](float p, int k) {                       // a_k = d loss / d p_k (zero outside the clamp windows)
      const float s_k = ce_target(k, l, pr), qv = 1.0f - p;
      return -((p > 1e-8f && p < 1e8f) ? s_k / p : 0.f) + ((qv > 1e-8f && qv < 1e8f) ? (1.0f - s_k) / qv : 0.f);
    };
    float dot = 0.f;
    for (int k = 0; k < K; ++k) {
      const float p = expf(__ldg(src + (int64_t)k * HW) - m) * inv_z;
      dot = fmaf(slope(p, k), p, dot);
    }
    for (int k = 0; k < K; ++k) {
      const float p = expf(__ldg(src + (int64_t)k * HW) - m) * inv_z;
      dst[(int64_t)k * HW] = p * (slope(p, k) - dot) * g;
    }
  }
}

}  // namespace
}  // namespace acan

extern "C" int acan_ce_loss_fwd(const float* scores, const int64_t* label, float* loss_out, float* count_out, float* ws,
                                int B, int K, int64_t HW, int ignore_index, float eps, int prior_kind, void* stream) {
  ACAN_CHECK_ARG(scores && label && loss_out && count_out && ws, "acan_ce_loss_fwd: null pointer");
  ACAN_CHECK_ARG(B > 0 && K > 1 && HW > 0, "acan_ce_loss_fwd: B=%d K=%d HW=%lld invalid", B, K, (long long)HW);
  ACAN_CHECK_ARG(eps >= 0.f && eps <= 1.f && (prior_kind == 0 || prior_kind == 1), "acan_ce_loss_fwd: eps=%g prior_kind=%d invalid", eps, prior_kind);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t total = (int64_t)B * HW;
  const int grid = grid_for(total);
  const CePrior pr{eps, prior_kind, 1.0f / (float)(K - 1), (float)K / 10.0f};
  ce_loss_fwd_kernel<<<grid, kThreads, 0, st>>>(scores, label, reinterpret_cast<float2*>(ws), K, HW, total, ignore_index, pr);
  if (int e = launch_status("acan_ce_loss_fwd")) return e;
  ord_loss_final_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const float2*>(ws), grid, loss_out, count_out);
  return launch_status("acan_ce_loss_fwd/final");
}

extern "C" int acan_ce_loss_bwd(const float* scores, const int64_t* label, const float* count, const float* grad_out, float* grad_scores,
                                int B, int K, int64_t HW, int ignore_index, float eps, int prior_kind, void* stream) {
  ACAN_CHECK_ARG(scores && label && count && grad_out && grad_scores, "acan_ce_loss_bwd: null pointer");
  ACAN_CHECK_ARG(B > 0 && K > 1 && HW > 0, "acan_ce_loss_bwd: sizes invalid");
  ACAN_CHECK_ARG(eps >= 0.f && eps <= 1.f && (prior_kind == 0 || prior_kind == 1), "acan_ce_loss_bwd: eps=%g prior_kind=%d invalid", eps, prior_kind);
  const int64_t total = (int64_t)B * HW;
  const CePrior pr{eps, prior_kind, 1.0f / (float)(K - 1), (float)K / 10.0f};
  ce_loss_bwd_kernel<<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(scores, label, count, grad_out, grad_scores, K, HW, total,
                                                                                        ignore_index, pr);
  return launch_status("acan_ce_loss_bwd");
}

extern "C" int acan_ordinal_loss_ws_floats(void) { return 2 * acan::kNumSMs * acan::kBlocksPerSM + 8; }

extern "C" int acan_ordinal_loss_fwd(const float* logits, const int64_t* label, float* loss_out, float* count_out, float* ws,
                                     int B, int K, int64_t HW, int ignore_index, void* stream) {
  ACAN_CHECK_ARG(logits && label && loss_out && count_out && ws, "acan_ordinal_loss_fwd: null pointer");
  ACAN_CHECK_ARG(B > 0 && K > 0 && HW > 0 && HW % 4 == 0, "acan_ordinal_loss_fwd: need positive sizes and HW %% 4 == 0 (HW=%lld)", (long long)HW);
  ACAN_CHECK_ARG(reinterpret_cast<uintptr_t>(logits) % 16 == 0 && reinterpret_cast<uintptr_t>(label) % 16 == 0 && reinterpret_cast<uintptr_t>(ws) % 8 == 0,
                 "acan_ordinal_loss_fwd: alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t hw4 = HW / 4, total = (int64_t)B * hw4;
  const int grid = grid_for(total);
  ord_loss_fwd_kernel<<<grid, kThreads, 0, st>>>(reinterpret_cast<const float4*>(logits), label, reinterpret_cast<float2*>(ws), K, hw4, total, ignore_index);
  if (int e = launch_status("acan_ordinal_loss_fwd")) return e;
  ord_loss_final_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const float2*>(ws), grid, loss_out, count_out);
  return launch_status("acan_ordinal_loss_fwd/final");
}

extern "C" int acan_ordinal_loss_bwd(const float* logits, const int64_t* label, const float* count, const float* grad_out, float* grad_logits,
                                     int B, int K, int64_t HW, int ignore_index, void* stream) {
  ACAN_CHECK_ARG(logits && label && count && grad_out && grad_logits, "acan_ordinal_loss_bwd: null pointer");
  ACAN_CHECK_ARG(B > 0 && K > 0 && HW > 0 && HW % 4 == 0, "acan_ordinal_loss_bwd: need positive sizes and HW %% 4 == 0");
  const int64_t hw4 = HW / 4, total = (int64_t)B * K * hw4;
  ord_loss_bwd_kernel<<<grid_for(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(logits), label, count, grad_out, reinterpret_cast<float4*>(grad_logits), K, hw4, total, ignore_index);
  return launch_status("acan_ordinal_loss_bwd");
}
