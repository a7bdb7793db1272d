// bn.cu -- training-mode BatchNorm2d over channels-last half-precision activations, fused with the ReLU (and the residual
// add) that follows it: the Q/K projection's BN + ReLU of the context aggregation module (reference sadecoder.py:25-30, the
// "BN batch-stat reduction in train" of SURVEY §8 row a1 / f-3), and -- the same arithmetic, nn.BatchNorm2d in train() --
// every conv -> BN -> ReLU of the backbone the module sits on (model.py:128-160, 216-226).
//
// x is [M, C] (M = B*H*W positions, C channels contiguous: what a channels_last cuDNN convolution writes), fp16 or bf16.
// All four kernels are HBM-bound column passes; a thread owns 8 consecutive channels (one 16 B load per row):
//   forward : bn_stats_kernel   per-channel sum / sum of squares (fp32 partials per row block; the LAST row block of a
//                               channel slab -- self-resetting arrival counter -- reduces them in double, writes mean,
//                               1/sqrt(var+eps), the running statistics and the folded (scale, shift))      reads x once
//             bn_apply_kernel   y = [relu]( x*scale + shift [+ residual] )                                  reads x, writes y
//   backward: bn_bwd_reduce_kernel  g = dy * [y > 0];  sum g, sum g*xhat  (same last-block finalisation -> dgamma, dbeta,
//                               and the two per-channel coefficients of dx)                                 reads dy, y, x
//             bn_bwd_apply_kernel   dx = scale * (g - a - xhat * b)  [+ g written out for the residual branch]
// Algorithmic bytes per element (2-byte activations): forward 2 + 4 (+2 residual) = 6..8 B, backward 6 + 8 (+2) = 14..16 B,
// against ~4 launches x several passes each of the framework's generic channels-last kernels.
#include "common.cuh"
#include <cuda_bf16.h>
#include <atomic>
#include <cstdlib>

namespace acan {
namespace {

constexpr int kBnSlab = 256;          // channels per block column (32 lanes x 8 channels)
constexpr int kBnRedY = 16;           // reduction kernels: 32 x 16 threads
constexpr int kBnAppY = 8;            // apply kernels: 32 x 8 threads
constexpr int kBnMaxRowBlocks = 128;
constexpr int kBnFinalDepth = 16;     // row-block partials one thread of the finalising block sums (all loaded before the first add)
constexpr int kBnCounterSlots = 64;   // one arrival counter per 256-channel slab: C <= 16384

template <bool BF16>
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (BF16) {
      f[2 * e] = __uint_as_float(w[e] << 16);
      f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
    } else {
      const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
      f[2 * e] = t.x;
      f[2 * e + 1] = t.y;
    }
  }
}

template <bool BF16>
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (BF16) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
      w[e] = *reinterpret_cast<const uint32_t*>(&h);
    } else {
      const __half2 h = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
      w[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// Geometry shared by the four kernels.  A block is 32 x Y threads; lane x owns the 8-channel group
// `slab * 32 + x % groups` where groups = min(C/8 - slab*32, 32); when a slab has fewer than 32 groups (C = 64, 128) the
// spare lanes take further rows (rows_per_warp = 32 / groups when that divides), so every lane loads 16 B.
struct ColGeom {
  int groups, rows_per_warp, cg, rsub;
};
__device__ __forceinline__ ColGeom col_geom(int C, int slab_groups = 32) {
  ColGeom g;
  const int c8 = C >> 3;
  g.groups = min(c8 - static_cast<int>(blockIdx.x) * slab_groups, slab_groups);
  g.rows_per_warp = (32 % g.groups) == 0 ? 32 / g.groups : 1;     // otherwise the lanes beyond `groups` idle
  g.cg = static_cast<int>(threadIdx.x) % g.groups;
  g.rsub = static_cast<int>(threadIdx.x) / g.groups;        // < rows_per_warp for active lanes
  return g;
}

// block-level reduction of per-thread (a[8], b[8]) over the row dimension -> valid for threads (y=0, x < groups).
// Two steps, both with all 512 threads: thread (value v, lane x) sums the 16 warps' entries, then the <= 4 row sub-groups of a warp
// are added by the receiving threads (the serial version -- 16 x rows_per_warp x 16 loads by `groups` threads -- cost ~1 us of the
// ~10 us latency chain of a cooperative BN kernel).
__device__ __forceinline__ void block_col_reduce(float (&a)[8], float (&b)[8], const ColGeom& g, bool active, float* sm /* [16][kBnRedY][32] */) {
  // element-major layout: the 32 lanes of a warp touch 32 consecutive words in every access (no bank conflicts)
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sm[(e * kBnRedY + threadIdx.y) * 32 + threadIdx.x] = active ? a[e] : 0.f;
    sm[((8 + e) * kBnRedY + threadIdx.y) * 32 + threadIdx.x] = active ? b[e] : 0.f;
  }
  __syncthreads();
  static_assert(kBnRedY == 16, "one thread per (value, lane) pair");
  float t = 0.f;
  {
    const float* col = sm + (threadIdx.y * kBnRedY) * 32 + threadIdx.x;      // value index = threadIdx.y (0..15), lane = threadIdx.x
#pragma unroll
    for (int y = 0; y < kBnRedY; ++y) t += col[y * 32];
  }
  __syncthreads();
  sm[threadIdx.y * 32 + threadIdx.x] = t;
  __syncthreads();
  if (threadIdx.y == 0 && threadIdx.x < g.groups) {
#pragma unroll
    for (int e = 0; e < 8; ++e) { a[e] = 0.f; b[e] = 0.f; }
    for (int r = 0; r < g.rows_per_warp; ++r) {
      const int lane = r * g.groups + threadIdx.x;
#pragma unroll
      for (int e = 0; e < 8; ++e) { a[e] += sm[e * 32 + lane]; b[e] += sm[(8 + e) * 32 + lane]; }
    }
  }
}

// Last row block of a slab: sum the row-block partials of its `nch` channels in double, in a fixed order, with every thread of
// the block: thread t takes channel t % nch and the partials t / nch, t / nch + nsl, ...; the nsl slices meet in shared memory.
// Returns the totals to the threads t < nch (others get garbage they must not use).
__device__ __forceinline__ void final_col_sum(const float* __restrict__ partial, int n_rb, int C, int c0, int nch, double* smd /* [2*512] */, double& S, double& Q) {
  const int tid = threadIdx.y * 32 + threadIdx.x, nthr = 32 * kBnRedY;
  const int nsl = nthr / nch;                                  // >= 2 (nch <= 256)
  const int cl = tid % nch, sl = tid / nch;
  double s = 0.0, q = 0.0;
  if (sl < nsl) {
    // the host caps the row blocks at kBnFinalDepth * nsl: all of a thread's partials are in flight at once
    float2 v[kBnFinalDepth];
#pragma unroll
    for (int k = 0; k < kBnFinalDepth; ++k) {
      const int rb = sl + k * nsl;
      v[k] = rb < n_rb ? __ldcg(reinterpret_cast<const float2*>(partial) + static_cast<size_t>(rb) * C + c0 + cl) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < kBnFinalDepth; ++k) { s += v[k].x; q += v[k].y; }
  }
  smd[tid] = s;
  smd[nthr + tid] = q;
  __syncthreads();
  S = 0.0; Q = 0.0;
  if (tid < nch) {
    for (int k = 0; k < nsl; ++k) { S += smd[k * nch + tid]; Q += smd[nthr + k * nch + tid]; }
  }
}

// partial[(rb * C + c) * 2 + {0,1}];  counters: one int per slab, zero between calls (the last block resets its own)
template <bool BF16>
__global__ void __launch_bounds__(32 * kBnRedY, 2)
bn_stats_kernel(const uint4* __restrict__ x, const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ partial,
                int* __restrict__ counters, float* __restrict__ save_mean, float* __restrict__ save_invstd, float* __restrict__ scale_shift,
                float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, float eps, int64_t M, int C) {
  __shared__ __align__(16) float sm[kBnRedY * 32 * 16];
  __shared__ int is_last;
  const ColGeom g = col_geom(C);
  const bool active = static_cast<int>(threadIdx.x) < g.groups * g.rows_per_warp;
  const int c8 = C >> 3;
  const int64_t row_step = static_cast<int64_t>(gridDim.y) * kBnRedY * g.rows_per_warp;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
  if (active) {
    const uint4* col = x + blockIdx.x * 32 + g.cg;
    int64_t r = (static_cast<int64_t>(blockIdx.y) * kBnRedY + threadIdx.y) * g.rows_per_warp + g.rsub;
    for (; r < M; r += 8 * row_step) {                            // 8 independent 16 B loads in flight; rows beyond M read as 0
      uint4 u[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t rr = r + i * row_step;
        u[i] = ldg_nc_u4(col + (rr < M ? rr : r) * c8);
        if (rr >= M) u[i] = make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float f[8];
        unpack8<BF16>(u[i], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
      }
    }
  }
  block_col_reduce(s, q, g, active, sm);
  if (threadIdx.y == 0 && static_cast<int>(threadIdx.x) < g.groups) {
    float* p = partial + (static_cast<size_t>(blockIdx.y) * C + (blockIdx.x * 32 + threadIdx.x) * 8) * 2;
#pragma unroll
    for (int e = 0; e < 8; ++e) { p[2 * e] = s[e]; p[2 * e + 1] = q[e]; }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) is_last = (atomicAdd(counters + blockIdx.x, 1) == static_cast<int>(gridDim.y) - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double S, Q;
  const int nch = g.groups * 8, c0 = blockIdx.x * kBnSlab;
  final_col_sum(partial, gridDim.y, C, c0, nch, reinterpret_cast<double*>(sm), S, Q);
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < nch) {
    const int c = c0 + tid;
    const double mean = S / static_cast<double>(M);
    double var = Q / static_cast<double>(M) - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float invstd = static_cast<float>(rsqrt(var + static_cast<double>(eps)));
    save_mean[c] = static_cast<float>(mean);
    save_invstd[c] = invstd;
    const float sc = weight[c] * invstd;
    scale_shift[c] = sc;
    scale_shift[C + c] = bias[c] - static_cast<float>(mean) * sc;
    if (running_mean != nullptr) {
      const double unbiased = M > 1 ? var * static_cast<double>(M) / static_cast<double>(M - 1) : var;
      running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
      running_var[c] = (1.0f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
    }
  }
  if (tid == 0) counters[blockIdx.x] = 0;
}

// y = [relu](x * scale + shift [+ residual]); a thread keeps its 8 channels' (scale, shift) in registers and walks the rows
template <bool BF16>
__global__ void __launch_bounds__(32 * kBnAppY)
bn_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ residual, uint4* __restrict__ y, const float* __restrict__ scale_shift,
                int relu, int64_t M, int C) {
  const ColGeom g = col_geom(C);
  if (static_cast<int>(threadIdx.x) >= g.groups * g.rows_per_warp) return;
  const int c8 = C >> 3, cbase = (blockIdx.x * 32 + g.cg) * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { sc[e] = __ldg(scale_shift + cbase + e); sh[e] = __ldg(scale_shift + C + cbase + e); }
  const int64_t off = blockIdx.x * 32 + g.cg;
  const int64_t row_step = static_cast<int64_t>(gridDim.y) * kBnAppY * g.rows_per_warp;
  int64_t r = (static_cast<int64_t>(blockIdx.y) * kBnAppY + threadIdx.y) * g.rows_per_warp + g.rsub;
  auto one = [&](const uint4& ux, const uint4& ur) {
    float f[8], rr[8];
    unpack8<BF16>(ux, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], sc[e], sh[e]);
    if (residual != nullptr) {
      unpack8<BF16>(ur, rr);
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] += rr[e];
    }
    if (relu) {
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
    }
    return pack8<BF16>(f);
  };
  for (; r + 3 * row_step < M; r += 4 * row_step) {
    uint4 ux[4], ur[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t idx = (r + i * row_step) * c8 + off;
      ux[i] = ldg_nc_u4(x + idx);
      ur[i] = (residual != nullptr) ? ldg_nc_u4(residual + idx) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) y[(r + i * row_step) * c8 + off] = one(ux[i], ur[i]);
  }
  for (; r < M; r += row_step) {
    const int64_t idx = r * c8 + off;
    const uint4 ux = ldg_nc_u4(x + idx);
    const uint4 ur = (residual != nullptr) ? ldg_nc_u4(residual + idx) : make_uint4(0, 0, 0, 0);
    y[idx] = one(ux, ur);
  }
}

// coef (used by bn_bwd_apply_kernel): dx = A g + Bc x + Cc per channel -> coef[0..C) = A = w invstd,
// [C..2C) = Bc = -A invstd b, [2C..3C) = Cc = A (invstd b mean - a),   a = sum g / M,  b = sum(g xhat) / M
template <bool BF16>
__global__ void __launch_bounds__(32 * kBnRedY)
bn_bwd_reduce_kernel(const uint4* __restrict__ x, const uint4* __restrict__ y, const uint4* __restrict__ gy, const float* __restrict__ weight,
                     const float* __restrict__ save_mean, const float* __restrict__ save_invstd, float* __restrict__ partial, int* __restrict__ counters,
                     float* __restrict__ grad_weight, float* __restrict__ grad_bias, float* __restrict__ coef, int64_t M, int C) {
  __shared__ __align__(16) float sm[kBnRedY * 32 * 16];
  __shared__ int is_last;
  const ColGeom g = col_geom(C);
  const bool active = static_cast<int>(threadIdx.x) < g.groups * g.rows_per_warp;
  const int c8 = C >> 3;
  const int64_t row_step = static_cast<int64_t>(gridDim.y) * kBnRedY * g.rows_per_warp;
  float sg[8], sx[8];                                       // sum g, sum g * (x - mean)   (invstd applied once at the end)
#pragma unroll
  for (int e = 0; e < 8; ++e) { sg[e] = 0.f; sx[e] = 0.f; }
  if (active) {
    const int cbase = (blockIdx.x * 32 + g.cg) * 8;
    float mu[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) mu[e] = __ldg(save_mean + cbase + e);
    const int64_t off = blockIdx.x * 32 + g.cg;
    int64_t r = (static_cast<int64_t>(blockIdx.y) * kBnRedY + threadIdx.y) * g.rows_per_warp + g.rsub;
    auto acc = [&](const uint4& ux, const uint4& ug, const uint4& uy) {
      float fx[8], fg[8], fy[8];
      unpack8<BF16>(ux, fx);
      unpack8<BF16>(ug, fg);
      if (y != nullptr) {
        unpack8<BF16>(uy, fy);
#pragma unroll
        for (int e = 0; e < 8; ++e) fg[e] = fy[e] > 0.f ? fg[e] : 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { sg[e] += fg[e]; sx[e] = fmaf(fg[e], fx[e] - mu[e], sx[e]); }
    };
    for (; r < M; r += 4 * row_step) {                             // 4 rows x up to 3 tensors in flight; rows beyond M: g = 0
      uint4 ux[4], ug[4], uy[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t rr = r + i * row_step;
        const int64_t idx = (rr < M ? rr : r) * c8 + off;
        ux[i] = ldg_nc_u4(x + idx);
        ug[i] = ldg_nc_u4(gy + idx);
        uy[i] = (y != nullptr) ? ldg_nc_u4(y + idx) : make_uint4(0, 0, 0, 0);
        if (rr >= M) ug[i] = make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) acc(ux[i], ug[i], uy[i]);
    }
  }
  block_col_reduce(sg, sx, g, active, sm);
  if (threadIdx.y == 0 && static_cast<int>(threadIdx.x) < g.groups) {
    float* p = partial + (static_cast<size_t>(blockIdx.y) * C + (blockIdx.x * 32 + threadIdx.x) * 8) * 2;
#pragma unroll
    for (int e = 0; e < 8; ++e) { p[2 * e] = sg[e]; p[2 * e + 1] = sx[e]; }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) is_last = (atomicAdd(counters + blockIdx.x, 1) == static_cast<int>(gridDim.y) - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double S, Q;
  const int nch = g.groups * 8, c0 = blockIdx.x * kBnSlab;
  final_col_sum(partial, gridDim.y, C, c0, nch, reinterpret_cast<double*>(sm), S, Q);
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < nch) {
    const int c = c0 + tid;
    const double is = static_cast<double>(save_invstd[c]), mean = static_cast<double>(save_mean[c]), w = static_cast<double>(weight[c]);
    const double sum_gx = Q * is;                            // sum g * xhat
    grad_bias[c] = static_cast<float>(S);
    grad_weight[c] = static_cast<float>(sum_gx);
    const double a = S / static_cast<double>(M), b = sum_gx / static_cast<double>(M), A = w * is;
    coef[c] = static_cast<float>(A);
    coef[C + c] = static_cast<float>(-A * is * b);
    coef[2 * C + c] = static_cast<float>(A * (is * b * mean - a));
  }
  if (tid == 0) counters[blockIdx.x] = 0;
}

template <bool BF16>
__global__ void __launch_bounds__(32 * kBnAppY)
bn_bwd_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ y, const uint4* __restrict__ gy, const float* __restrict__ coef,
                    uint4* __restrict__ gx, uint4* __restrict__ gres, int64_t M, int C) {
  const ColGeom g = col_geom(C);
  if (static_cast<int>(threadIdx.x) >= g.groups * g.rows_per_warp) return;
  const int c8 = C >> 3, cbase = (blockIdx.x * 32 + g.cg) * 8;
  float A[8], Bc[8], Cc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { A[e] = __ldg(coef + cbase + e); Bc[e] = __ldg(coef + C + cbase + e); Cc[e] = __ldg(coef + 2 * C + cbase + e); }
  const int64_t off = blockIdx.x * 32 + g.cg;
  const int64_t row_step = static_cast<int64_t>(gridDim.y) * kBnAppY * g.rows_per_warp;
  int64_t r = (static_cast<int64_t>(blockIdx.y) * kBnAppY + threadIdx.y) * g.rows_per_warp + g.rsub;
  auto one = [&](int64_t idx, const uint4& ux, const uint4& ug, const uint4& uy) {
    float fx[8], fg[8], fy[8], o[8];
    unpack8<BF16>(ux, fx);
    unpack8<BF16>(ug, fg);
    if (y != nullptr) {
      unpack8<BF16>(uy, fy);
#pragma unroll
      for (int e = 0; e < 8; ++e) fg[e] = fy[e] > 0.f ? fg[e] : 0.f;
      if (gres != nullptr) gres[idx] = pack8<BF16>(fg);      // gradient of the residual branch: the masked upstream gradient
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = fmaf(A[e], fg[e], fmaf(Bc[e], fx[e], Cc[e]));
    gx[idx] = pack8<BF16>(o);
  };
  for (; r + row_step < M; r += 2 * row_step) {
    uint4 ux[2], ug[2], uy[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t idx = (r + i * row_step) * c8 + off;
      ux[i] = ldg_nc_u4(x + idx);
      ug[i] = ldg_nc_u4(gy + idx);
      uy[i] = (y != nullptr) ? ldg_nc_u4(y + idx) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) one((r + i * row_step) * c8 + off, ux[i], ug[i], uy[i]);
  }
  for (; r < M; r += row_step) {
    const int64_t idx = r * c8 + off;
    const uint4 ux = ldg_nc_u4(x + idx), ug = ldg_nc_u4(gy + idx);
    const uint4 uy = (y != nullptr) ? ldg_nc_u4(y + idx) : make_uint4(0, 0, 0, 0);
    one(idx, ux, ug, uy);
  }
}

// ---------------------------------------------------------------- single-launch variants (cooperative grid)
// The activations of one layer (3 .. 46 MB at the C4 shapes) fit the 126 MB L2, and at 5 .. 25 us per pass the two-launch
// form above is dominated by launch gaps and by the serial "last block finalises" tail.  Here one cooperative launch does
// both passes: every block reduces its rows, the blocks of a channel slab meet at a barrier, EVERY block then reduces the
// slab's partials itself (in parallel, same fixed order -> identical results) and walks its rows again -- now L2 hits --
// to write the output.  The grid is at most one block per SM (checked by the cooperative launch), so the barrier cannot
// deadlock; a bounded spin traps instead of hanging if that is ever violated.
// Sense-reversing: `arrive` counts the blocks of the slab and is reset by the last one, `gen` flips once per barrier and is never reset
// (any start value works).  One release-atomic and acquire-loads on the critical path; the version with an arrive and a depart counter
// and two full fences cost ~2 us more per kernel.
__device__ __forceinline__ void slab_barrier(int* arrive, int* gen, int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) {
    int g0, old;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(g0) : "l"(gen) : "memory");       // cannot flip before this block has arrived
    asm volatile("atom.add.acq_rel.gpu.global.s32 %0, [%1], 1;" : "=r"(old) : "l"(arrive) : "memory");
    if (old == nblocks - 1) {
      asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" :: "l"(arrive), "r"(0) : "memory");
      asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(gen), "r"(g0 + 1) : "memory");
    } else {
      const long long t0 = clock64();
      int gn;
      do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(gn) : "l"(gen) : "memory");
        if (gn == g0 && clock64() - t0 > 4000000000LL) {      // ~2 s: blocks were not co-resident
          printf("acan: BN slab barrier timeout (slab %d block %d of %d)\n", blockIdx.x, blockIdx.y, nblocks);
          __trap();
        }
      } while (gn == g0);
    }
  }
  __syncthreads();
}

template <bool BF16>
__global__ void __launch_bounds__(32 * kBnRedY)
bn_fwd_coop_kernel(const uint4* __restrict__ x, const uint4* __restrict__ residual, uint4* __restrict__ y, const float* __restrict__ weight,
                   const float* __restrict__ bias, float* __restrict__ partial, int* __restrict__ counters, float* __restrict__ save_mean,
                   float* __restrict__ save_invstd, float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, float eps,
                   int relu, int64_t M, int C, int slab_groups) {
  __shared__ __align__(16) float sm[kBnRedY * 32 * 16];
  __shared__ float sc_s[kBnSlab], sh_s[kBnSlab];
  const ColGeom g = col_geom(C, slab_groups);
  const bool active = static_cast<int>(threadIdx.x) < g.groups * g.rows_per_warp;
  const int c8 = C >> 3;
  const int64_t row_step = static_cast<int64_t>(gridDim.y) * kBnRedY * g.rows_per_warp;
  const int64_t r0 = (static_cast<int64_t>(blockIdx.y) * kBnRedY + threadIdx.y) * g.rows_per_warp + g.rsub;
  const int64_t off = blockIdx.x * slab_groups + g.cg;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
  if (active) {
    for (int64_t r = r0; r < M; r += 8 * row_step) {
      uint4 u[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t rr = r + i * row_step;
        u[i] = ldg_nc_u4(x + (rr < M ? rr : r) * c8 + off);
        if (rr >= M) u[i] = make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float f[8];
        unpack8<BF16>(u[i], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
      }
    }
  }
  block_col_reduce(s, q, g, active, sm);
  if (threadIdx.y == 0 && static_cast<int>(threadIdx.x) < g.groups) {
    float* p = partial + (static_cast<size_t>(blockIdx.y) * C + (blockIdx.x * slab_groups + threadIdx.x) * 8) * 2;
#pragma unroll
    for (int e = 0; e < 8; ++e) { p[2 * e] = s[e]; p[2 * e + 1] = q[e]; }
  }
  slab_barrier(counters + blockIdx.x, counters + kBnCounterSlots + blockIdx.x, gridDim.y);
  double S, Q;
  const int nch = g.groups * 8, c0 = blockIdx.x * slab_groups * 8;
  final_col_sum(partial, gridDim.y, C, c0, nch, reinterpret_cast<double*>(sm), S, Q);
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < nch) {
    const int c = c0 + tid;
    const double mean = S / static_cast<double>(M);
    double var = Q / static_cast<double>(M) - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float invstd = static_cast<float>(rsqrt(var + static_cast<double>(eps)));
    const float scv = __ldg(weight + c) * invstd;
    sc_s[tid] = scv;
    sh_s[tid] = __ldg(bias + c) - static_cast<float>(mean) * scv;
    if (blockIdx.y == 0) {
      save_mean[c] = static_cast<float>(mean);
      save_invstd[c] = invstd;
      if (running_mean != nullptr) {
        const double unbiased = M > 1 ? var * static_cast<double>(M) / static_cast<double>(M - 1) : var;
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
      }
    }
  }
  __syncthreads();
  if (!active) return;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { sc[e] = sc_s[g.cg * 8 + e]; sh[e] = sh_s[g.cg * 8 + e]; }
  for (int64_t r = r0; r < M; r += 4 * row_step) {             // second pass over the same rows: L2 hits
    uint4 ux[4], ur[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t rr = r + i * row_step;
      const int64_t idx = (rr < M ? rr : r) * c8 + off;
      ux[i] = ldg_nc_u4(x + idx);
      ur[i] = (residual != nullptr) ? ldg_nc_u4(residual + idx) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t rr = r + i * row_step;
      if (rr < M) {
        float f[8], rv[8];
        unpack8<BF16>(ux[i], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], sc[e], sh[e]);
        if (residual != nullptr) {
          unpack8<BF16>(ur[i], rv);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] += rv[e];
        }
        if (relu) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
        }
        y[rr * c8 + off] = pack8<BF16>(f);
      }
    }
  }
}

// ---------------------------------------------------------------- forward, activations staged in shared memory by TMA
// When a block's share of the layer fits its shared memory (148 x ~190 KB = 28 MB: every layer of a ResNet at the configs[3] batch), the
// block pulls its rows in with a handful of tensor-map copies (64 rows x one channel slab each: nothing passes through registers, and
// all of the block's bytes are in flight at once instead of 8 loads per thread per round trip), takes the statistics from shared
// memory as the boxes land, and after the slab barrier normalises the SAME shared-memory copy: x is read from HBM / L2 exactly once.
// Rows are contiguous per block here ([row_lo, row_hi)); lanes map to channels and rows as in the kernels above.
constexpr int kBnBoxRows = 64;        // rows per TMA box (a multiple of kBnRedY * rows_per_warp for every slab width)
constexpr int kBnMaxBoxes = 24;
constexpr int kBnTmaSmemMax = 190 * 1024;

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

template <bool BF16>
__global__ void __launch_bounds__(32 * kBnRedY, 1)
bn_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmap_x, const uint4* __restrict__ residual, uint4* __restrict__ y,
                  const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ partial, int* __restrict__ counters,
                  float* __restrict__ save_mean, float* __restrict__ save_invstd, float* __restrict__ running_mean, float* __restrict__ running_var,
                  float momentum, float eps, int relu, int64_t M, int C, int slab_groups, int rows_per_block) {
  extern __shared__ uint8_t bn_dyn_smem[];
  uint8_t* cache = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(bn_dyn_smem) + 1023) & ~static_cast<uintptr_t>(1023));
  __shared__ __align__(16) float sm[kBnRedY * 32 * 16];
  __shared__ float sc_s[kBnSlab], sh_s[kBnSlab];
  __shared__ __align__(8) uint64_t bars[kBnMaxBoxes];
  ColGeom g = col_geom(C, slab_groups);
  if (g.rows_per_warp > 4) g.rows_per_warp = 4;                        // slabs of 1, 2 or 4 channel groups: a sweep must not exceed a box (64 rows)
  const bool active = static_cast<int>(threadIdx.x) < g.groups * g.rows_per_warp;
  const int c8 = C >> 3;
  const int pitch = slab_groups * 16;                                  // bytes per staged row (the box is slab_groups * 8 channels wide)
  const uint32_t box_bytes = static_cast<uint32_t>(kBnBoxRows) * pitch;
  const int64_t row_lo = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t row_hi = (row_lo + rows_per_block < M) ? row_lo + rows_per_block : M;
  const int nbox = static_cast<int>((row_hi - row_lo + kBnBoxRows - 1) / kBnBoxRows);
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid == 0) {
    tma_prefetch_desc(&tmap_x);
    for (int j = 0; j < nbox; ++j) mbar_init(&bars[j], 1);
    fence_barrier_init();
    for (int j = 0; j < nbox; ++j) {
      mbar_expect_tx(&bars[j], box_bytes);                             // rows beyond M arrive as zeros and count all the same
      tma_load_2d(cache + static_cast<size_t>(j) * box_bytes, &tmap_x, &bars[j], blockIdx.x * slab_groups * 8,
                  static_cast<int>(row_lo) + j * kBnBoxRows);
    }
  }
  __syncthreads();
  const int ri = kBnRedY * g.rows_per_warp;                            // rows one sweep of the block covers: 16, 32 or 64
  const int sweeps = kBnBoxRows / ri;
  const int my_row = threadIdx.y * g.rows_per_warp + g.rsub;           // row of this thread inside a sweep
  const uint8_t* my_cache = cache + static_cast<size_t>(my_row) * pitch + g.cg * 16;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { s[e] = 0.f; q[e] = 0.f; }
  for (int j = 0; j < nbox; ++j) {
    mbar_wait(&bars[j], 0);
    if (active) {
      for (int k = 0; k < sweeps; ++k) {
        const int i = j * kBnBoxRows + k * ri;
        if (row_lo + i + my_row < row_hi) {
          const uint4 u = *reinterpret_cast<const uint4*>(my_cache + static_cast<size_t>(i) * pitch);
          float f[8];
          unpack8<BF16>(u, f);
#pragma unroll
          for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
        }
      }
    }
  }
  block_col_reduce(s, q, g, active, sm);
  if (threadIdx.y == 0 && static_cast<int>(threadIdx.x) < g.groups) {
    float* p = partial + (static_cast<size_t>(blockIdx.y) * C + (blockIdx.x * slab_groups + threadIdx.x) * 8) * 2;
#pragma unroll
    for (int e = 0; e < 8; ++e) { p[2 * e] = s[e]; p[2 * e + 1] = q[e]; }
  }
  slab_barrier(counters + blockIdx.x, counters + kBnCounterSlots + blockIdx.x, gridDim.y);
  double S, Q;
  const int nch = g.groups * 8, c0 = blockIdx.x * slab_groups * 8;
  final_col_sum(partial, gridDim.y, C, c0, nch, reinterpret_cast<double*>(sm), S, Q);
  if (tid < nch) {
    const int c = c0 + tid;
    const double mean = S / static_cast<double>(M);
    double var = Q / static_cast<double>(M) - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const float invstd = static_cast<float>(rsqrt(var + static_cast<double>(eps)));
    const float scv = __ldg(weight + c) * invstd;
    sc_s[tid] = scv;
    sh_s[tid] = __ldg(bias + c) - static_cast<float>(mean) * scv;
    if (blockIdx.y == 0) {
      save_mean[c] = static_cast<float>(mean);
      save_invstd[c] = invstd;
      if (running_mean != nullptr) {
        const double unbiased = M > 1 ? var * static_cast<double>(M) / static_cast<double>(M - 1) : var;
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
      }
    }
  }
  __syncthreads();
  if (!active) return;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { sc[e] = sc_s[g.cg * 8 + e]; sh[e] = sh_s[g.cg * 8 + e]; }
  const int64_t off = blockIdx.x * slab_groups + g.cg;
  const int total_sweeps = nbox * sweeps;
  for (int t0 = 0; t0 < total_sweeps; t0 += 4) {                       // four rows per thread in flight (the residual comes from global memory)
    uint4 ur[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t rr = row_lo + static_cast<int64_t>(t0 + i) * ri + my_row;
      ur[i] = (residual != nullptr && rr < row_hi) ? ldg_nc_u4(residual + rr * c8 + off) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t rr = row_lo + static_cast<int64_t>(t0 + i) * ri + my_row;
      if (rr < row_hi) {
        const uint4 ux = *reinterpret_cast<const uint4*>(my_cache + static_cast<size_t>(t0 + i) * ri * pitch);
        float f[8], rv[8];
        unpack8<BF16>(ux, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = fmaf(f[e], sc[e], sh[e]);
        if (residual != nullptr) {
          unpack8<BF16>(ur[i], rv);
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] += rv[e];
        }
        if (relu) {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
        }
        y[rr * c8 + off] = pack8<BF16>(f);
      }
    }
  }
}

template <bool BF16>
__global__ void __launch_bounds__(32 * kBnRedY)
bn_bwd_coop_kernel(const uint4* __restrict__ x, const uint4* __restrict__ y, const uint4* __restrict__ gy, const float* __restrict__ weight,
                   const float* __restrict__ save_mean, const float* __restrict__ save_invstd, float* __restrict__ partial, int* __restrict__ counters,
                   float* __restrict__ grad_weight, float* __restrict__ grad_bias, uint4* __restrict__ gx, uint4* __restrict__ gres, int64_t M, int C, int slab_groups) {
  __shared__ __align__(16) float sm[kBnRedY * 32 * 16];
  __shared__ float a_s[kBnSlab], b_s[kBnSlab], c_s[kBnSlab];
  const ColGeom g = col_geom(C, slab_groups);
  const bool active = static_cast<int>(threadIdx.x) < g.groups * g.rows_per_warp;
  const int c8 = C >> 3;
  const int64_t row_step = static_cast<int64_t>(gridDim.y) * kBnRedY * g.rows_per_warp;
  const int64_t r0 = (static_cast<int64_t>(blockIdx.y) * kBnRedY + threadIdx.y) * g.rows_per_warp + g.rsub;
  const int64_t off = blockIdx.x * slab_groups + g.cg;
  float sg[8], sx[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { sg[e] = 0.f; sx[e] = 0.f; }
  if (active) {
    const int cbase = (blockIdx.x * slab_groups + g.cg) * 8;
    float mu[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) mu[e] = __ldg(save_mean + cbase + e);
    for (int64_t r = r0; r < M; r += 4 * row_step) {
      uint4 ux[4], ug[4], uy[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t rr = r + i * row_step;
        const int64_t idx = (rr < M ? rr : r) * c8 + off;
        ux[i] = ldg_nc_u4(x + idx);
        ug[i] = ldg_nc_u4(gy + idx);
        uy[i] = (y != nullptr) ? ldg_nc_u4(y + idx) : make_uint4(0, 0, 0, 0);
        if (rr >= M) ug[i] = make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float fx[8], fg[8], fy[8];
        unpack8<BF16>(ux[i], fx);
        unpack8<BF16>(ug[i], fg);
        if (y != nullptr) {
          unpack8<BF16>(uy[i], fy);
#pragma unroll
          for (int e = 0; e < 8; ++e) fg[e] = fy[e] > 0.f ? fg[e] : 0.f;
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) { sg[e] += fg[e]; sx[e] = fmaf(fg[e], fx[e] - mu[e], sx[e]); }
      }
    }
  }
  block_col_reduce(sg, sx, g, active, sm);
  if (threadIdx.y == 0 && static_cast<int>(threadIdx.x) < g.groups) {
    float* p = partial + (static_cast<size_t>(blockIdx.y) * C + (blockIdx.x * slab_groups + threadIdx.x) * 8) * 2;
#pragma unroll
    for (int e = 0; e < 8; ++e) { p[2 * e] = sg[e]; p[2 * e + 1] = sx[e]; }
  }
  slab_barrier(counters + blockIdx.x, counters + kBnCounterSlots + blockIdx.x, gridDim.y);
  double S, Q;
  const int nch = g.groups * 8, c0 = blockIdx.x * slab_groups * 8;
  final_col_sum(partial, gridDim.y, C, c0, nch, reinterpret_cast<double*>(sm), S, Q);
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (tid < nch) {
    const int c = c0 + tid;
    const double is = static_cast<double>(save_invstd[c]), mean = static_cast<double>(save_mean[c]), w = static_cast<double>(weight[c]);
    const double sum_gx = Q * is;
    const double a = S / static_cast<double>(M), b = sum_gx / static_cast<double>(M), A = w * is;
    a_s[tid] = static_cast<float>(A);
    b_s[tid] = static_cast<float>(-A * is * b);
    c_s[tid] = static_cast<float>(A * (is * b * mean - a));
    if (blockIdx.y == 0) {
      grad_bias[c] = static_cast<float>(S);
      grad_weight[c] = static_cast<float>(sum_gx);
    }
  }
  __syncthreads();
  if (!active) return;
  float A[8], Bc[8], Cc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { A[e] = a_s[g.cg * 8 + e]; Bc[e] = b_s[g.cg * 8 + e]; Cc[e] = c_s[g.cg * 8 + e]; }
  for (int64_t r = r0; r < M; r += 2 * row_step) {             // second pass: L2 hits
    uint4 ux[2], ug[2], uy[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t rr = r + i * row_step;
      const int64_t idx = (rr < M ? rr : r) * c8 + off;
      ux[i] = ldg_nc_u4(x + idx);
      ug[i] = ldg_nc_u4(gy + idx);
      uy[i] = (y != nullptr) ? ldg_nc_u4(y + idx) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t rr = r + i * row_step;
      if (rr < M) {
        const int64_t idx = rr * c8 + off;
        float fx[8], fg[8], fy[8], o[8];
        unpack8<BF16>(ux[i], fx);
        unpack8<BF16>(ug[i], fg);
        if (y != nullptr) {
          unpack8<BF16>(uy[i], fy);
#pragma unroll
          for (int e = 0; e < 8; ++e) fg[e] = fy[e] > 0.f ? fg[e] : 0.f;
          if (gres != nullptr) gres[idx] = pack8<BF16>(fg);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = fmaf(A[e], fg[e], fmaf(Bc[e], fx[e], Cc[e]));
        gx[idx] = pack8<BF16>(o);
      }
    }
  }
}

// cooperative grids: at most one block per SM in total, and no more row blocks than the finalisation sums per thread.
// Narrow layers (C <= 512) get narrower channel slabs (8 or 16 lane groups instead of 32; the spare lanes of a warp take further
// rows) so that the grid still covers the machine: with 256-channel slabs a C = 256 layer would run on 32 SMs.
// SMs the single-launch kernels may cover (acan_bn_set_sm_limit): a cooperative grid only starts when ALL its blocks fit at once, so next
// to a long-running kernel that keeps some SMs to itself (an NCCL collective overlapping the backward pass) a 148-block grid waits for
// that kernel to END.  The multi-GPU training step leaves NCCL's SMs out of these grids instead.
std::atomic<int> g_bn_sm_limit{kNumSMs};

struct CoopGrid { int slab_groups, slabs, row_blocks; };
CoopGrid coop_grid(int64_t M, int C, int rows_per_thread) {
  const int sms = g_bn_sm_limit.load(std::memory_order_relaxed);
  CoopGrid best = {32, (C + kBnSlab - 1) / kBnSlab, 1};
  long best_blocks = -1;
  const int cands[3] = {32, 16, 8};
  for (int L : cands) {
    const int c8 = C / 8;
    const int slabs = (c8 + L - 1) / L;
    if (slabs > sms || slabs > kBnCounterSlots) continue;
    const int groups = c8 < L ? c8 : L;
    const int rpw = (32 % groups) == 0 ? 32 / groups : 1;
    const int64_t rows_per_block_min = static_cast<int64_t>(kBnRedY) * rpw * rows_per_thread;
    int64_t rb = sms / slabs;
    const int64_t max_rb = (M + rows_per_block_min - 1) / rows_per_block_min;
    const int nch = groups * 8;
    const int64_t fin = static_cast<int64_t>(kBnFinalDepth) * ((32 * kBnRedY) / nch);
    if (rb > max_rb) rb = max_rb;
    if (rb > fin) rb = fin;
    if (rb > kBnMaxRowBlocks) rb = kBnMaxRowBlocks;
    if (rb < 1) rb = 1;
    const long blocks = static_cast<long>(slabs) * rb;
    if (blocks > best_blocks) { best_blocks = blocks; best = {L, slabs, static_cast<int>(rb)}; }      // ties keep the wider slab
  }
  return best;
}

// ACAN_BN_PLAIN_LAUNCH=1 (builds with -DACAN_EXPERIMENT only): launch the single-pass kernels as ordinary grids (experiment).  The grid is still at most one block per SM, so
// every block becomes resident as soon as the kernels ahead of it in the stream drain (nothing that runs later can hold an SM against it)
// and the barrier's bounded spin still traps instead of hanging; an ordinary launch may start its first blocks under the previous
// kernel's tail, which a cooperative launch (whole grid resident at once) cannot.
bool bn_plain_launch() {
#ifdef ACAN_EXPERIMENT
  static int v = -1;
  if (v < 0) { const char* e = getenv("ACAN_BN_PLAIN_LAUNCH"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
#else
  return false;             // the shipped library never reads the environment
#endif
}

bool bn_use_coop(int C) {
  int v = 1;
#ifdef ACAN_EXPERIMENT
  { const char* e = getenv("ACAN_BN_COOP"); v = (e && e[0] == '0') ? 0 : 1; }
#endif
  return v == 1 && (C + kBnSlab - 1) / kBnSlab <= g_bn_sm_limit.load(std::memory_order_relaxed) && (C + kBnSlab - 1) / kBnSlab <= kBnCounterSlots;   // the 32-group candidate always fits then
}

// row blocks so that the whole grid is ~`per_sm` blocks per SM, with at least `rows_per_thread` rows for every thread of a block
// (block height `by`; a warp covers 32 / groups rows when the slab has fewer than 32 channel groups)
int row_blocks_for(int64_t M, int C, int per_sm, int by, int rows_per_thread, int cap) {
  const int slabs = (C + kBnSlab - 1) / kBnSlab;
  const int groups = (C / 8) < 32 ? (C / 8) : 32;
  const int rpw = (32 % groups) == 0 ? 32 / groups : 1;
  const int64_t rows_per_block_min = static_cast<int64_t>(by) * rpw * rows_per_thread;
  int64_t rb = (static_cast<int64_t>(kNumSMs) * per_sm + slabs - 1) / slabs;
  const int64_t max_rb = (M + rows_per_block_min - 1) / rows_per_block_min;
  if (rb > max_rb) rb = max_rb;
  if (cap == kBnMaxRowBlocks) {                      // reduction kernels: the finalising block sums kBnFinalDepth partials per thread
    const int nch = C < kBnSlab ? C : kBnSlab;
    const int fin = kBnFinalDepth * ((32 * kBnRedY) / nch);
    if (rb > fin) rb = fin;
  }
  if (rb > cap) rb = cap;
  if (rb < 1) rb = 1;
  return static_cast<int>(rb);
}

// [M, C] activations as a 2-D tensor map: boxes of kBnBoxRows rows x box_ch channels, dense in shared memory, rows / channels beyond
// the tensor read as zeros
typedef CUresult (*BnEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_bn_tmap(CUtensorMap* m, const void* base, bool bf16, int64_t M, int C, int box_ch) {
  static std::atomic<void*> fn_cache{nullptr};
  void* p = fn_cache.load(std::memory_order_acquire);
  if (!p) {
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
      (void)cudaGetLastError();
      return fail(ACAN_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
    }
    fn_cache.store(p, std::memory_order_release);
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(M)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(C) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_ch), static_cast<cuuint32_t>(kBnBoxRows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<BnEncodeTiledFn>(p)(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base),
                                                    dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ACAN_ERR_DRIVER, "cuTensorMapEncodeTiled (BN activations) failed with CUresult %d (M=%lld C=%d box=%d)", (int)r, (long long)M, C, box_ch);
  return 0;
}

int check_bn(const char* who, int64_t M, int C, int dtype) {
  if (M <= 0 || C <= 0 || (C % 8) != 0) return fail(ACAN_ERR_BAD_ARG, "%s: need M > 0 and C a positive multiple of 8 (M=%lld C=%d)", who, (long long)M, C);
  if (dtype != 0 && dtype != 1) return fail(ACAN_ERR_BAD_ARG, "%s: dtype %d (0 = fp16, 1 = bf16)", who, dtype);
  if ((C + kBnSlab - 1) / kBnSlab > kBnCounterSlots) return fail(ACAN_ERR_BAD_ARG, "%s: C=%d exceeds %d channels", who, C, kBnCounterSlots * kBnSlab);
  return 0;
}

}  // namespace
}  // namespace acan

using namespace acan;

// scratch layout (floats): [kBnCounterSlots arrival counters (ints; zero between calls) + kBnCounterSlots barrier generation words, at a
// C-independent place so that one zero-initialised buffer serves layers of every width] [row-block partials kBnMaxRowBlocks x C x 2]
// [(scale, shift) | (A, Bc, Cc): 3 x C]
extern "C" int acan_bn_set_sm_limit(int n_sms) {
  const int prev = g_bn_sm_limit.load(std::memory_order_relaxed);
  if (n_sms > 0) g_bn_sm_limit.store(n_sms < 32 ? 32 : (n_sms > kNumSMs ? kNumSMs : n_sms), std::memory_order_relaxed);
  return prev;
}

extern "C" size_t acan_bn_ws_floats(int C) {
  if (C <= 0) return 0;
  return 2 * static_cast<size_t>(kBnCounterSlots) + static_cast<size_t>(kBnMaxRowBlocks) * C * 2 + 3 * static_cast<size_t>(C);
}

extern "C" int acan_bn_train_fwd(const void* x, int dtype, const float* weight, const float* bias, const void* residual, void* y,
                                 float* save_mean, float* save_invstd, float* running_mean, float* running_var,
                                 float momentum, float eps, int relu, float* ws, int64_t M, int C, void* stream) {
  ACAN_CHECK_ARG(x && weight && bias && y && save_mean && save_invstd && ws, "acan_bn_train_fwd: null pointer");
  ACAN_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "acan_bn_train_fwd: running_mean and running_var go together");
  if (int e = check_bn("acan_bn_train_fwd", M, C, dtype)) return e;
  ACAN_CHECK_ARG(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0 && reinterpret_cast<uintptr_t>(residual) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(ws) % 16 == 0, "acan_bn_train_fwd: tensors must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* counters = reinterpret_cast<int*>(ws);
  float* partial = ws + 2 * kBnCounterSlots;
  float* scale_shift = partial + static_cast<size_t>(kBnMaxRowBlocks) * C * 2;
  const int slabs = (C + kBnSlab - 1) / kBnSlab;
  const dim3 rgrid(slabs, row_blocks_for(M, C, 2, kBnRedY, 8, kBnMaxRowBlocks)), rblock(32, kBnRedY);
  const dim3 agrid(slabs, row_blocks_for(M, C, 8, kBnAppY, 2, 65535)), ablock(32, kBnAppY);
  const uint4* xr = static_cast<const uint4*>(x);
  const uint4* rr = static_cast<const uint4*>(residual);
  if (bn_use_coop(C)) {
    CoopGrid cg = coop_grid(M, C, 8);
    uint4* yw = static_cast<uint4*>(y);
    {  // the layer fits the shared memory of the grid: stage it there by TMA and read x once (bn_fwd_tma_kernel)
      int sg = cg.slab_groups < C / 8 ? cg.slab_groups : C / 8;
      const int pitch = sg * 16;
      int64_t R = (M + cg.row_blocks - 1) / cg.row_blocks;
      R = (R + kBnBoxRows - 1) / kBnBoxRows * kBnBoxRows;
      if (R * pitch <= kBnTmaSmemMax && R / kBnBoxRows <= kBnMaxBoxes && M < (1ll << 31) - 65536) {
        CUtensorMap tm;
        if (int e = make_bn_tmap(&tm, x, dtype == 1, M, C, sg * 8)) return e;
        int rows_per_block = static_cast<int>(R);
        const dim3 tgrid(cg.slabs, static_cast<unsigned>((M + R - 1) / R));
        const size_t smem = static_cast<size_t>(R) * pitch + 1024;
        const void* fn = dtype == 1 ? reinterpret_cast<const void*>(bn_fwd_tma_kernel<true>) : reinterpret_cast<const void*>(bn_fwd_tma_kernel<false>);
        static std::atomic<uint64_t> done[2];
        int dev = 0;
        ACAN_CUDA(cudaGetDevice(&dev));
        const uint64_t bit = 1ull << (dev & 63);
        if (!(done[dtype].load(std::memory_order_acquire) & bit)) {
          ACAN_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kBnTmaSmemMax + 1024));
          done[dtype].fetch_or(bit, std::memory_order_release);
        }
        void* args[] = {&tm, &rr, &yw, &weight, &bias, &partial, &counters, &save_mean, &save_invstd, &running_mean, &running_var,
                        &momentum, &eps, &relu, &M, &C, &sg, &rows_per_block};
        ACAN_CUDA(cudaLaunchCooperativeKernel(fn, tgrid, rblock, args, smem, st));
        return launch_status("bn_fwd_tma_kernel");
      }
    }
    const dim3 cgrid(cg.slabs, cg.row_blocks);
    void* args[] = {&xr, &rr, &yw, &weight, &bias, &partial, &counters, &save_mean, &save_invstd, &running_mean, &running_var,
                    &momentum, &eps, &relu, &M, &C, &cg.slab_groups};
    const void* fn = dtype == 1 ? reinterpret_cast<const void*>(bn_fwd_coop_kernel<true>) : reinterpret_cast<const void*>(bn_fwd_coop_kernel<false>);
    ACAN_CUDA(bn_plain_launch() ? cudaLaunchKernel(fn, cgrid, rblock, args, 0, st) : cudaLaunchCooperativeKernel(fn, cgrid, rblock, args, 0, st));
    return launch_status("bn_fwd_coop_kernel");
  }
  if (dtype == 1) {
    bn_stats_kernel<true><<<rgrid, rblock, 0, st>>>(xr, weight, bias, partial, counters, save_mean, save_invstd, scale_shift, running_mean, running_var,
                                                    momentum, eps, M, C);
    if (int e = launch_status("bn_stats_kernel")) return e;
    bn_apply_kernel<true><<<agrid, ablock, 0, st>>>(xr, rr, static_cast<uint4*>(y), scale_shift, relu, M, C);
  } else {
    bn_stats_kernel<false><<<rgrid, rblock, 0, st>>>(xr, weight, bias, partial, counters, save_mean, save_invstd, scale_shift, running_mean, running_var,
                                                     momentum, eps, M, C);
    if (int e = launch_status("bn_stats_kernel")) return e;
    bn_apply_kernel<false><<<agrid, ablock, 0, st>>>(xr, rr, static_cast<uint4*>(y), scale_shift, relu, M, C);
  }
  return launch_status("bn_apply_kernel");
}

extern "C" int acan_bn_train_bwd(const void* x, const void* y, const void* grad_y, int dtype, const float* weight,
                                 const float* save_mean, const float* save_invstd, void* grad_x, void* grad_residual,
                                 float* grad_weight, float* grad_bias, float* ws, int64_t M, int C, void* stream) {
  ACAN_CHECK_ARG(x && grad_y && weight && save_mean && save_invstd && grad_x && grad_weight && grad_bias && ws, "acan_bn_train_bwd: null pointer");
  ACAN_CHECK_ARG(grad_residual == nullptr || y != nullptr, "acan_bn_train_bwd: grad_residual needs y (the ReLU mask)");
  if (int e = check_bn("acan_bn_train_bwd", M, C, dtype)) return e;
  ACAN_CHECK_ARG(reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0 && reinterpret_cast<uintptr_t>(grad_y) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(grad_x) % 16 == 0 && reinterpret_cast<uintptr_t>(grad_residual) % 16 == 0 && reinterpret_cast<uintptr_t>(ws) % 16 == 0,
                 "acan_bn_train_bwd: tensors must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int* counters = reinterpret_cast<int*>(ws);
  float* partial = ws + 2 * kBnCounterSlots;
  float* coef = partial + static_cast<size_t>(kBnMaxRowBlocks) * C * 2;
  const int slabs = (C + kBnSlab - 1) / kBnSlab;
  const dim3 rgrid(slabs, row_blocks_for(M, C, 1, kBnRedY, 4, kBnMaxRowBlocks)), rblock(32, kBnRedY);   // 94 registers x 512 threads: one block per SM
  const dim3 agrid(slabs, row_blocks_for(M, C, 8, kBnAppY, 2, 65535)), ablock(32, kBnAppY);
  const uint4* xr = static_cast<const uint4*>(x);
  const uint4* yr = static_cast<const uint4*>(y);
  const uint4* gr = static_cast<const uint4*>(grad_y);
  if (bn_use_coop(C)) {
    CoopGrid cg = coop_grid(M, C, 4);
    const dim3 cgrid(cg.slabs, cg.row_blocks);
    uint4* gxw = static_cast<uint4*>(grad_x);
    uint4* grw = static_cast<uint4*>(grad_residual);
    void* args[] = {&xr, &yr, &gr, &weight, &save_mean, &save_invstd, &partial, &counters, &grad_weight, &grad_bias, &gxw, &grw, &M, &C, &cg.slab_groups};
    const void* fn = dtype == 1 ? reinterpret_cast<const void*>(bn_bwd_coop_kernel<true>) : reinterpret_cast<const void*>(bn_bwd_coop_kernel<false>);
    ACAN_CUDA(bn_plain_launch() ? cudaLaunchKernel(fn, cgrid, rblock, args, 0, st) : cudaLaunchCooperativeKernel(fn, cgrid, rblock, args, 0, st));
    return launch_status("bn_bwd_coop_kernel");
  }
  if (dtype == 1) {
    bn_bwd_reduce_kernel<true><<<rgrid, rblock, 0, st>>>(xr, yr, gr, weight, save_mean, save_invstd, partial, counters, grad_weight, grad_bias, coef, M, C);
    if (int e = launch_status("bn_bwd_reduce_kernel")) return e;
    bn_bwd_apply_kernel<true><<<agrid, ablock, 0, st>>>(xr, yr, gr, coef, static_cast<uint4*>(grad_x), static_cast<uint4*>(grad_residual), M, C);
  } else {
    bn_bwd_reduce_kernel<false><<<rgrid, rblock, 0, st>>>(xr, yr, gr, weight, save_mean, save_invstd, partial, counters, grad_weight, grad_bias, coef, M, C);
    if (int e = launch_status("bn_bwd_reduce_kernel")) return e;
    bn_bwd_apply_kernel<false><<<agrid, ablock, 0, st>>>(xr, yr, gr, coef, static_cast<uint4*>(grad_x), static_cast<uint4*>(grad_residual), M, C);
  }
  return launch_status("bn_bwd_apply_kernel");
}
