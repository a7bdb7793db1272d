// attention.cu -- pixel-to-pixel affinity, row softmax and aggregation on tcgen05 tensor cores.
//
// Replaces PixelAttentionBlock_.forward_att (reference sadecoder.py:42-49) and the matmul / permute of
// SelfAttentionBlock_.forward (sadecoder.py:85-87), with Q = K = F because f_query is f_key (:30):
//     S = F^T F * dk^-1/2        [N,N] per image
//     P = softmax_rows(S)        (= sim_map)
//     O[c,i] = sum_j V[c,j] P[i,j]    written directly in NCHW ([B,dv,N]), no transpose pass
// and, fused into the same tensor-core pass, the attention loss of losses.py:12-62.
//
// Why three tensor-core passes instead of one flash-style kernel: the aggregation accumulator for a
// 128-query tile is 128 x dv = 128 x 2048 fp32 = 1 MB, four times the 256 KB of TMEM an SM has, so a single
// kernel would recompute S once per dv slice (>= 5x).  Instead
//   pass 1  STATS : S tiles -> per-row (max, sum-exp) partials -> combined to (m_i, l_i)       2*N^2*dk  FLOP
//   pass 2  PROBS : S tiles again -> P = exp2(S*c - m_i)/l_i, final, fp16, written once          2*N^2*dk  FLOP
//                   (+ optional fp32 sim_map, + optional fused KL against the on-the-fly target)
//   pass 3  PV    : O^T = V P^T as a plain TMA/UMMA GEMM, accumulators never rescaled           2*N^2*dv  FLOP
// P (2*N^2 bytes / image, 4 MB at N = 1408) round-trips through the 126 MB L2.  Executed / algorithmic
// FLOPs = (2*dk + dv) / (dk + dv) = 1.2.  Exact row statistics up front are also what the KL clamp
//  and the backward pass need.
//
// All passes (forward and backward) are one warp-specialised kernel template (gemm_kernel<MODE, CG, AM, BMN>), 320 threads:
//   warp 0   : TMA producer   (cp.async.bulk.tensor.3d, SWIZZLE_128B, mbarrier complete_tx; K-major or MN-major operand boxes)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, M = 128 per CTA, N = BN <= 256, K = 16)
//   warps 2-9: epilogue, two groups of four warps (tcgen05.ld 32x32b.x32 -> registers -> SWIZZLE_128B staging in shared
//              memory -> TMA store), TMEM double-buffered so the epilogue of tile t overlaps the MMAs of tile t+1
// persistent over a static tile schedule; CG = 2 (default) runs CTA pairs (cluster of 2, cta_group::2: a 256 x BN tile per pair).
// The probabilities are fp16 operands (11-bit significand: the rounding of TF32, at the bf16 MMA rate) with fp32 accumulation;
// bf16 probabilities miss the 1e-3 parity bar on sim_map.  V and dO therefore have to be fp16 too: the instruction
// descriptor has one format field per operand, but an fp16 x bf16 kind::f16 MMA traps as an illegal instruction on sm_100a (measured).
#include "common.cuh"
#include <cuda_bf16.h>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <atomic>

namespace acan {

// from kl.cu
__global__ void nearest_bins_kernel(const float* __restrict__ label, float* __restrict__ bins, int B, int H, int W);
__global__ void ordered_sum_kernel(const float* __restrict__ vals, int64_t n, float scale, float* __restrict__ out);

namespace {

constexpr int BM = 128;                 // UMMA M (TMEM lanes)
constexpr int BK = 64;                  // K elements per pipeline stage: 64 fp16 = one 128 B swizzle row
constexpr int UMMA_K = 16;
constexpr int kMaxBN = 256;
constexpr int kMinBN = 64;
constexpr int kTmemCols = 512;          // 2 accumulator buffers x 256 fp32 columns
constexpr int kGemmThreads = 320;       // TMA warp + MMA warp + 2 x 4 epilogue warps
constexpr int kMaxStages = 8;
constexpr int kSmemBudget = 225 * 1024; // operand ring + epilogue staging; barriers + alignment slack (~1.2 KB) come on top (227 KB max)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kShiftHeadroom = 14.0f, kSafeDiagMax = 72.0f;   // statistics-free probability pass, see diag_kernel
constexpr float kKeepScale = 1024.0f;   // training path: normalised probabilities are stored as fp16(P * 2^10) (tails stay normal numbers)

// Experiment knobs (mainloop-only timing, per-tile traces, tile-width / launch-mode overrides from the environment) exist only
// in builds made with -DACAN_EXPERIMENT: the shipped library never reads the environment and never skips an epilogue.
#ifdef ACAN_EXPERIMENT
constexpr bool kExperiment = true;
#else
constexpr bool kExperiment = false;
#endif

// MODE_PV : O^T[dv, Nq] = V[dv, Nk] P[Nq, Nk]^T, both K-major (NCHW world), fp32 out.
// MODE_PVT: O[Nq, dv]  = P[Nq, Nk] V[Nk, dv], B = V read in place from channels-last memory as an MN-major operand, fp16 out.
// MODE_DS : backward: acc = dP[Nq, Nk] = dO^T V; epilogue dS = P o (dP - delta_i) (+ attention-loss term), scaled fp16 out.
// MODE_PVT32: as MODE_PVT (accumulator rows = rows of A, row-major [rows, cols] output) with fp32 output.
enum Mode { MODE_STATS = 0, MODE_PROBS = 1, MODE_PV = 2, MODE_PVT = 3, MODE_DS = 4, MODE_PVT32 = 5 };
enum DType { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2 };

constexpr int kOutChunkBytes = 128 * 128;   // epilogue staging: 128 rows x 128 B (one SWIZZLE_128B TMA-store box)
constexpr int kDmaxSlots = 64;          // images whose diagonal maximum a probability-pass CTA caches in shared memory
constexpr int kOutBuffers = 2;          // TMA-store staging ring (4 buffers measured slower: they cost an operand stage)

struct GemmArgs {
  int batch, m_tiles, n_tiles, k_steps;   // m_tiles counts tiles of 128*CG rows
  int m_rows;          // valid rows of A per image (N queries, or dv channels)
  int n_cols;          // valid rows of B per image (N keys, or N queries)
  int bn;              // tile width (multiple of 32, of 64 when P is stored; <= 256)
  int stages;
  uint32_t idesc;
  float c;             // scale * log2(e)
  int k_half;          // AM == 2 (symmetrised A): k-steps [0, k_half) read A K-major through tmap_a, [k_half, k_steps) read the
                       //   same matrix transposed (MN-major boxes through tmap_a2); B k-coordinates restart at k_half
  uint32_t idesc2;     //   instruction descriptor of the second half (a_major = MN)
  int skip_epilogue;   // experiments only (ACAN_SKIP_EPI=1): release the accumulator without reading it
  const float* dpart;  // fast path: [B, n_dpart] per-CTA maxima of the diagonal d_i = S_ii * c (diag_kernel); an image is "safe" when its
  int n_dpart;         //   maximum is <= kSafeDiagMax.  MODE_STATS skips the tiles of safe images (and exits at once when all are).
  const float* diag;   // MODE_PROBS (unnorm): [B, N] d_i, the row shift is derived from it in the epilogue
  float2* stats_w;     // fast path: row_stats being produced -- PROBS writes .x (shift_i), PVT / combine_l write .y (l_i)
  float2* stats_w2;    //   second copy kept in the workspace for acan_attn_rows / acan_attention_loss_fused (may be null)
  const float* l_in;   // MODE_PVT: [B, l_parts, N] partial row sums of p~ written by the PROBS pass; 1 / sum is applied per query
  int l_parts;
  int unnorm;          // MODE_PROBS: write p~ = exp2(S*c - shift_i) un-normalised and emit per-row partial sums
  float* lpart;        //   [B, 2*n_tiles, N] those partial sums
  const float* ds_delta; // MODE_DS: [B, N] delta_i = sum_c dO[c,i] O[c,i]
  const __half* ds_p;    //   [B, N, N] fp16 probabilities
  const float* ds_klT;   //   [B, N] in-window target mass T_i (attention-loss term), with logbin / lnz above
  const float* ds_gl;    //   device scalar dLoss/dloss3 (null = no attention-loss term);  gk = *ds_gl * ds_inv_bn
  float ds_inv_bn;       //   1 / (B N)
  const float* gscale; // backward: device pair (s, 1/s), the power-of-two gradient scale that keeps fp16 dO / dS in range
  const float* inv_l;  // MODE_PV / MODE_PVT: [B, N] 1 / l_i applied per query while the accumulator leaves TMEM (null = already normalised)
  float out_scale;     // MODE_PVT / MODE_PVT32: constant factor on the output (host sets 1)
  const float* out_gscale;   //   device scalar multiplied in as well (null = 1)
  int out_bf16;        // MODE_PVT / MODE_DS: 16-bit output is bf16 instead of fp16
  float* lsum_out;     // MODE_PVT / MODE_PVT32 with l_in: [B, N] the row sums go here (training path: the divisor of the stored probabilities)
  float p_scale;       // MODE_PROBS (normalised): stored probabilities = P * p_scale (1, or kKeepScale on the training path)
  const float* ds_lsum;  // MODE_DS (channels-last backward): [B, N] sum_j of the stored fp16 probabilities; p_ij = ds_p_ij / ds_lsum_i
  int ds_new;          //   1: scale factors come from gscale[0..4] = (s_ds, 1/s_ds, s_ds/s_op, 1/s_op, delta_mul)
  int pdl;             // host only: launch as a programmatic dependent of the previous kernel in the stream (one of ours, see griddep_wait)
  long long* trace;    // experiments only (ACAN_TRACE=1): per-tile clock64 stamps of CTA 0: [tile][4] = mma start, mma issued, epi start, epi end
  // MODE_STATS
  float2* part;        // [B, n_tiles, m_rows] (m2, l) partials
  // MODE_PROBS
  const float2* stats; // [B, N] final (m2, l)
  int store_p;         // P (fp16) goes out through tmap_out
  float* sim_out;      // [B, N, N] fp32 or null (rare path: plain vector stores)
  const float* logbin; // [B, N] ln(bin) (-inf for bin 0) or null  -> fused KL
  const float* lnz;    // [B, N] ln(Z_i) of the target affinity (unused when bin_i == 0)
  float* kl_part;      // [B, 2*n_tiles, N] per-row partial losses
  float* klw_part;     // null, or [B, 2*n_tiles, N] per-row partial sums of the in-window target mass T_i (backward)
};

struct __align__(8) PipeBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
  float dmax[kDmaxSlots];   // probability pass, inference path: per-image maxima of the diagonal
};

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ uint32_t pack_bf162(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack16(float a, float b, bool bf16) { return bf16 ? pack_bf162(a, b) : pack_half2(a, b); }

// 16-byte chunk j (0..7) of row r (0..127) inside a 128 B-per-row SWIZZLE_128B staging buffer
__device__ __forceinline__ uint4* staged_chunk(uint8_t* buf, int r, int j) {
  return reinterpret_cast<uint4*>(buf + r * 128 + ((j ^ (r & 7)) << 4));
}

// Sum of the two fp16 values packed in `u`, WITHOUT conversion instructions: the 15 magnitude bits of each half are placed where an fp32
// keeps its exponent / mantissa, which reads as value * 2^-112 (normal and subnormal halves alike; the probabilities are >= 0 so there is
// no sign to carry).  The caller multiplies the accumulated sum by 2^112 once.  Two logic ops + one add per element on the full-rate
// integer / FMA pipes -- the half -> float conversions of the obvious version share the quarter-rate pipe with the ex2 of the same
// epilogue and made the probability pass epilogue-bound (profiles/r02_notes.md).
__device__ __forceinline__ float half2_sum_scaled(uint32_t u) {
  const float lo = __uint_as_float((u << 13) & 0x0fffe000u);
  const float hi = __uint_as_float((u >> 3) & 0x0fffe000u);
  return lo + hi;
}
constexpr float kHalfSumRescale = 5192296858534827628530496329220096.0f;   // 2^112

// One 32-column half of a 64-column staged chunk of the un-normalised probability pass: p~ = exp2(S*c - shift_i) -> fp16 ->
// swizzled staging row; returns the sum of the ROUNDED values (what the aggregation pass will read: measurably more accurate
// than summing the fp32 values).  Columns >= nv lie beyond N: p~ = 0.
__device__ __forceinline__ float probs_unnorm_half(float (&v)[32], int nv, float c, float m2, uint8_t* buf, int row_in_tile, int half) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], c, -m2);
  if (nv < 32) {
#pragma unroll
    for (int j = 0; j < 32; ++j) if (j >= nv) v[j] = -INFINITY;
  }
  uint32_t h[16];
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    h[j] = pack_half2(ex2_approx(v[2 * j]), ex2_approx(v[2 * j + 1]));
    h[j + 1] = pack_half2(ex2_approx(v[2 * j + 2]), ex2_approx(v[2 * j + 3]));
    s0 += half2_sum_scaled(h[j]);
    s1 += half2_sum_scaled(h[j + 1]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
    *staged_chunk(buf, row_in_tile, half * 4 + q) = make_uint4(h[q * 4 + 0], h[q * 4 + 1], h[q * 4 + 2], h[q * 4 + 3]);
  return (s0 + s1) * kHalfSumRescale;
}

// max_i d_i of image b from the per-CTA maxima diag_kernel left (<= 13 values for N <= 1664)
__device__ __forceinline__ float image_dmax(const float* __restrict__ dpart, int n_dpart, int b) {
  float m = 0.f;
  for (int t = 0; t < n_dpart; ++t) m = fmaxf(m, __ldg(dpart + static_cast<size_t>(b) * n_dpart + t));
  return m;
}

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile --
// each CTA stages its own 128 rows of A and HALF of the B tile, the leader issues M = 256 MMAs that read both
// shared memories and write both TMEMs, so the L2 -> SM operand traffic per FLOP drops by (128+BN)/(128+BN/2).
// AM = 1 / BMN: that operand is read as an MN-major tile (64-element-wide {MN x 64 K-row} TMA boxes, 8 KB apart).
// AM = 2: A is a square matrix that enters the product symmetrised, acc = (A + A^T) B: the k-loop runs over it twice, first
// K-major (tmap_a) then MN-major (tmap_a2), both halves into the same accumulator (the two dF GEMMs of the backward pass).
template <int MODE, int CG, int AM, bool BMN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_out, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const uint32_t a_bytes = BM * BK * 2;
  const uint32_t b_rows = static_cast<uint32_t>(g.bn) / CG;          // rows of the B tile this CTA stages
  const uint32_t b_bytes = b_rows * BK * 2;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* out_stage = smem + static_cast<size_t>(g.stages) * stage_bytes;
  PipeBarriers* bars = reinterpret_cast<PipeBarriers*>(out_stage + kOutBuffers * kOutChunkBytes);
  float* const dmax_s = bars->dmax;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  griddep_launch_dependents();       // the next pass may start its prologue as soon as this grid's CTAs are resident
  if (MODE == MODE_STATS && g.dpart != nullptr) {
    griddep_wait();
    // fallback statistics pass of the inference path: nothing to do unless some image's diagonal exceeds the safe bound
    // (every CTA of the grid, and both CTAs of a pair, reach the same verdict from the same values)
    float m = 0.f;
    for (int i = threadIdx.x; i < g.batch * g.n_dpart; i += kGemmThreads) m = fmaxf(m, __ldg(g.dpart + i));
    if (!__syncthreads_or(m > kSafeDiagMax)) return;
  }
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int total_tiles = g.batch * g.m_tiles * g.n_tiles;
  const int first_tile = blockIdx.x / CG, tile_step = gridDim.x / CG;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    if (AM == 2) tma_prefetch_desc(&tmap_a2);
    tma_prefetch_desc(&tmap_b);
    if (MODE != MODE_STATS) tma_prefetch_desc(&tmap_out);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < g.stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(&bars->tmem_full[s], 1); mbar_init(&bars->tmem_empty[s], 8 * CG); }
      fence_barrier_init();
    }
    __syncwarp();
    if (CG == 2) tmem_alloc_2sm(&bars->tmem_base, kTmemCols); else tmem_alloc(&bars->tmem_base, kTmemCols);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();      // barrier inits + TMEM address visible (pair-wide)
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  griddep_wait();                    // PDL: everything above overlapped the previous kernel's tail; its results are visible from here

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one thread per CTA)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int nt = tile % g.n_tiles;
        const int mt = (tile / g.n_tiles) % g.m_tiles;
        const int b = tile / (g.n_tiles * g.m_tiles);
        const int a_row = (mt * CG + static_cast<int>(cta_rank)) * BM;
        const int b_row = nt * g.bn + static_cast<int>(cta_rank * b_rows);
        if (MODE == MODE_STATS && g.dpart != nullptr && image_dmax(g.dpart, g.n_dpart, b) <= kSafeDiagMax) continue;
        for (int ks = 0; ks < g.k_steps; ++ks) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
          const bool a_mn = (AM == 1) || (AM == 2 && ks >= g.k_half);      // this k-step reads A as an MN-major tile
          const int kk = (AM == 2 && ks >= g.k_half) ? ks - g.k_half : ks;  // k coordinate (both halves of AM == 2 walk the same k range)
          uint32_t leader_bar = 0;
          if (CG == 1) {
            mbar_expect_tx(&bars->full[stage], stage_bytes);
          } else {
            // both CTAs' bytes complete on the LEADER's barrier; the leader arms it for the pair's total
            if (cta_rank == 0) mbar_expect_tx(&bars->full[stage], stage_bytes * 2);
            leader_bar = mapa_u32(smem_u32(&bars->full[stage]), 0);
          }
          auto ld = [&](void* dst, const CUtensorMap* m, int c0, int c1) {
            if (CG == 1) tma_load_3d(dst, m, &bars->full[stage], c0, c1, b);
            else         tma_load_3d_2sm(dst, m, leader_bar, c0, c1, b);
          };
          if (a_mn) {                            // MN-major A: two {64 rows-of-A x 64 k} boxes
            const CUtensorMap* ma = (AM == 2) ? &tmap_a2 : &tmap_a;
            for (uint32_t s = 0; s < BM / 64; ++s) ld(sa + s * 8192, ma, a_row + static_cast<int>(s) * 64, kk * BK);
          } else {
            ld(sa, &tmap_a, kk * BK, a_row);
          }
          if constexpr (BMN) {                   // MN-major B: one {64 rows-of-B x 64 k} box per 64 rows
            for (uint32_t s = 0; s < b_rows / 64; ++s) ld(sa + a_bytes + s * 8192, &tmap_b, b_row + static_cast<int>(s) * 64, kk * BK);
          } else {
            ld(sa + a_bytes, &tmap_b, kk * BK, b_row);
          }
          if (++stage == static_cast<uint32_t>(g.stages)) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread; leader CTA of a pair)
    if (lane == 0 && cta_rank == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        if (MODE == MODE_STATS && g.dpart != nullptr && image_dmax(g.dpart, g.n_dpart, tile / (g.n_tiles * g.m_tiles)) <= kSafeDiagMax) continue;
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        if (kExperiment && g.trace && blockIdx.x == 0) g.trace[(tile / tile_step) * 4 + 0] = clock64();
        const uint32_t d_tmem = tmem_base + acc * kMaxBN;
        for (int ks = 0; ks < g.k_steps; ++ks) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
          const uint32_t sb = sa + a_bytes;
          const bool a_mn = (AM == 1) || (AM == 2 && ks >= g.k_half);
          const uint32_t idesc = (AM == 2 && ks >= g.k_half) ? g.idesc2 : g.idesc;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = a_mn ? umma_desc_mn_sw128(sa + k * UMMA_K * 128, 8192) : umma_desc_k_sw128(sa + k * UMMA_K * 2);
            const uint64_t db = BMN ? umma_desc_mn_sw128(sb + k * UMMA_K * 128, 8192) : umma_desc_k_sw128(sb + k * UMMA_K * 2);
            if (CG == 1) umma_f16(d_tmem, da, db, idesc, (ks | k) != 0 ? 1u : 0u);
            else         umma_f16_2sm(d_tmem, da, db, idesc, (ks | k) != 0 ? 1u : 0u);
          }
          // smem slot free (in both CTAs of a pair) once these MMAs have read it
          if (CG == 1) umma_commit(&bars->empty[stage]); else umma_commit_2sm(&bars->empty[stage], 3);
          if (++stage == static_cast<uint32_t>(g.stages)) { stage = 0; phase ^= 1; }
        }
        if (CG == 1) umma_commit(&bars->tmem_full[acc]); else umma_commit_2sm(&bars->tmem_full[acc], 3);   // -> epilogue(s)
        if (kExperiment && g.trace && blockIdx.x == 0) g.trace[(tile / tile_step) * 4 + 1] = clock64();
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue: 2 groups x 4 warps
    // Each group covers the 128 TMEM lanes (warp % 4 = lane quarter) and owns alternate column chunks of the tile, its
    // own staging buffer and its own named barrier, so the two groups run decoupled and two warps share every SM
    // sub-partition: one group's tcgen05.ld / MUFU latency hides under the other's issue (the 4-warp version measured
    // ~1140 cycles of exposed latency per 64-column chunk and made the QK passes epilogue-bound).
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int group = (warp - 2) >> 2;                  // 0: warps 2-5, 1: warps 6-9
    const int row_in_tile = quarter * 32 + lane;
    const bool store_leader = ((threadIdx.x - 64) & 127) == 0;   // lane 0 of each group's first warp issues its TMA stores
    uint8_t* const buf = out_stage + group * kOutChunkBytes;
    const int bar_id = 1 + group;
    uint32_t acc = 0, acc_phase = 0;
    if (MODE == MODE_PROBS && g.unnorm == 1) {            // per-image diagonal maxima, once per CTA (epilogue warps only)
      const int t = threadIdx.x - 64;
      if (t < g.batch && t < kDmaxSlots) dmax_s[t] = image_dmax(g.dpart, g.n_dpart, t);
      named_bar_sync(3, 256);
    }
    const uint32_t leader_empty0 = (CG == 2) ? mapa_u32(smem_u32(&bars->tmem_empty[0]), 0) : 0u;
    const uint32_t leader_empty1 = (CG == 2) ? mapa_u32(smem_u32(&bars->tmem_empty[1]), 0) : 0u;
    for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
      const int nt = tile % g.n_tiles;
      const int mt = (tile / g.n_tiles) % g.m_tiles;
      const int b = tile / (g.n_tiles * g.m_tiles);
      const int row0 = (mt * CG + static_cast<int>(cta_rank)) * BM;
      const int row = row0 + row_in_tile;
      const bool row_ok = row < g.m_rows;
      const int col0 = nt * g.bn;
      const int nvalid = min(g.bn, g.n_cols - col0);   // >= 1 by construction of n_tiles
      const int chunks = (nvalid + 31) >> 5;
      const size_t part_idx = (static_cast<size_t>(b) * (2 * g.n_tiles) + 2 * nt + group) * g.m_rows + row;   // per (key tile, group) partials
      if (MODE == MODE_STATS && g.dpart != nullptr && image_dmax(g.dpart, g.n_dpart, b) <= kSafeDiagMax) continue;

      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      if (kExperiment && g.trace && blockIdx.x == 0 && threadIdx.x == 64) g.trace[(tile / tile_step) * 4 + 2] = clock64();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kMaxBN;

      if (kExperiment && g.skip_epilogue == 1) {
        // nothing: mainloop-only timing experiment
      } else if constexpr (MODE == MODE_STATS) {
        float m = -INFINITY, l = 0.f;                   // m in raw dot-product units
        for (int ch = group; ch < chunks; ch += 2) {
          float v[32];
          tmem_ld_32x32(taddr + ch * 32, v);
          tmem_ld_wait();
          const int nv = nvalid - ch * 32;
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) { if (j >= nv) v[j] = -INFINITY; cm = fmaxf(cm, v[j]); }
          if (cm > m) { l *= ex2_approx((m - cm) * g.c); m = cm; }
          const float mc = m * g.c;
          float s = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) s += ex2_approx(fmaf(v[j], g.c, -mc));
          l += s;
        }
        if (row_ok) g.part[part_idx] = make_float2(m * g.c, l);        // (-inf, 0) when this group had no chunk
      } else if constexpr (MODE == MODE_PROBS) {
        float m2 = 0.f, inv_l = 0.f, la_i = 0.f, lnz_i = 0.f, log2_l = 0.f;
        const size_t brow = static_cast<size_t>(b) * g.m_rows + (row_ok ? row : 0);
        if (g.unnorm == 2) {
          // training path: exact statistics are known; folding the normalisation into the shift, p^ = exp2(S c - (m + log2 l - log2 p_scale)),
          // gives fp16(P * p_scale) through the same light epilogue as the inference path (no per-element multiply, no second branch)
          inv_l = 1.0f;
          const float2 st = g.stats[brow];
          m2 = st.x + lg2_approx(st.y) - lg2_approx(g.p_scale);
        } else if (g.unnorm) {
          // Cauchy-Schwarz shift (see diag_kernel); images above the safe bound take the exact row maximum the fallback
          // statistics pass left in `part` (same tile width, so the same 2 * n_tiles partials per row)
          inv_l = 1.0f;
          const float dmax = b < kDmaxSlots ? dmax_s[b] : image_dmax(g.dpart, g.n_dpart, b);
          if (dmax <= kSafeDiagMax) {
            m2 = sqrtf(g.diag[brow] * dmax) - kShiftHeadroom;
          } else {
            m2 = -INFINITY;
            const float2* pp = g.part + static_cast<size_t>(b) * (2 * g.n_tiles) * g.m_rows + (row_ok ? row : 0);
            for (int t = 0; t < 2 * g.n_tiles; ++t) m2 = fmaxf(m2, pp[static_cast<size_t>(t) * g.m_rows].x);
          }
          if (row_ok && nt == 0 && group == 0) {
            g.stats_w[brow].x = m2;
            if (g.stats_w2) g.stats_w2[brow].x = m2;
          }
        } else {
          const float2 st = g.stats[brow];
          m2 = st.x;
          inv_l = g.p_scale / st.y;
          log2_l = lg2_approx(st.y);
        }
        float l_acc = 0.f;                               // un-normalised mode: sum of the fp16-rounded p~ this group wrote
        const bool kl = g.logbin != nullptr;
        bool kl_row = false;
        if (kl) { la_i = g.logbin[brow]; lnz_i = g.lnz[brow]; kl_row = row_ok && la_i > -INFINITY; }
        float kl_acc = 0.f, klw_acc = 0.f;
        const float* la_cols = kl ? g.logbin + static_cast<size_t>(b) * g.n_cols + col0 : nullptr;
        float* srow = g.sim_out ? g.sim_out + (brow * g.n_cols + col0) : nullptr;
        if (g.unnorm) {
          // inference path: both halves of a 64-column chunk are fetched from TMEM before the staging buffer is even free
          for (int ch = 2 * group; ch < chunks; ch += 4) {
            float v0[32], v1[32];
            const bool two = ch + 1 < chunks;                 // warp-uniform (an odd tail half lies beyond N: the store clips it)
            tmem_ld_32x32(taddr + ch * 32, v0);
            if (two) tmem_ld_32x32(taddr + (ch + 1) * 32, v1);
            if (store_leader) tma_store_wait_read<0>();       // this group's previous store has drained `buf`
            tmem_ld_wait();
            named_bar_sync(bar_id, 128);
            l_acc += probs_unnorm_half(v0, nvalid - ch * 32, g.c, m2, buf, row_in_tile, 0);
            if (two) l_acc += probs_unnorm_half(v1, nvalid - (ch + 1) * 32, g.c, m2, buf, row_in_tile, 1);
            fence_proxy_async();                              // generic-proxy smem writes -> visible to the TMA engine
            named_bar_sync(bar_id, 128);
            if (store_leader && !(kExperiment && g.skip_epilogue == 2)) {       // rows / columns beyond N are clipped by the tensor map
              tma_store_3d(&tmap_out, buf, col0 + ch * 32, row0, b);
              tma_store_commit();
            }
          }
        } else
        // 64 columns (= 128 B of fp16) per staged TMA store; this group takes every other 64-column chunk
        for (int ch = 2 * group; ch < chunks; ch += 4) {
          if (g.store_p) {
            if (store_leader) tma_store_wait_read<0>();       // this group's previous store has drained `buf`
            named_bar_sync(bar_id, 128);
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int chh = ch + half;
            if (chh < chunks) {
              float v[32];
              tmem_ld_32x32(taddr + chh * 32, v);
              tmem_ld_wait();
              const int nv = nvalid - chh * 32;         // valid columns in this chunk (multiple of 8 since N % 8 == 0)
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j], g.c, -m2);      // log2 of the un-normalised probability
              if (kl_row) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (j < nv) {
                    const float la_j = __ldg(la_cols + chh * 32 + j);
                    const float lnr = -fabsf(la_i - la_j);                       // ln min(a_i/a_j, a_j/a_i); -inf if a_j = 0
                    const float lnw = lnr - lnz_i;                               // ln W_ij
                    const float lnq = (v[j] - log2_l) * kLn2;                    // ln Q_ij, exact from the logits
                    const float w = ex2_approx(lnw * kLog2e);                    // 0 when a_j = 0
                    const float lr = fminf(fmaxf(lnw - lnq, -18.420680744f), 18.420680744f);   // ln clamp(W/Q, 1e-8, 1e8)
                    kl_acc += (w > 0.f) ? w * lr : 0.f;
                    klw_acc += (w > 0.f && fabsf(lnw - lnq) < 18.420680744f) ? w : 0.f;
                  }
                }
              }
              if (!(kExperiment && g.skip_epilogue == 3)) {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = ex2_approx(v[j]) * inv_l;      // inv_l carries p_scale
              }
              if (g.store_p) {
                uint32_t hh[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) hh[j] = pack_half2(v[2 * j], v[2 * j + 1]);
                if (g.lpart) {                   // training path: row sums of exactly the values the later passes will read
                  float s0 = 0.f, s1 = 0.f;      // (N % 8 == 0: a packed pair lies inside or outside the valid columns together)
#pragma unroll
                  for (int j = 0; j < 16; j += 2) {
                    s0 += (2 * j < nv) ? half2_sum_scaled(hh[j]) : 0.f;
                    s1 += (2 * j + 2 < nv) ? half2_sum_scaled(hh[j + 1]) : 0.f;
                  }
                  l_acc += (s0 + s1) * kHalfSumRescale;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  *staged_chunk(buf, row_in_tile, half * 4 + q) = make_uint4(hh[q * 4 + 0], hh[q * 4 + 1], hh[q * 4 + 2], hh[q * 4 + 3]);
              }
              if (srow && row_ok) {
                float4* dst = reinterpret_cast<float4*>(srow + chh * 32);
                const float us = g.p_scale != 1.0f ? 1.0f / g.p_scale : 1.0f;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  if (q * 4 < nv) dst[q] = make_float4(v[q * 4 + 0] * us, v[q * 4 + 1] * us, v[q * 4 + 2] * us, v[q * 4 + 3] * us);
              }
            }
          }
          if (g.store_p) {
            fence_proxy_async();                       // generic-proxy smem writes -> visible to the TMA engine
            named_bar_sync(bar_id, 128);
            if (store_leader && !(kExperiment && g.skip_epilogue == 2)) { // rows / columns beyond N are clipped by the tensor map
              tma_store_3d(&tmap_out, buf, col0 + ch * 32, row0, b);
              tma_store_commit();
            }
          }
        }
        if (kl && row_ok) g.kl_part[part_idx] = kl_acc;
        if (kl && g.klw_part && row_ok) g.klw_part[part_idx] = klw_acc;
        if ((g.unnorm || g.lpart) && row_ok) g.lpart[part_idx] = l_acc;
      } else if constexpr (MODE == MODE_DS) {
        // acc = dP_ij.  dS_ij = P_ij (dP_ij - delta_i) - gk (W_ij [in clamp window] - P_ij T_i),  gk = dLoss/dloss3 / (B N)
        const size_t brow = static_cast<size_t>(b) * g.m_rows + (row_ok ? row : 0);
        // legacy (fp32-NCHW ABI): dP already carries the scale s (dO was packed scaled), delta is unscaled.
        // ds_new (channels-last ABI): delta was computed from the operand the GEMM read; acc and delta take s_ds / s_op, the loss term s_ds.
        const float gs = g.gscale ? __ldg(g.gscale) : 1.0f;
        const float acc_mul = g.ds_new ? __ldg(g.gscale + 2) : 1.0f;
        const float delta_i = g.ds_delta ? g.ds_delta[brow] * (g.ds_new ? __ldg(g.gscale + 4) : gs) : 0.f;
        const float gks = (g.ds_gl ? __ldg(g.ds_gl) * g.ds_inv_bn : 0.f) * gs;
        const float pn = g.ds_lsum ? 1.0f / g.ds_lsum[brow] : 1.0f;       // stored probabilities -> normalised
        const bool kl = g.logbin != nullptr;
        float la_i = 0.f, lnz_i = 0.f, t_i = 0.f;
        if (kl) { la_i = g.logbin[brow]; lnz_i = g.lnz[brow]; t_i = g.ds_klT[brow]; }
        const bool kl_row = kl && la_i > -INFINITY;
        const float* la_cols = kl ? g.logbin + static_cast<size_t>(b) * g.n_cols + col0 : nullptr;
        const __half* prow = g.ds_p + (brow * g.n_cols + col0);
        for (int ch = 2 * group; ch < chunks; ch += 4) {
          if (store_leader) tma_store_wait_read<0>();
          named_bar_sync(bar_id, 128);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int chh = ch + half;
            if (chh < chunks) {
              float v[32];
              tmem_ld_32x32(taddr + chh * 32, v);
              tmem_ld_wait();
              const int nv = nvalid - chh * 32;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 pu = make_uint4(0, 0, 0, 0);
                if (row_ok && q * 8 < nv) pu = __ldg(reinterpret_cast<const uint4*>(prow + chh * 32) + q);
                const uint32_t pw[4] = {pu.x, pu.y, pu.z, pu.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float2 pp = __half22float2(*reinterpret_cast<const __half2*>(&pw[e]));
                  pp.x *= pn; pp.y *= pn;
                  const int j = q * 8 + e * 2;
                  float d0 = pp.x * (v[j] * acc_mul - delta_i), d1 = pp.y * (v[j + 1] * acc_mul - delta_i);
                  if (kl_row) {
                    const float lw0 = -fabsf(la_i - __ldg(la_cols + chh * 32 + j)) - lnz_i, lw1 = -fabsf(la_i - __ldg(la_cols + chh * 32 + j + 1)) - lnz_i;
                    const float w0 = ex2_approx(lw0 * kLog2e), w1 = ex2_approx(lw1 * kLog2e);
                    // inside the clamp window 1e-8 < W / Q < 1e8 (losses.py:12,30), tested without a logarithm: Q 1e-8 < W < Q 1e8
                    const float in0 = (w0 > pp.x * 1e-8f && w0 < pp.x * 1e8f) ? w0 : 0.f;
                    const float in1 = (w1 > pp.y * 1e-8f && w1 < pp.y * 1e8f) ? w1 : 0.f;
                    d0 -= gks * (in0 - pp.x * t_i);
                    d1 -= gks * (in1 - pp.y * t_i);
                  } else if (kl) {
                    // all-zero target row: only the softmax coupling term survives (T_i = 0 -> nothing)
                  }
                  v[j] = d0; v[j + 1] = d1;
                }
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  o[e] = pack_half2(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
                }
                *staged_chunk(buf, row_in_tile, half * 4 + q) = make_uint4(o[0], o[1], o[2], o[3]);
              }
            }
          }
          fence_proxy_async();
          named_bar_sync(bar_id, 128);
          if (store_leader) {
            tma_store_3d(&tmap_out, buf, col0 + ch * 32, row0, b);
            tma_store_commit();
          }
        }
      } else if constexpr (MODE == MODE_PVT || MODE == MODE_PVT32) {   // rows = rows of A (query / key positions), columns = channels
        float rs = g.out_scale;
        if (g.out_gscale) rs *= __ldg(g.out_gscale);
        if (g.l_in) {      // l_i = sum (fixed order) of the partial sums of the rounded probabilities the probability pass wrote
          const float* lp = g.l_in + static_cast<size_t>(b) * g.l_parts * g.m_rows + (row_ok ? row : 0);
          float l = 0.f;
          for (int t = 0; t < g.l_parts; ++t) l += __ldg(lp + static_cast<size_t>(t) * g.m_rows);
          rs *= 1.0f / l;
          if (row_ok && nt == 0 && group == 0) {
            const size_t brow = static_cast<size_t>(b) * g.m_rows + row;
            if (g.stats_w) g.stats_w[brow].y = l;
            if (g.stats_w2) g.stats_w2[brow].y = l;
            if (g.lsum_out) g.lsum_out[brow] = l;
          }
        }
        if constexpr (MODE == MODE_PVT32) {       // 32 fp32 columns (128 B) per staged store; the groups take alternate chunks
          for (int ch = group; ch < chunks; ch += 2) {
            float v[32];
            tmem_ld_32x32(taddr + ch * 32, v);
            if (store_leader) tma_store_wait_read<0>();
            tmem_ld_wait();
            named_bar_sync(bar_id, 128);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *staged_chunk(buf, row_in_tile, q) = make_uint4(__float_as_uint(v[q * 4 + 0] * rs), __float_as_uint(v[q * 4 + 1] * rs),
                                                              __float_as_uint(v[q * 4 + 2] * rs), __float_as_uint(v[q * 4 + 3] * rs));
            fence_proxy_async();
            named_bar_sync(bar_id, 128);
            if (store_leader) {
              tma_store_3d(&tmap_out, buf, col0 + ch * 32, row0, b);
              tma_store_commit();
            }
          }
        } else
        for (int ch = 2 * group; ch < chunks; ch += 4) {   // 64 16-bit columns per staged store
          float v[2][32];
          const bool two = ch + 1 < chunks;                   // warp-uniform
          tmem_ld_32x32(taddr + ch * 32, v[0]);
          if (two) tmem_ld_32x32(taddr + (ch + 1) * 32, v[1]);
          if (store_leader) tma_store_wait_read<0>();
          tmem_ld_wait();
          named_bar_sync(bar_id, 128);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (half == 0 || two) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *staged_chunk(buf, row_in_tile, half * 4 + q) =
                    make_uint4(pack16(v[half][q * 8 + 0] * rs, v[half][q * 8 + 1] * rs, g.out_bf16), pack16(v[half][q * 8 + 2] * rs, v[half][q * 8 + 3] * rs, g.out_bf16),
                               pack16(v[half][q * 8 + 4] * rs, v[half][q * 8 + 5] * rs, g.out_bf16), pack16(v[half][q * 8 + 6] * rs, v[half][q * 8 + 7] * rs, g.out_bf16));
            }
          }
          fence_proxy_async();
          named_bar_sync(bar_id, 128);
          if (store_leader) {
            tma_store_3d(&tmap_out, buf, col0 + ch * 32, row0, b);
            tma_store_commit();
          }
        }
      } else {  // MODE_PV: rows = value channels, columns = query positions; 32 fp32 columns (128 B) per staged store
        for (int ch = group; ch < chunks; ch += 2) {
          if (store_leader) tma_store_wait_read<0>();
          named_bar_sync(bar_id, 128);
          float v[32];
          tmem_ld_32x32(taddr + ch * 32, v);
          tmem_ld_wait();
          if (g.gscale) {                                 // backward outputs: undo the gradient scale
            const float us = __ldg(g.gscale + 1);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= us;
          }
          if (g.inv_l) {                                  // columns are the queries here
            const float* il = g.inv_l + static_cast<size_t>(b) * g.n_cols + col0 + ch * 32;
            const int nvq = nvalid - ch * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= (j < nvq) ? __ldg(il + j) : 0.f;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *staged_chunk(buf, row_in_tile, q) = make_uint4(__float_as_uint(v[q * 4 + 0]), __float_as_uint(v[q * 4 + 1]),
                                                            __float_as_uint(v[q * 4 + 2]), __float_as_uint(v[q * 4 + 3]));
          fence_proxy_async();
          named_bar_sync(bar_id, 128);
          if (store_leader) {
            tma_store_3d(&tmap_out, buf, col0 + ch * 32, row0, b);
            tma_store_commit();
          }
        }
      }

      if (kExperiment && g.trace && blockIdx.x == 0 && threadIdx.x == 64) g.trace[(tile / tile_step) * 4 + 3] = clock64();
      tc_fence_before();
      // one arrival per epilogue warp (8 x CG) on the LEADER's barrier releases the accumulator buffer
      __syncwarp();
      if (lane == 0) {
        if (CG == 1) mbar_arrive(&bars->tmem_empty[acc]);
        else         mbar_arrive_cluster(acc == 0 ? leader_empty0 : leader_empty1);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (store_leader) tma_store_wait_all();             // staged bytes must have left shared memory before the CTA retires
  }

  __syncwarp();                                            // single-lane role loops rejoin their warps
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // the peer may still signal our barriers until here
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------ fused inference forward (one persistent launch)
// EXPERIMENT KEPT AS AN OPT-IN (keep = 2), NOT THE DEFAULT.  Hypothesis: the probability pass is bound by its epilogue (ex2, staging), the
// aggregation pass by its operand feed; one persistent grid of CTA pairs walking a list in which the probability tiles of image b+2 are
// interleaved with the aggregation tiles of image b would hide the first under the second and pay one tail instead of two.  P~ still makes
// its fp16 round trip through L2 (the aggregation accumulator of a query tile, 128 x 2048 fp32, is four times the TMEM of an SM).
// Measured (B200, configs[1], profiles/r02_notes.md): results bit-identical to the two launches, 165 us against 148 us.  Why it loses:
// BOTH passes are bound by the same resource -- shared-memory bandwidth: per 64-deep k-step a CTA's shared memory takes 32 KB of TMA fill
// and serves 48 KB of UMMA operand reads (its A, and the whole B tile: half for its own tensor core, half for its partner's), 80 KB at
// 128 B/clock = 640 cycles against 512 cycles of MMA (the 80 % both passes and cuBLAS itself reach); the probability epilogue's staging
// traffic (128 KB per tile) lands on the same port.  Nothing complementary is left to overlap, and tiles of two sizes (8 vs 22 k-steps)
// balance worse over 74 pairs than two homogeneous launches do.
//
// Ordering: an aggregation tile (b, mt) reads the 256 rows of P~ that the probability tiles (b, mt, *) wrote -- from other SMs.  Every
// epilogue group of a probability tile, once its TMA stores have COMPLETED (cp.async.bulk.wait_group 0, not .read) and its row-sum
// partials are written, bumps done[b][mt] with a release; the TMA producer of an aggregation tile spins (acquire, bounded) until the
// count is nP x 2 groups x 2 CTAs, then crosses to the async proxy and loads.  The list is walked in increasing order by every pair and
// every dependency points backwards in it, so the wait cannot cycle (the grid is co-resident: <= 1 CTA per SM, clusters of 2).
struct FusedArgs {
  int batch, n, dv;
  int mP, nP, nV;          // tiles per image: probability mP x nP (256 x 256), aggregation mP x nV
  int kP, kV;              // 64-deep k-steps: dk / 64, ceil(N / 64)
  int stages;
  uint32_t idesc_p, idesc_v;
  float c;
  const float* dpart; int n_dpart;
  const float* diag;
  const float2* part;      // exact per-(key tile, group) maxima of images above the safe bound (fallback statistics pass)
  float2* stats_w; float2* stats_w2;
  float* lpart;            // [B, 2 nP, N]
  int* done;               // [B, mP] zeroed by the host before the launch
  int out_bf16;
};

struct FusedTile { int type, b, mt, nt; };   // type 0: probability tile, 1: aggregation tile

// Tile t of the static list.  The list is a sequence of segments; segment s holds the probability tiles of image s (s < B) and the
// aggregation tiles of image s - kFusedLag (s >= kFusedLag), spread evenly through each other.  The lag keeps every dependency (an
// aggregation tile needs the probability tiles of its row block COMPLETED, epilogue and stores included) more than two waves of the
// 74 resident pairs behind its consumer: with a lag of one segment the producers found their row block still in flight and the
// pipeline drained at every tile (measured: 189 us against 146 us for two launches).
constexpr int kFusedLag = 2;
__device__ __forceinline__ FusedTile fused_decode(const FusedArgs& g, int t) {
  const int TP = g.mP * g.nP, TV = g.mP * g.nV;
  FusedTile r;
  r.type = 1; r.b = g.batch - 1;
  int idx = 0, t0 = 0;
  for (int s = 0; s < g.batch + kFusedLag; ++s) {
    const int np = s < g.batch ? TP : 0, nv = s >= kFusedLag ? TV : 0, size = np + nv;
    if (t < t0 + size) {
      const int j = t - t0;
      if (np == 0) { r.type = 1; r.b = s - kFusedLag; idx = j; }
      else if (nv == 0) { r.type = 0; r.b = s; idx = j; }
      else {
        const int cp1 = ((j + 1) * TP) / size, cp0 = (j * TP) / size;      // probability tiles spread evenly through the segment
        if (cp1 > cp0) { r.type = 0; r.b = s; idx = cp0; }
        else { r.type = 1; r.b = s - kFusedLag; idx = j - cp0; }
      }
      break;
    }
    t0 += size;
  }
  const int nn = r.type == 0 ? g.nP : g.nV;
  r.mt = idx / nn;
  r.nt = idx - r.mt * nn;
  return r;
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(kGemmThreads, 1)
fused_fwd_kernel(const __grid_constant__ CUtensorMap tmap_f, const __grid_constant__ CUtensorMap tmap_p, const __grid_constant__ CUtensorMap tmap_v,
                 const __grid_constant__ CUtensorMap tmap_o, const FusedArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  constexpr uint32_t a_bytes = BM * BK * 2, b_bytes = 128 * BK * 2, stage_bytes = a_bytes + b_bytes;   // BN = 256: each CTA stages half of B
  constexpr int BN = 256;
  uint8_t* out_stage = smem + static_cast<size_t>(g.stages) * stage_bytes;
  PipeBarriers* bars = reinterpret_cast<PipeBarriers*>(out_stage + kOutBuffers * kOutChunkBytes);
  float* const dmax_s = bars->dmax;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  griddep_launch_dependents();
  const uint32_t cta_rank = cluster_ctarank();
  const int total_tiles = g.batch * g.mP * (g.nP + g.nV);
  const int first_tile = blockIdx.x / 2, tile_step = gridDim.x / 2;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap_f); tma_prefetch_desc(&tmap_p); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_o); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < g.stages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
      for (int s = 0; s < 2; ++s) { mbar_init(&bars->tmem_full[s], 1); mbar_init(&bars->tmem_empty[s], 16); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_2sm(&bars->tmem_base, kTmemCols);
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const FusedTile T = fused_decode(g, tile);
        const int a_row = (T.mt * 2 + static_cast<int>(cta_rank)) * BM;
        const int b_row = T.nt * BN + static_cast<int>(cta_rank) * 128;
        if (T.type == 1) {     // the probabilities of this row block must be in L2, written by other SMs' TMA stores
          const int* flag = g.done + T.b * g.mP + T.mt;
          const int want = g.nP * 4;
          if (ld_acquire_gpu(flag) < want) {
            const long long t0 = clock64();
            while (ld_acquire_gpu(flag) < want) {
              __nanosleep(64);
              if (clock64() - t0 > 4000000000LL) { printf("acan: fused forward: dependency timeout (image %d row block %d)\n", T.b, T.mt); __trap(); }
            }
          }
          asm volatile("fence.proxy.async.global;" ::: "memory");      // generic-proxy acquire -> the async-proxy (TMA) loads that follow
        }
        const int ksteps = T.type == 0 ? g.kP : g.kV;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&bars->empty[stage], phase ^ 1);
          uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
          if (cta_rank == 0) mbar_expect_tx(&bars->full[stage], stage_bytes * 2);
          const uint32_t leader_bar = mapa_u32(smem_u32(&bars->full[stage]), 0);
          if (T.type == 0) {
            tma_load_3d_2sm(sa, &tmap_f, leader_bar, ks * BK, a_row, T.b);
            tma_load_3d_2sm(sa + a_bytes, &tmap_f, leader_bar, ks * BK, b_row, T.b);
          } else {
            tma_load_3d_2sm(sa, &tmap_p, leader_bar, ks * BK, a_row, T.b);
            tma_load_3d_2sm(sa + a_bytes, &tmap_v, leader_bar, b_row, ks * BK, T.b);
            tma_load_3d_2sm(sa + a_bytes + 8192, &tmap_v, leader_bar, b_row + 64, ks * BK, T.b);
          }
          if (++stage == static_cast<uint32_t>(g.stages)) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const FusedTile T = fused_decode(g, tile);
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kMaxBN;
        const int ksteps = T.type == 0 ? g.kP : g.kV;
        const uint32_t idesc = T.type == 0 ? g.idesc_p : g.idesc_v;
        for (int ks = 0; ks < ksteps; ++ks) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
          const uint32_t sb = sa + a_bytes;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = umma_desc_k_sw128(sa + k * UMMA_K * 2);
            const uint64_t db = T.type == 0 ? umma_desc_k_sw128(sb + k * UMMA_K * 2) : umma_desc_mn_sw128(sb + k * UMMA_K * 128, 8192);
            umma_f16_2sm(d_tmem, da, db, idesc, (ks | k) != 0 ? 1u : 0u);
          }
          umma_commit_2sm(&bars->empty[stage], 3);
          if (++stage == static_cast<uint32_t>(g.stages)) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm(&bars->tmem_full[acc], 3);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int quarter = warp & 3, group = (warp - 2) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const bool store_leader = ((threadIdx.x - 64) & 127) == 0;
    uint8_t* const buf = out_stage + group * kOutChunkBytes;
    const int bar_id = 1 + group;
    uint32_t acc = 0, acc_phase = 0;
    {
      const int t = threadIdx.x - 64;
      if (t < g.batch && t < kDmaxSlots) dmax_s[t] = image_dmax(g.dpart, g.n_dpart, t);
      named_bar_sync(3, 256);
    }
    const uint32_t leader_empty0 = mapa_u32(smem_u32(&bars->tmem_empty[0]), 0), leader_empty1 = mapa_u32(smem_u32(&bars->tmem_empty[1]), 0);
    for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
      const FusedTile T = fused_decode(g, tile);
      const int row0 = (T.mt * 2 + static_cast<int>(cta_rank)) * BM;
      const int row = row0 + row_in_tile;
      const bool row_ok = row < g.n;
      const int col0 = T.nt * BN;
      const int ncols = T.type == 0 ? g.n : g.dv;
      const int nvalid = min(BN, ncols - col0);
      const int chunks = (nvalid + 31) >> 5;
      const size_t brow = static_cast<size_t>(T.b) * g.n + (row_ok ? row : 0);
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kMaxBN;
      if (T.type == 0) {
        float m2;
        const float dmax = T.b < kDmaxSlots ? dmax_s[T.b] : image_dmax(g.dpart, g.n_dpart, T.b);
        if (dmax <= kSafeDiagMax) {
          m2 = sqrtf(g.diag[brow] * dmax) - kShiftHeadroom;
        } else {
          m2 = -INFINITY;
          const float2* pp = g.part + static_cast<size_t>(T.b) * (2 * g.nP) * g.n + (row_ok ? row : 0);
          for (int t = 0; t < 2 * g.nP; ++t) m2 = fmaxf(m2, pp[static_cast<size_t>(t) * g.n].x);
        }
        if (row_ok && T.nt == 0 && group == 0) {
          g.stats_w[brow].x = m2;
          if (g.stats_w2) g.stats_w2[brow].x = m2;
        }
        float l_acc = 0.f;
        for (int ch = 2 * group; ch < chunks; ch += 4) {
          float v0[32], v1[32];
          const bool two = ch + 1 < chunks;
          tmem_ld_32x32(taddr + ch * 32, v0);
          if (two) tmem_ld_32x32(taddr + (ch + 1) * 32, v1);
          if (store_leader) tma_store_wait_read<0>();
          tmem_ld_wait();
          named_bar_sync(bar_id, 128);
          l_acc += probs_unnorm_half(v0, nvalid - ch * 32, g.c, m2, buf, row_in_tile, 0);
          if (two) l_acc += probs_unnorm_half(v1, nvalid - (ch + 1) * 32, g.c, m2, buf, row_in_tile, 1);
          fence_proxy_async();
          named_bar_sync(bar_id, 128);
          if (store_leader) {
            tma_store_3d(&tmap_p, buf, col0 + ch * 32, row0, T.b);
            tma_store_commit();
          }
        }
        if (row_ok) g.lpart[(static_cast<size_t>(T.b) * (2 * g.nP) + 2 * T.nt + group) * g.n + row] = l_acc;
        // accumulator can go back to the MMA warp before the stores have landed
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc == 0 ? leader_empty0 : leader_empty1);
        // publish: stores complete -> (group barrier: every thread's row-sum partial is written) -> release
        if (store_leader) tma_store_wait_all();
        named_bar_sync(bar_id, 128);
        if (store_leader) {
          asm volatile("fence.proxy.async.global;" ::: "memory");
          red_release_gpu_add(g.done + T.b * g.mP + T.mt, 1);
        }
      } else {
        const int l_parts = 2 * g.nP;
        const float* lp = g.lpart + static_cast<size_t>(T.b) * l_parts * g.n + (row_ok ? row : 0);
        float l = 0.f;
        for (int t = 0; t < l_parts; ++t) l += __ldcg(lp + static_cast<size_t>(t) * g.n);     // written by other SMs in this launch: L2, not L1
        const float rs = 1.0f / l;
        if (row_ok && T.nt == 0 && group == 0) {
          g.stats_w[brow].y = l;
          if (g.stats_w2) g.stats_w2[brow].y = l;
        }
        for (int ch = 2 * group; ch < chunks; ch += 4) {
          float v[2][32];
          const bool two = ch + 1 < chunks;
          tmem_ld_32x32(taddr + ch * 32, v[0]);
          if (two) tmem_ld_32x32(taddr + (ch + 1) * 32, v[1]);
          if (store_leader) tma_store_wait_read<0>();
          tmem_ld_wait();
          named_bar_sync(bar_id, 128);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            if (half == 0 || two) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *staged_chunk(buf, row_in_tile, half * 4 + q) =
                    make_uint4(pack16(v[half][q * 8 + 0] * rs, v[half][q * 8 + 1] * rs, g.out_bf16), pack16(v[half][q * 8 + 2] * rs, v[half][q * 8 + 3] * rs, g.out_bf16),
                               pack16(v[half][q * 8 + 4] * rs, v[half][q * 8 + 5] * rs, g.out_bf16), pack16(v[half][q * 8 + 6] * rs, v[half][q * 8 + 7] * rs, g.out_bf16));
            }
          }
          fence_proxy_async();
          named_bar_sync(bar_id, 128);
          if (store_leader) {
            tma_store_3d(&tmap_o, buf, col0 + ch * 32, row0, T.b);
            tma_store_commit();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc == 0 ? leader_empty0 : leader_empty1);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (store_leader) tma_store_wait_all();
  }
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------ small kernels

// F fp32 [B, dk, N]  ->  Fh fp16 [B, N, dk]   (position-major rows: the K-major operand of S = F^T F)
__global__ void __launch_bounds__(256) pack_feat_kernel(const float* __restrict__ f, __half* __restrict__ fh, int dk, int N) {
  __shared__ float tile[64][65];
  const int b = blockIdx.z, c0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const float* src = f + static_cast<size_t>(b) * dk * N;
  for (int i = threadIdx.x; i < 64 * 64; i += 256) {
    const int c = i >> 6, n = i & 63;
    tile[c][n] = (n0 + n < N) ? src[static_cast<size_t>(c0 + c) * N + n0 + n] : 0.f;
  }
  __syncthreads();
  __half* dst = fh + static_cast<size_t>(b) * N * dk;
  for (int i = threadIdx.x; i < 64 * 32; i += 256) {
    const int n = i >> 5, c2 = (i & 31) * 2;
    if (n0 + n < N) {
      // clamp to the fp16 range: post-BN/ReLU activations never get there, +-inf would poison the softmax
      const float a = fminf(fmaxf(tile[c2][n], -65504.f), 65504.f), bq = fminf(fmaxf(tile[c2 + 1][n], -65504.f), 65504.f);
      *reinterpret_cast<__half2*>(dst + static_cast<size_t>(n0 + n) * dk + c0 + c2) = __floats2half2_rn(a, bq);
    }
  }
}

// V fp32 [B, dv, N] -> Vh fp16, same layout (keys contiguous: the K-major A operand of O^T = V P^T)
__global__ void __launch_bounds__(256) pack_value_kernel(const float4* __restrict__ v, uint4* __restrict__ vh, int64_t n8) {
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n8; i += static_cast<int64_t>(gridDim.x) * 256) {
    const float4 a = ldg_stream(v + 2 * i), b = ldg_stream(v + 2 * i + 1);
    auto cl = [](float x) { return fminf(fmaxf(x, -65504.f), 65504.f); };
    vh[i] = make_uint4(pack_half2(cl(a.x), cl(a.y)), pack_half2(cl(a.z), cl(a.w)), pack_half2(cl(b.x), cl(b.y)), pack_half2(cl(b.z), cl(b.w)));
  }
}

// (m2, l) partials over key tiles -> final row statistics
__global__ void __launch_bounds__(256) combine_stats_kernel(const float2* __restrict__ part, float2* __restrict__ stats, float2* __restrict__ stats2,
                                                             int n_tiles, int N, int B) {
  const int64_t total = static_cast<int64_t>(B) * N;
  for (int64_t q = blockIdx.x * 256LL + threadIdx.x; q < total; q += static_cast<int64_t>(gridDim.x) * 256) {
    const int b = static_cast<int>(q / N), i = static_cast<int>(q - static_cast<int64_t>(b) * N);
    const float2* p = part + static_cast<size_t>(b) * n_tiles * N + i;
    float m = -INFINITY;
    for (int t = 0; t < n_tiles; ++t) m = fmaxf(m, p[static_cast<size_t>(t) * N].x);
    float l = 0.f;
    for (int t = 0; t < n_tiles; ++t) { const float2 v = p[static_cast<size_t>(t) * N]; l += v.y * ex2_approx(v.x - m); }
    stats[q] = make_float2(m, l);
    if (stats2) stats2[q] = make_float2(m, l);
  }
}

// Cauchy-Schwarz row shift (inference fast path, no statistics pass): with d_i = S_ii * c (log2 units) and Q = K = F,
//   d_i <= max_j S_ij * c <= sqrt(d_i * d_max),
// so shift_i = sqrt(d_i * d_max) - 14 keeps p~ = exp2(S_ij*c - shift_i) <= 2^14 (no fp16 overflow) and every entry within
// 2^-10 of the row maximum a normal fp16 as long as sqrt(d_i d_max) - d_i <= d_max / 4 <= 18, i.e. d_max <= 72 (S up to
// ~50 in natural units; random-init / BN'd features give ~16).  Images above that limit take the exact max instead.
//
// grid (ceil(N/128), B), 1024 threads: each warp computes d_i for 4 rows; the CTA leaves the maximum of its 128 rows in
// dpart[b, blockIdx.x].  The consumers (probability-pass epilogue, fallback statistics pass) reduce those <= 13 values per
// image themselves: no atomics, no counters to reset, no second kernel.
constexpr int kDiagRows = 128, kDiagThreads = 512;
__global__ void __launch_bounds__(kDiagThreads, 2) diag_kernel(const __half* __restrict__ fh, float* __restrict__ diag, float* __restrict__ dpart,
                                                             int N, int dk, float c) {
  __shared__ float red[kDiagThreads / 32];
  griddep_launch_dependents();
  const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5, b = blockIdx.y;
  const __half* base = fh + static_cast<size_t>(b) * N * dk;
  float* d = diag + static_cast<size_t>(b) * N;
  const int k8 = dk / 8;                                   // 16-byte chunks per row; dk % 64 == 0 so k8 % 8 == 0
  float wm = 0.f;
#pragma unroll 1
  for (int r = 0; r < kDiagRows / (kDiagThreads / 32); r += 4) {          // 8 rows per warp, 4 at a time
    const int i0 = blockIdx.x * kDiagRows + wv * (kDiagRows / (kDiagThreads / 32)) + r;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < k8; k += 64) {                  // 8 independent, unconditional 16 B loads in flight per lane
      const bool second = k + 32 < k8;
      uint4 u[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4* rowp = reinterpret_cast<const uint4*>(base + static_cast<size_t>(min(i0 + q, N - 1)) * dk);
        u[2 * q] = ldg_nc_u4(rowp + k);
        u[2 * q + 1] = ldg_nc_u4(rowp + (second ? k + 32 : k));
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t w[4] = {u[q].x, u[q].y, u[q].z, u[q].w};
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
          a0 = fmaf(t.x, t.x, a0);
          a1 = fmaf(t.y, t.y, a1);
        }
        if ((q & 1) == 0 || second) s[q >> 1] += a0 + a1;
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float t = warp_sum(s[q]) * c;
      if (i0 + q < N) { wm = fmaxf(wm, t); if (lane == 0) d[i0 + q] = t; }
    }
  }
  if (lane == 0) red[wv] = wm;
  __syncthreads();
  if (wv == 0) {
    const float m = warp_max(lane < kDiagThreads / 32 ? red[lane] : 0.f);
    if (lane == 0) dpart[static_cast<size_t>(b) * gridDim.x + blockIdx.x] = m;
  }
}

// l_i = sum of the partial sums of the rounded p~; stats <- (shift_i, l_i), inv_l <- 1 / l_i   (fp32-NCHW aggregation pass only:
// the channels-last pass, whose accumulator rows are the queries, does this in its epilogue)
__global__ void __launch_bounds__(256) combine_l_kernel(const float* __restrict__ lpart, float2* __restrict__ stats, float2* __restrict__ stats2,
                                                         float* __restrict__ inv_l, int n_parts, int N, int64_t total) {
  for (int64_t q = blockIdx.x * 256LL + threadIdx.x; q < total; q += static_cast<int64_t>(gridDim.x) * 256) {
    const int b = static_cast<int>(q / N), i = static_cast<int>(q - static_cast<int64_t>(b) * N);
    const float* p = lpart + static_cast<size_t>(b) * n_parts * N + i;
    float l = 0.f;
    for (int t = 0; t < n_parts; ++t) l += p[static_cast<size_t>(t) * N];
    stats[q].y = l;
    if (stats2) stats2[q].y = l;
    inv_l[q] = 1.0f / l;
  }
}

// bf16 / fp32 -> fp16, same layout, times *gscale (null = 1), clamped to the fp16 range (exact for bf16 values whose product stays a
// normal fp16: a power-of-two scale only moves the exponent)
template <bool SRC_BF16>
__global__ void __launch_bounds__(256) to_half_kernel(const void* __restrict__ src, uint4* __restrict__ out, int64_t n8, const float* __restrict__ gscale = nullptr) {
  const float sc = gscale ? __ldg(gscale) : 1.0f;
  auto cl = [sc](float x) { return fminf(fmaxf(x * sc, -65504.f), 65504.f); };
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n8; i += static_cast<int64_t>(gridDim.x) * 256) {
    float f[8];
    if (SRC_BF16) {
      const uint4 u = ldg_nc_u4(reinterpret_cast<const uint4*>(src) + i);
      const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { f[2 * e] = __uint_as_float(wv[e] << 16); f[2 * e + 1] = __uint_as_float(wv[e] & 0xffff0000u); }
    } else {
      const float4 a = ldg_stream(reinterpret_cast<const float4*>(src) + 2 * i), b = ldg_stream(reinterpret_cast<const float4*>(src) + 2 * i + 1);
      f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    out[i] = make_uint4(pack_half2(cl(f[0]), cl(f[1])), pack_half2(cl(f[2]), cl(f[3])), pack_half2(cl(f[4]), cl(f[5])), pack_half2(cl(f[6]), cl(f[7])));
  }
}

// ---- backward helpers
// fp32 -> fp16 / bf16, same layout (clamped to the fp16 range for fp16)
template <bool BF16>
__global__ void __launch_bounds__(256) pack_cast_kernel(const float4* __restrict__ v, uint4* __restrict__ out, int64_t n8, const float* __restrict__ gscale) {
  const float sc = gscale ? __ldg(gscale) : 1.0f;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n8; i += static_cast<int64_t>(gridDim.x) * 256) {
    float4 a = ldg_stream(v + 2 * i), b = ldg_stream(v + 2 * i + 1);
    a.x *= sc; a.y *= sc; a.z *= sc; a.w *= sc; b.x *= sc; b.y *= sc; b.z *= sc; b.w *= sc;
    if (BF16) {
      auto pk = [](float x, float y) { const __nv_bfloat162 h = __floats2bfloat162_rn(x, y); return *reinterpret_cast<const uint32_t*>(&h); };
      out[i] = make_uint4(pk(a.x, a.y), pk(a.z, a.w), pk(b.x, b.y), pk(b.z, b.w));
    } else {
      auto cl = [](float x) { return fminf(fmaxf(x, -65504.f), 65504.f); };
      out[i] = make_uint4(pack_half2(cl(a.x), cl(a.y)), pack_half2(cl(a.z), cl(a.w)), pack_half2(cl(b.x), cl(b.y)), pack_half2(cl(b.z), cl(b.w)));
    }
  }
}

// delta[b,i] = sum_c dO[b,c,i] * O[b,c,i]   (= sum_j P_ij dP_ij, the softmax-backward row term)
// grid (ceil(N/64), B): a lane owns 2 adjacent positions (8 B loads), the 8 warps stride over the channels, 8 channel rows of both
// tensors in flight per thread (16 loads; the register target in __launch_bounds__ lets ptxas keep them batched)
__global__ void __launch_bounds__(256, 2) delta_kernel(const float* __restrict__ go, const float* __restrict__ o, float* __restrict__ delta,
                                                        int* __restrict__ amax_bits, int C, int N) {
  __shared__ float2 red[8][33];
  const int b = blockIdx.y, lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
  const int i = blockIdx.x * 64 + lane * 2;
  float2 s = make_float2(0.f, 0.f);
  float am = 0.f;
  if (i < N) {                                           // N % 8 == 0: a pair of positions is inside or outside together
    const size_t base = static_cast<size_t>(b) * C * N + i;
    int c = wv;
    for (; c + 56 < C; c += 64) {
      float2 gv[8], ov[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        gv[u] = __ldg(reinterpret_cast<const float2*>(go + base + static_cast<size_t>(c + 8 * u) * N));
        ov[u] = __ldg(reinterpret_cast<const float2*>(o + base + static_cast<size_t>(c + 8 * u) * N));
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        s.x = fmaf(gv[u].x, ov[u].x, s.x);
        s.y = fmaf(gv[u].y, ov[u].y, s.y);
        am = fmaxf(am, fmaxf(fabsf(gv[u].x), fabsf(gv[u].y)));
      }
    }
    for (; c < C; c += 8) {
      const float2 gv = __ldg(reinterpret_cast<const float2*>(go + base + static_cast<size_t>(c) * N));
      const float2 ov = __ldg(reinterpret_cast<const float2*>(o + base + static_cast<size_t>(c) * N));
      s.x = fmaf(gv.x, ov.x, s.x);
      s.y = fmaf(gv.y, ov.y, s.y);
      am = fmaxf(am, fmaxf(fabsf(gv.x), fabsf(gv.y)));
    }
  }
  am = warp_max(am);
  if (lane == 0 && am > 0.f) atomicMax(amax_bits, __float_as_int(am));   // |dO| max over the tensor (one atomic per warp)
  red[wv][lane] = s;
  __syncthreads();
  if (wv == 0 && i < N) {
    float2 t = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) { t.x += red[k][lane].x; t.y += red[k][lane].y; }
    *reinterpret_cast<float2*>(delta + static_cast<size_t>(b) * N + i) = t;
  }
}

// out[q] = sum_t parts[b, t, i]
__global__ void __launch_bounds__(256) sum_parts_kernel(const float* __restrict__ parts, float* __restrict__ out, int n_parts, int N, int64_t total) {
  for (int64_t q = blockIdx.x * 256LL + threadIdx.x; q < total; q += static_cast<int64_t>(gridDim.x) * 256) {
    const int b = static_cast<int>(q / N), i = static_cast<int>(q - static_cast<int64_t>(b) * N);
    const float* p = parts + static_cast<size_t>(b) * n_parts * N + i;
    float s = 0.f;
    for (int t = 0; t < n_parts; ++t) s += p[static_cast<size_t>(t) * N];
    out[q] = s;
  }
}

// (s, 1/s): power of two that brings amax |dO| to 2^-2 -- or, without an aggregation gradient, the attention-loss gradient scale to ~1;
// chosen on the device: the loss gradient arrives as a device scalar, the step never synchronises with the host
__global__ void gscale_kernel(const int* __restrict__ amax_bits, const float* __restrict__ grad_loss, float inv_bn, float* __restrict__ sc) {
  float s = 1.0f;
  if (amax_bits != nullptr) {
    const float am = __int_as_float(*amax_bits);
    s = (am > 0.f && am < INFINITY) ? exp2f(-2.0f - ceilf(log2f(am))) : 1.0f;
  } else if (grad_loss != nullptr) {
    // loss-only backward: |dS| <= gk = |dLoss/dloss3| / (B N); bring it to ~1
    const float gk = fabsf(*grad_loss * inv_bn);
    s = (gk > 0.f && gk < INFINITY) ? exp2f(-ceilf(log2f(gk))) : 1.0f;
  }
  sc[0] = s;
  sc[1] = 1.0f / s;
}

__global__ void __launch_bounds__(256) scaled_sum_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out, float scale_in,
                                                          const float* __restrict__ gscale, int64_t n4) {
  const float scale = scale_in * __ldg(gscale + 1);
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * 256) {
    const float4 x = a[i], y = b[i];
    out[i] = make_float4(scale * (x.x + y.x), scale * (x.y + y.y), scale * (x.z + y.z), scale * (x.w + y.w));
  }
}

// attention-loss-only dS (no aggregation gradient): dS_ij = -gk (W_ij [in window] - P_ij T_i), one CTA per row, scaled fp16 out
__global__ void __launch_bounds__(256) kl_ds_kernel(const __half* __restrict__ p, const float* __restrict__ logbin, const float* __restrict__ lnz,
                                                     const float* __restrict__ klT, __half* __restrict__ ds, const float* __restrict__ grad_loss,
                                                     float inv_bn, const float* __restrict__ gscale, int N, const float* __restrict__ lsum = nullptr) {
  const float gk = __ldg(grad_loss) * inv_bn * __ldg(gscale);
  const int64_t row = blockIdx.x;
  const float pn = lsum ? 1.0f / lsum[row] : 1.0f;
  const int b = static_cast<int>(row / N);
  const float la_i = logbin[row], lnz_i = lnz[row], t_i = klT[row];
  const float* la = logbin + static_cast<size_t>(b) * N;
  const __half* pr = p + row * N;
  __half* dr = ds + row * N;
  for (int j = threadIdx.x; j < N; j += 256) {
    const float pv = __half2float(pr[j]) * pn;
    float d = 0.f;
    if (la_i > -INFINITY) {
      const float lw = -fabsf(la_i - la[j]) - lnz_i;
      const float w = ex2_approx(lw * kLog2e);
      const float in = (w > 0.f && fabsf(lw - lg2_approx(pv) * kLn2) < 18.420680744f) ? w : 0.f;
      d = -gk * (in - pv * t_i);
    }
    dr[j] = __float2half_rn(d);
  }
}

// channels-last backward: delta[row] = sum_c dO[row, c] * O[row, c] with dO the 16-bit operand the dP GEMM reads (so that
// delta_i == sum_j p_ij dP_ij holds for the values the tensor cores see) and O the fp32 aggregation output; max |dO| on the side.
// One warp per row, a lane owns 8 consecutive channels per step, 4 steps in flight.
template <bool GO_BF16>
__global__ void __launch_bounds__(256, 4) delta_nhwc_kernel(const uint4* __restrict__ go, const float4* __restrict__ o, float* __restrict__ delta,
                                                             int* __restrict__ amax_bits, int64_t rows, int C) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint4* g = go + row * (C / 8);
  const float4* x = o + row * (C / 4);
  float acc = 0.f, am = 0.f;
  auto term = [&](const uint4& u, const float4& a, const float4& b) {
    const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
    const float of[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float g0, g1;
      if (GO_BF16) { g0 = __uint_as_float(wv[e] << 16); g1 = __uint_as_float(wv[e] & 0xffff0000u); }
      else { const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&wv[e])); g0 = t.x; g1 = t.y; }
      acc = fmaf(g0, of[2 * e], acc);
      acc = fmaf(g1, of[2 * e + 1], acc);
      am = fmaxf(am, fmaxf(fabsf(g0), fabsf(g1)));
    }
  };
  const int c8 = C / 8;
  int k = lane;
  for (; k + 96 < c8; k += 128) {
    uint4 u[4]; float4 a[4], b[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { u[q] = ldg_nc_u4(g + k + 32 * q); a[q] = __ldg(x + 2 * (k + 32 * q)); b[q] = __ldg(x + 2 * (k + 32 * q) + 1); }
#pragma unroll
    for (int q = 0; q < 4; ++q) term(u[q], a[q], b[q]);
  }
  for (; k < c8; k += 32) term(ldg_nc_u4(g + k), __ldg(x + 2 * k), __ldg(x + 2 * k + 1));
  acc = warp_sum(acc);
  am = warp_max(am);
  if (lane == 0) {
    delta[row] = acc;
    if (amax_bits && am > 0.f) atomicMax(amax_bits, __float_as_int(am));
  }
}

// max |x| over an fp32 tensor (fp32 grad_ctx: the operand pack needs its scale first)
__global__ void __launch_bounds__(256) amax_f32_kernel(const float4* __restrict__ x, int64_t n4, int* __restrict__ amax_bits) {
  float am = 0.f;
  for (int64_t i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += static_cast<int64_t>(gridDim.x) * 256) {
    const float4 v = ldg_stream(x + i);
    am = fmaxf(am, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  am = warp_max(am);
  if ((threadIdx.x & 31) == 0 && am > 0.f) atomicMax(amax_bits, __float_as_int(am));
}

// channels-last backward scales, chosen on the device: sc = (s_ds, 1/s_ds, s_ds/s_op, 1/s_op, delta_mul).
//   s_ds brings max |dO| to 2^-2 (dS = p (dP - delta) is stored as fp16(dS * s_ds)); without an aggregation gradient it brings the
//   attention-loss gradient scale to ~1.   s_op = s_ds when dO was packed to fp16 (scaled while packing), 1 when it is read in place.
//   delta_mul takes delta to the scale of dS: s_ds when delta was computed from the unscaled gradient, s_ds / s_op from the packed one.
__global__ void gscale4_kernel(const int* __restrict__ amax_bits, int packed, int delta_from_packed, const float* __restrict__ grad_loss, float inv_bn,
                               float* __restrict__ sc) {
  float s = 1.0f;
  if (amax_bits != nullptr) {
    const float am = __int_as_float(*amax_bits);
    s = (am > 0.f && am < INFINITY) ? exp2f(-2.0f - ceilf(log2f(am))) : 1.0f;
  } else if (grad_loss != nullptr) {
    const float gk = fabsf(*grad_loss * inv_bn);
    s = (gk > 0.f && gk < INFINITY) ? exp2f(-ceilf(log2f(gk))) : 1.0f;
  }
  const float s_op = packed ? s : 1.0f;
  sc[0] = s;
  sc[1] = 1.0f / s;
  sc[2] = s / s_op;
  sc[3] = 1.0f / s_op;
  sc[4] = delta_from_packed ? s / s_op : s;
}

// per row: ln(bin) and ln(Z_i), Z_i = sum_j min(a_i/a_j, a_j/a_i)   (target-affinity normaliser, losses.py:40-46)
__global__ void __launch_bounds__(256) gt_norm_kernel(const float* __restrict__ bins, float* __restrict__ logbin, float* __restrict__ lnz, int N) {
  // grid (ceil(N/8), B): the image's log-bins are staged in shared memory; one WARP per row sums exp(-|la_i - la_j|) over j in lane
  // strides with a fixed shuffle tree (one thread per row left the pass latency-bound: 11 k threads, N sequential ex2 each)
  extern __shared__ float la[];
  const int b = blockIdx.y;
  const float* a = bins + static_cast<size_t>(b) * N;
  for (int j = threadIdx.x; j < N; j += 256) { const float v = a[j]; la[j] = v > 0.f ? logf(v) * kLog2e : -INFINITY; }   // log2 units
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= N) return;
  const float li = la[i];
  float z = 0.f;
  if (li > -INFINITY) {
    for (int j = lane; j < N; j += 32) z += ex2_approx(-fabsf(li - la[j]));      // exp2(-inf) = 0 for bin 0
  }
  z = warp_sum(z);
  if (lane == 0) {
    logbin[static_cast<size_t>(b) * N + i] = li > -INFINITY ? li * kLn2 : -INFINITY;
    lnz[static_cast<size_t>(b) * N + i] = z > 0.f ? lg2_approx(z) * kLn2 : 0.f;
  }
}

// selected rows of sim_map on CUDA cores (3 rows per image for the TensorBoard figure, vis_utils.py:107-121)
__global__ void __launch_bounds__(256) sim_rows_kernel(const __half* __restrict__ fh, const float2* __restrict__ stats,
                                                        const int* __restrict__ rows, float* __restrict__ out,
                                                        int n_rows, int N, int dk, float c) {
  extern __shared__ float qrow[];
  const int b = blockIdx.y, r = blockIdx.x;
  const int i = min(max(rows[r], 0), N - 1);              // callers validate the indices; never read out of bounds anyway
  const __half* base = fh + static_cast<size_t>(b) * N * dk;
  for (int k = threadIdx.x; k < dk; k += 256) qrow[k] = __half2float(base[static_cast<size_t>(i) * dk + k]);
  __syncthreads();
  const float2 st = stats[static_cast<size_t>(b) * N + i];
  const int lane = threadIdx.x & 31;
  for (int j = threadIdx.x >> 5; j < N; j += 8) {
    const __half* kj = base + static_cast<size_t>(j) * dk;
    float acc = 0.f;
    for (int k = lane * 2; k < dk; k += 64) {
      const float2 kv = __half22float2(*reinterpret_cast<const __half2*>(kj + k));
      acc += kv.x * qrow[k] + kv.y * qrow[k + 1];
    }
    acc = warp_sum(acc);
    if (lane == 0) out[(static_cast<size_t>(b) * n_rows + r) * N + j] = ex2_approx(fmaf(acc, c, -st.x)) / st.y;
  }
}

// ------------------------------------------------------------------------------------ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// tensor [batch][rows][inner] (inner contiguous) of fp16 (elem_bytes 2) or fp32 (4); box = {128 B of elements, box_rows, 1},
// SWIZZLE_128B, out-of-bounds elements read as 0 / are not written
int make_tmap(CUtensorMap* m, const void* base, int elem_bytes, int inner, int rows, int batch, size_t row_pitch_bytes, size_t batch_pitch_bytes,
              int box_rows, bool bf16 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(ACAN_ERR_DRIVER, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(batch)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(row_pitch_bytes), static_cast<cuuint64_t>(batch_pitch_bytes)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(128 / elem_bytes), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, elem_bytes == 2 ? (bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16) : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ACAN_ERR_DRIVER, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%d rows=%d batch=%d box_rows=%d)", (int)r, inner, rows, batch, box_rows);
  return 0;
}

// Experiment builds (-DACAN_EXPERIMENT) read overrides from the environment (ACAN_BN_QK / ACAN_BN_PV tile widths, ACAN_SKIP_EPI,
// ACAN_TRACE, ACAN_EXACT_STATS, ACAN_PDL=0, ACAN_CTA_GROUP=1); the shipped library returns the default without looking.
int env_int(const char* name, int dflt) {
  if (!kExperiment) return dflt;
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

bool use_pdl() {            // programmatic dependent launch between the passes of one call
  static const bool v = env_int("ACAN_PDL", 1) != 0;
  return v;
}

int cta_group() {           // CTA pairs (cta_group::2) unless an experiment build asks for single-CTA tiles
  static const int cg = env_int("ACAN_CTA_GROUP", 2) == 1 ? 1 : 2;
  return cg;
}

// Tile width.  Per 64-deep k-step a CTA spends max(MMA time, operand-feed time) cycles:
//   MMA   : 128*cg x bn x 64 MACs at 4096 MAC/cycle/SM                      = 2 * bn
//   feed  : (16 KB of A + bn/cg rows x 128 B of B) at ~51 B/cycle/SM, the L2 -> SM rate both the bn = 160 (59 % of
//           peak) and bn = 256 (80 %) measurements of the aggregation pass fit (profiles/r01_notes.md)
// and the pass costs (waves over the 148 / cg tile slots) x that.  Wide tiles win unless they quantise badly.
int choose_bn(int n_cols, int other_tiles, int cg, int step, int start = kMinBN) {
  int best = start;
  double best_cost = 1e30;
  const int slots = kNumSMs / cg;
  for (int bn = start; bn <= kMaxBN; bn += step) {
    const int nt = (n_cols + bn - 1) / bn;
    const long tiles = static_cast<long>(nt) * other_tiles;
    const long waves = (tiles + slots - 1) / slots;
    const double mma = 2.0 * bn, feed = (16384.0 + 128.0 * bn / cg) / 51.0;
    const double cost = static_cast<double>(waves) * ((mma > feed ? mma : feed) + 40.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

struct Workspace {
  __half* fh;      // [B, N, dk]
  __half* vh;      // [B, dv, N]
  __half* p;       // [B, N, N]
  float2* part;    // [B, ceil(N/kMinBN), N]
  float2* stats;   // [B, N]   (copy of row_stats, used by acan_attn_rows / loss_fused)
  float* bins;     // [B, N]
  float* logbin;   // [B, N]
  float* lnz;      // [B, N]
  float* kl_part;  // [B, 2*ceil(N/kMinBN), N]   (also the l partials of the un-normalised probability pass)
  float* diag;     // [B, N]  S_ii * c
  float* inv_l;    // [B, N]
  float* dpart;    // [B, ceil(N/128)]  per-CTA maxima of the diagonal (diag_kernel)
  float* lpart;    // [B, 2*ceil(N/kMinBN), N]  training path: partial row sums of the stored (rounded) probabilities
  float* lsum;     // [B, N]  their total: the divisor that makes the stored rows sum to exactly 1 (aggregation, backward)
  float* klw_part; // [B, 2*ceil(N/kMinBN), N]  partial in-window target mass (attention loss)
  float* klT;      // [B, N]  T_i = sum_j W_ij [W_ij / Q_ij inside the clamp window]: left by the loss forward for the backward
  int* done;       // [B, ceil(N/256)]  fused forward: completed probability-tile epilogues per 256-row block
  size_t bytes;
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

Workspace carve(void* base, int B, int N, int dk, int dv) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return reinterpret_cast<uint8_t*>(base) + o; };
  const size_t nt_max = 2 * ((N + kMinBN - 1) / kMinBN);      // one partial per (key tile, epilogue group)
  w.fh = reinterpret_cast<__half*>(take(static_cast<size_t>(B) * N * dk * 2));
  w.vh = reinterpret_cast<__half*>(take(static_cast<size_t>(B) * dv * N * 2));
  w.p = reinterpret_cast<__half*>(take(static_cast<size_t>(B) * N * N * 2));
  w.part = reinterpret_cast<float2*>(take(static_cast<size_t>(B) * nt_max * N * 8));
  w.stats = reinterpret_cast<float2*>(take(static_cast<size_t>(B) * N * 8));
  w.bins = reinterpret_cast<float*>(take(static_cast<size_t>(B) * N * 4));
  w.logbin = reinterpret_cast<float*>(take(static_cast<size_t>(B) * N * 4));
  w.lnz = reinterpret_cast<float*>(take(static_cast<size_t>(B) * N * 4));
  w.kl_part = reinterpret_cast<float*>(take(static_cast<size_t>(B) * nt_max * N * 4));
  w.diag = reinterpret_cast<float*>(take(static_cast<size_t>(B) * N * 4));
  w.inv_l = reinterpret_cast<float*>(take(static_cast<size_t>(B) * N * 4));
  w.dpart = reinterpret_cast<float*>(take(static_cast<size_t>(B) * ((N + kDiagRows - 1) / kDiagRows) * 4));
  w.lpart = reinterpret_cast<float*>(take(static_cast<size_t>(B) * nt_max * N * 4));
  w.lsum = reinterpret_cast<float*>(take(static_cast<size_t>(B) * N * 4));
  w.klw_part = reinterpret_cast<float*>(take(static_cast<size_t>(B) * nt_max * N * 4));
  w.klT = reinterpret_cast<float*>(take(static_cast<size_t>(B) * N * 4));
  w.done = reinterpret_cast<int*>(take(static_cast<size_t>(B) * ((N + 255) / 256) * 4));
  w.bytes = off;
  return w;
}

template <int MODE, int CG, int AM, bool BMN>
int launch_gemm_cg(const CUtensorMap& ta, const CUtensorMap& ta2, const CUtensorMap& tb, const CUtensorMap& tout, GemmArgs& g, cudaStream_t st) {
  const size_t stage_bytes = static_cast<size_t>(BM) * BK * 2 + static_cast<size_t>(g.bn / CG) * BK * 2;
  int stages = static_cast<int>((kSmemBudget - kOutBuffers * kOutChunkBytes) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  g.stages = stages;
  const size_t smem = stages * stage_bytes + kOutBuffers * kOutChunkBytes + sizeof(PipeBarriers) + 1024;
  {  // opt-in to > 48 KB of dynamic shared memory: a per-device, per-function attribute (this static is per kernel instantiation;
     // one bit per device ordinal)
    static std::atomic<uint64_t> done{0};
    int dev = 0;
    ACAN_CUDA(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (!(done.load(std::memory_order_acquire) & bit)) {
      ACAN_CUDA((cudaFuncSetAttribute(gemm_kernel<MODE, CG, AM, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
      done.fetch_or(bit, std::memory_order_release);
    }
  }
  if (g.out_scale == 0.f) g.out_scale = 1.0f;
  if (g.p_scale == 0.f) g.p_scale = 1.0f;
  if (AM != 2) g.k_half = g.k_steps;
  const int tiles = g.batch * g.m_tiles * g.n_tiles;
  const int slots = kNumSMs / CG;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>((tiles < slots ? tiles : slots) * CG));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g.pdl && use_pdl()) ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_kernel<MODE, CG, AM, BMN>, ta, ta2, tb, tout, g);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail((int)e, "gemm_kernel launch: %s", cudaGetErrorString(e)); }
  return launch_status("acan gemm_kernel");
}

template <int MODE, int AM = 0, bool BMN = (MODE == MODE_PVT)>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, GemmArgs& g, cudaStream_t st, int cg, const CUtensorMap* ta2 = nullptr) {
  const CUtensorMap& a2 = ta2 ? *ta2 : ta;
  return cg == 2 ? launch_gemm_cg<MODE, 2, AM, BMN>(ta, a2, tb, tout, g, st) : launch_gemm_cg<MODE, 1, AM, BMN>(ta, a2, tb, tout, g, st);
}

int check_attn_shape(const char* who, int B, int N, int dk, int dv) {
  if (B <= 0 || N < 8 || (N % 8) != 0) return fail(ACAN_ERR_BAD_ARG, "%s: need B > 0 and N a positive multiple of 8 (B=%d N=%d)", who, B, N);
  if (dk <= 0 || (dk % 64) != 0) return fail(ACAN_ERR_BAD_ARG, "%s: dk=%d must be a positive multiple of 64", who, dk);
  if (dv < 0 || (dv % 128) != 0) return fail(ACAN_ERR_BAD_ARG, "%s: dv=%d must be a multiple of 128", who, dv);
  if (static_cast<int64_t>(B) * N > 0x7fffffffLL / 2) return fail(ACAN_ERR_BAD_ARG, "%s: B*N too large", who);
  return 0;
}

// passes 1+2 over S = Fh Fh^T: statistics, then probabilities / sim_map / fused KL
// keep: training path -- the stored probabilities are fp16(P * kKeepScale) and their row sums go to w.lpart (2 * n_tiles slices per row)
int run_affinity(const Workspace& w, const __half* fh, float2* stats, bool want_p, float* sim_map, bool kl, int B, int N, int dk, float scale, cudaStream_t st,
                 bool need_stats, int* key_tiles, float* klw_part = nullptr, bool keep = false) {
  const int cg = cta_group();
  const int m_tiles = (N + BM * cg - 1) / (BM * cg);
  const int bn = env_int("ACAN_BN_QK", choose_bn(N, B * m_tiles, cg, 64));   // P leaves in 64-column (128 B) TMA-store boxes
  CUtensorMap ta, tb, tp;
  if (int e = make_tmap(&ta, fh, 2, dk, N, B, static_cast<size_t>(dk) * 2, static_cast<size_t>(N) * dk * 2, BM)) return e;
  if (int e = make_tmap(&tb, fh, 2, dk, N, B, static_cast<size_t>(dk) * 2, static_cast<size_t>(N) * dk * 2, bn / cg)) return e;
  if (int e = make_tmap(&tp, w.p, 2, N, N, B, static_cast<size_t>(N) * 2, static_cast<size_t>(N) * N * 2, BM)) return e;
  GemmArgs g;
  std::memset(&g, 0, sizeof(g));
  g.batch = B; g.m_tiles = m_tiles; g.n_tiles = (N + bn - 1) / bn; g.k_steps = dk / BK;
  g.m_rows = N; g.n_cols = N; g.bn = bn; g.idesc = umma_idesc_f16(bn, BM * cg); g.c = scale * kLog2e;
  g.skip_epilogue = env_int("ACAN_SKIP_EPI", 0);
  if (need_stats) {
    ProfScope ps(PROF_STATS, st);
    g.part = w.part;
    if (int e = launch_gemm<MODE_STATS>(ta, tb, tp, g, st, cg)) return e;
    combine_stats_kernel<<<(B * N + 255) / 256, 256, 0, st>>>(w.part, stats, stats != w.stats ? w.stats : nullptr, 2 * g.n_tiles, N, B);
    if (int e = launch_status("combine_stats_kernel")) return e;
  }
  if (want_p || sim_map || kl) {
    g.stats = stats;
    g.store_p = want_p ? 1 : 0;
    g.sim_out = sim_map;
    g.logbin = kl ? w.logbin : nullptr;
    g.lnz = w.lnz;
    g.kl_part = w.kl_part;
    g.klw_part = klw_part;
    if (keep && want_p) {
      g.p_scale = kKeepScale; g.lpart = w.lpart;
      if (!kl && !sim_map) { g.unnorm = 2; g.pdl = 1; }     // plain probability pass: the light epilogue, chained to the statistics kernels
    }
#ifdef ACAN_EXPERIMENT
    static long long* trace_buf = nullptr;
    if (env_int("ACAN_TRACE", 0) && !trace_buf) cudaMalloc(&trace_buf, 512 * sizeof(long long));
    g.trace = env_int("ACAN_TRACE", 0) ? trace_buf : nullptr;
#endif
    {
      ProfScope ps(PROF_PROBS, st);
      if (int e = launch_gemm<MODE_PROBS>(ta, tb, tp, g, st, cg)) return e;
    }
#ifdef ACAN_EXPERIMENT
    if (g.trace) {
      long long h[512];
      cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
      int n = (g.batch * g.m_tiles * g.n_tiles + kNumSMs / cg - 1) / (kNumSMs / cg);
      if (n > 12) n = 12;
      fprintf(stderr, "PROBS trace, CTA 0 (cycles rel. to first MMA start): tile: mma_start mma_issued | epilogue start end\n");
      for (int i = 0; i < n; ++i)
        fprintf(stderr, "  %2d: %7lld %7lld | %7lld %7lld\n", i, h[i * 4] - h[0], h[i * 4 + 1] - h[0], h[i * 4 + 2] - h[0], h[i * 4 + 3] - h[0]);
      g.trace = nullptr;
    }
#endif
  }
  *key_tiles = 2 * g.n_tiles;   // kl_part slices per row (key tile x epilogue group): callers reduce over them
  return 0;
}

// Inference fast path over S = Fh Fh^T: no statistics pass.  Row shifts come from the Cauchy-Schwarz bound (diag_kernel);
// the probability pass writes un-normalised fp16 p~ and per-row partial sums of exactly those rounded values; the
// aggregation pass adds them up and divides (combine_l: by a separate kernel, for the fp32-NCHW aggregation pass whose
// accumulator columns are the queries).  Images whose diagonal is too large for the bound still get the exact row maximum:
// the statistics kernel is always launched, skips the tiles of safe images on the device and exits in its first
// instructions when every image is safe -- no host synchronisation either way.
int run_affinity_fast(const Workspace& w, const __half* fh, float2* stats, int B, int N, int dk, float scale, cudaStream_t st, bool combine_l, int* l_parts) {
  const int cg = cta_group();
  const int m_tiles = (N + BM * cg - 1) / (BM * cg);
  const int bn = env_int("ACAN_BN_QK", choose_bn(N, B * m_tiles, cg, 64));
  CUtensorMap ta, tb, tp;
  if (int e = make_tmap(&ta, fh, 2, dk, N, B, static_cast<size_t>(dk) * 2, static_cast<size_t>(N) * dk * 2, BM)) return e;
  if (int e = make_tmap(&tb, fh, 2, dk, N, B, static_cast<size_t>(dk) * 2, static_cast<size_t>(N) * dk * 2, bn / cg)) return e;
  if (int e = make_tmap(&tp, w.p, 2, N, N, B, static_cast<size_t>(N) * 2, static_cast<size_t>(N) * N * 2, BM)) return e;
  GemmArgs g;
  std::memset(&g, 0, sizeof(g));
  g.batch = B; g.m_tiles = m_tiles; g.n_tiles = (N + bn - 1) / bn; g.k_steps = dk / BK;
  g.m_rows = N; g.n_cols = N; g.bn = bn; g.idesc = umma_idesc_f16(bn, BM * cg); g.c = scale * kLog2e;
  g.skip_epilogue = env_int("ACAN_SKIP_EPI", 0);
  g.part = w.part;
  g.dpart = w.dpart;
  g.n_dpart = (N + kDiagRows - 1) / kDiagRows;
  {
    ProfScope ps(PROF_STATS, st);
    diag_kernel<<<dim3(g.n_dpart, B), kDiagThreads, 0, st>>>(fh, w.diag, w.dpart, N, dk, g.c);
    if (int e = launch_status("diag_kernel")) return e;
    g.pdl = 1;
    if (int e = launch_gemm<MODE_STATS>(ta, tb, tp, g, st, cg)) return e;     // exits at once unless an image is above the safe bound
  }
  {
    ProfScope ps(PROF_PROBS, st);
    g.diag = w.diag;
    g.stats_w = stats;
    g.stats_w2 = stats != w.stats ? w.stats : nullptr;
    g.store_p = 1;
    g.unnorm = 1;
    g.lpart = w.kl_part;
    if (int e = launch_gemm<MODE_PROBS>(ta, tb, tp, g, st, cg)) return e;
    if (combine_l) {
      const int64_t rows = static_cast<int64_t>(B) * N;
      combine_l_kernel<<<static_cast<int>((rows + 255) / 256), 256, 0, st>>>(w.kl_part, stats, g.stats_w2, w.inv_l, 2 * g.n_tiles, N, rows);
      if (int e = launch_status("combine_l_kernel")) return e;
    }
  }
  *l_parts = 2 * g.n_tiles;
  return 0;
}

// Inference forward as ONE persistent launch behind the diagonal / fallback-statistics kernels (fused_fwd_kernel): probability and
// aggregation tiles interleaved, CTA pairs, 256 x 256 tiles.  V and O are channels-last 16-bit, dv a multiple of 256.
int run_fused_forward(const Workspace& w, const __half* fh, const void* value, void* ctx, bool out_bf16, float2* stats, int B, int N, int dk, int dv,
                      float scale, cudaStream_t st) {
  constexpr int BN = 256;
  const int mP = (N + 255) / 256, nP = (N + BN - 1) / BN, nV = dv / BN;
  CUtensorMap tf, tfb, tp, tv, to;
  if (int e = make_tmap(&tf, fh, 2, dk, N, B, static_cast<size_t>(dk) * 2, static_cast<size_t>(N) * dk * 2, BM)) return e;
  if (int e = make_tmap(&tp, w.p, 2, N, N, B, static_cast<size_t>(N) * 2, static_cast<size_t>(N) * N * 2, BM)) return e;
  if (int e = make_tmap(&tv, value, 2, dv, N, B, static_cast<size_t>(dv) * 2, static_cast<size_t>(N) * dv * 2, 64)) return e;
  if (int e = make_tmap(&to, ctx, 2, dv, N, B, static_cast<size_t>(dv) * 2, static_cast<size_t>(N) * dv * 2, BM, out_bf16)) return e;
  tfb = tf;
  const float c = scale * kLog2e;
  const int n_dpart = (N + kDiagRows - 1) / kDiagRows;
  {
    ProfScope ps(PROF_STATS, st);
    ACAN_CUDA(cudaMemsetAsync(w.done, 0, static_cast<size_t>(B) * mP * sizeof(int), st));
    diag_kernel<<<dim3(n_dpart, B), kDiagThreads, 0, st>>>(fh, w.diag, w.dpart, N, dk, c);
    if (int e = launch_status("diag_kernel")) return e;
    // fallback statistics (exact row maxima of images whose diagonal exceeds the safe bound): same 256-wide key tiles as the fused kernel
    GemmArgs g;
    std::memset(&g, 0, sizeof(g));
    g.batch = B; g.m_tiles = mP; g.n_tiles = nP; g.k_steps = dk / BK;
    g.m_rows = N; g.n_cols = N; g.bn = BN; g.idesc = umma_idesc_f16(BN, BM * 2); g.c = c;
    g.part = w.part; g.dpart = w.dpart; g.n_dpart = n_dpart; g.pdl = 1;
    if (int e = launch_gemm<MODE_STATS>(tf, tfb, tp, g, st, 2)) return e;
  }
  FusedArgs f;
  std::memset(&f, 0, sizeof(f));
  f.batch = B; f.n = N; f.dv = dv; f.mP = mP; f.nP = nP; f.nV = nV; f.kP = dk / BK; f.kV = (N + BK - 1) / BK;
  f.idesc_p = umma_idesc_f16(BN, BM * 2); f.idesc_v = umma_idesc_f16(BN, BM * 2, true);
  f.c = c; f.dpart = w.dpart; f.n_dpart = n_dpart; f.diag = w.diag; f.part = w.part;
  f.stats_w = stats; f.stats_w2 = stats != w.stats ? w.stats : nullptr;
  f.lpart = w.kl_part; f.done = w.done; f.out_bf16 = out_bf16 ? 1 : 0;
  const size_t stage_bytes = static_cast<size_t>(BM) * BK * 2 + 128 * BK * 2;
  int stages = static_cast<int>((kSmemBudget - kOutBuffers * kOutChunkBytes) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  f.stages = stages;
  const size_t smem = stages * stage_bytes + kOutBuffers * kOutChunkBytes + sizeof(PipeBarriers) + 1024;
  {
    static std::atomic<uint64_t> done{0};
    int dev = 0;
    ACAN_CUDA(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (!(done.load(std::memory_order_acquire) & bit)) {
      ACAN_CUDA((cudaFuncSetAttribute(fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)));
      done.fetch_or(bit, std::memory_order_release);
    }
  }
  const int tiles = B * mP * (nP + nV), slots = kNumSMs / 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>((tiles < slots ? tiles : slots) * 2));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl() ? 2 : 1;
  ProfScope ps(PROF_PV, st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, fused_fwd_kernel, tf, tp, tv, to, f);
  if (e != cudaSuccess) { (void)cudaGetLastError(); return fail((int)e, "fused_fwd_kernel launch: %s", cudaGetErrorString(e)); }
  return launch_status("fused_fwd_kernel");
}

int prepare_target(const Workspace& w, const float* dis_label, int B, int H, int Wd, cudaStream_t st) {
  const int N = (H / 8) * (Wd / 8);
  nearest_bins_kernel<<<(B * N + 255) / 256, 256, 0, st>>>(dis_label, w.bins, B, H, Wd);
  if (int e = launch_status("nearest_bins_kernel")) return e;
  gt_norm_kernel<<<dim3((N + 7) / 8, B), 256, static_cast<size_t>(N) * sizeof(float), st>>>(w.bins, w.logbin, w.lnz, N);
  return launch_status("gt_norm_kernel");
}

}  // namespace
}  // namespace acan

using namespace acan;

extern "C" size_t acan_attn_workspace_bytes(int B, int N, int dk, int dv) {
  if (B <= 0 || N <= 0 || dk <= 0 || dv < 0) return 0;
  return carve(nullptr, B, N, dk, dv).bytes;
}

static int attn_forward_impl(const float* feat, const float* value, float* ctx, float* sim_map, float* row_stats,
                             const float* dis_label, int H, int Wd, float* loss_out,
                             void* workspace, size_t workspace_bytes, int B, int N, int dk, int dv, float scale, void* stream) {
  ACAN_CHECK_ARG(feat && row_stats && workspace, "acan_attn_fwd: null pointer");
  ACAN_CHECK_ARG((value == nullptr) == (ctx == nullptr), "acan_attn_fwd: value and ctx must be given together");
  if (int e = check_attn_shape("acan_attn_fwd", B, N, dk, value ? dv : 0)) return e;
  ACAN_CHECK_ARG(!(dis_label) || (loss_out && (H / 8) * (Wd / 8) == N && H % 8 == 0 && Wd % 8 == 0), "acan_attn_fwd: label [B,1,%d,%d] does not match N=%d", H, Wd, N);
  ACAN_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 1024 == 0, "acan_attn_fwd: workspace must be 1024-byte aligned");
  const Workspace w = carve(workspace, B, N, dk, dv);
  if (workspace_bytes < w.bytes) return fail(ACAN_ERR_WORKSPACE, "acan_attn_fwd: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
  if (int e = acan_device_check()) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float2* stats = reinterpret_cast<float2*>(row_stats);

  {
    ProfScope ps(PROF_PACK, st);
    pack_feat_kernel<<<dim3((N + 63) / 64, dk / 64, B), 256, 0, st>>>(feat, w.fh, dk, N);
    if (int e = launch_status("pack_feat_kernel")) return e;
    if (value) {
      const int64_t n8 = static_cast<int64_t>(B) * dv * N / 8;
      int64_t blocks = (n8 + 255) / 256;
      if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
      pack_value_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(reinterpret_cast<const float4*>(value), reinterpret_cast<uint4*>(w.vh), n8);
      if (int e = launch_status("pack_value_kernel")) return e;
    }
  }
  if (dis_label)
    if (int e = prepare_target(w, dis_label, B, H, Wd, st)) return e;

  int nt = 0;
  const bool fast = value && !sim_map && !dis_label && !env_int("ACAN_EXACT_STATS", 0);   // inference: no statistics pass
  int l_parts = 0;
  if (fast) {
    if (int e = run_affinity_fast(w, w.fh, stats, B, N, dk, scale, st, true, &l_parts)) return e;
  } else {
    if (int e = run_affinity(w, w.fh, stats, value != nullptr, sim_map, dis_label != nullptr, B, N, dk, scale, st, true, &nt)) return e;
  }
  if (dis_label) {
    ordered_sum_kernel<<<1, 1024, 0, st>>>(w.kl_part, static_cast<int64_t>(B) * nt * N, 1.0f / (static_cast<float>(B) * N), loss_out);
    if (int e = launch_status("ordered_sum_kernel")) return e;
  }
  if (value) {
    // pass 3: O^T[dv, Nq] = V[dv, Nk] * P[Nq, Nk]^T
    const int cg = cta_group();
    const int m_tiles = (dv + BM * cg - 1) / (BM * cg);
    const int bn = env_int("ACAN_BN_PV", choose_bn(N, B * m_tiles, cg, 32));   // O leaves in 32-column (128 B of fp32) TMA-store boxes
    CUtensorMap ta, tb, to;
    if (int e = make_tmap(&ta, w.vh, 2, N, dv, B, static_cast<size_t>(N) * 2, static_cast<size_t>(dv) * N * 2, BM)) return e;
    if (int e = make_tmap(&tb, w.p, 2, N, N, B, static_cast<size_t>(N) * 2, static_cast<size_t>(N) * N * 2, bn / cg)) return e;
    if (int e = make_tmap(&to, ctx, 4, N, dv, B, static_cast<size_t>(N) * 4, static_cast<size_t>(dv) * N * 4, BM)) return e;
    GemmArgs g;
    std::memset(&g, 0, sizeof(g));
    g.batch = B; g.m_tiles = m_tiles; g.n_tiles = (N + bn - 1) / bn; g.k_steps = (N + BK - 1) / BK;
    g.m_rows = dv; g.n_cols = N; g.bn = bn; g.idesc = umma_idesc_f16(bn, BM * cg); g.c = 0.f;
    g.skip_epilogue = env_int("ACAN_SKIP_EPI", 0);
    g.inv_l = fast ? w.inv_l : nullptr;
    ProfScope ps(PROF_PV, st);
    if (int e = launch_gemm<MODE_PV>(ta, tb, to, g, st, cg)) return e;
  }
  return 0;
}

extern "C" int acan_attn_fwd(const float* feat, const float* value, float* ctx, float* sim_map, float* row_stats,
                             void* workspace, size_t workspace_bytes, int B, int N, int dk, int dv, float scale, void* stream) {
  return attn_forward_impl(feat, value, ctx, sim_map, row_stats, nullptr, 0, 0, nullptr, workspace, workspace_bytes, B, N, dk, dv, scale, stream);
}

extern "C" int acan_attn_fwd_loss(const float* feat, const float* value, float* ctx, float* sim_map, float* row_stats,
                                  const float* dis_label, int H, int W, float* loss_out,
                                  void* workspace, size_t workspace_bytes, int B, int N, int dk, int dv, float scale, void* stream) {
  ACAN_CHECK_ARG(dis_label && loss_out, "acan_attn_fwd_loss: null label / loss pointer");
  return attn_forward_impl(feat, value, ctx, sim_map, row_stats, dis_label, H, W, loss_out, workspace, workspace_bytes, B, N, dk, dv, scale, stream);
}

// channels-last form: F and V are consumed where the convolutions left them (fp16, or bf16 for V), O is written channels-last.
//   keep = 0 (inference): statistics-free probabilities when no sim_map / loss is requested (run_affinity_fast).
//   keep = 1 (training):  exact two-pass statistics; the normalised probabilities stay in the workspace as fp16(P * 2^10) together
//                         with the row sums of those rounded values (lsum) -- the divisor used by the aggregation here and by
//                         acan_attn_bwd_nhwc, so that the rows the tensor cores read sum to exactly 1.
extern "C" int acan_attn_fwd_nhwc(const void* feat, int feat_dtype, const void* value, int value_dtype, void* ctx, int ctx_dtype,
                                  float* sim_map, float* row_stats, const float* dis_label, int H, int Wd, float* loss_out,
                                  void* workspace, size_t workspace_bytes, int B, int N, int dk, int dv, float scale, int keep, void* stream) {
  ACAN_CHECK_ARG(feat && row_stats && workspace, "acan_attn_fwd_nhwc: null pointer");
  ACAN_CHECK_ARG((value == nullptr) == (ctx == nullptr), "acan_attn_fwd_nhwc: value and ctx must be given together");
  ACAN_CHECK_ARG(feat_dtype == DT_F16 || feat_dtype == DT_BF16 || feat_dtype == DT_F32, "acan_attn_fwd_nhwc: feat_dtype %d", feat_dtype);
  // (an fp16 A with a bf16 B is NOT a legal kind::f16 MMA on sm_100a although the descriptor has one format field per operand: measured,
  //  the instruction traps as illegal -- so V shares the probabilities' format)
  ACAN_CHECK_ARG(!value || ((value_dtype == DT_F16 || value_dtype == DT_BF16) && (ctx_dtype == DT_F16 || ctx_dtype == DT_BF16 || ctx_dtype == DT_F32)),
                 "acan_attn_fwd_nhwc: value must be fp16 / bf16 (got %d), ctx fp16 / bf16 / fp32 (got %d)", value_dtype, ctx_dtype);
  if (int e = check_attn_shape("acan_attn_fwd_nhwc", B, N, dk, value ? dv : 0)) return e;
  ACAN_CHECK_ARG(!(dis_label) || (loss_out && (H / 8) * (Wd / 8) == N && H % 8 == 0 && Wd % 8 == 0), "acan_attn_fwd_nhwc: label [B,1,%d,%d] does not match N=%d", H, Wd, N);
  ACAN_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 1024 == 0 && reinterpret_cast<uintptr_t>(feat) % 16 == 0 &&
                 reinterpret_cast<uintptr_t>(value) % 16 == 0 && reinterpret_cast<uintptr_t>(ctx) % 16 == 0,
                 "acan_attn_fwd_nhwc: workspace must be 1024-byte, tensors 16-byte aligned");
  const Workspace w = carve(workspace, B, N, dk, dv);
  if (workspace_bytes < w.bytes) return fail(ACAN_ERR_WORKSPACE, "acan_attn_fwd_nhwc: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
  if (int e = acan_device_check()) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope core_scope(PROF_CORE, st);
  float2* stats = reinterpret_cast<float2*>(row_stats);
  const __half* fh = static_cast<const __half*>(feat);
  if (feat_dtype != DT_F16) {          // bf16 (autocast) / fp32 features: one elementwise pass into the workspace pack
    const int64_t n8 = static_cast<int64_t>(B) * N * dk / 8;
    int64_t blocks = (n8 + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    if (feat_dtype == DT_BF16) to_half_kernel<true><<<static_cast<int>(blocks), 256, 0, st>>>(feat, reinterpret_cast<uint4*>(w.fh), n8);
    else                       to_half_kernel<false><<<static_cast<int>(blocks), 256, 0, st>>>(feat, reinterpret_cast<uint4*>(w.fh), n8);
    if (int e = launch_status("to_half_kernel")) return e;
    fh = w.fh;
  }
  if (value && value_dtype == DT_BF16) {   // bf16 values (the autocast step): one exact conversion pass into the workspace's fp16 slot
    const int64_t n8 = static_cast<int64_t>(B) * N * dv / 8;
    int64_t blocks = (n8 + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    to_half_kernel<true><<<static_cast<int>(blocks), 256, 0, st>>>(value, reinterpret_cast<uint4*>(w.vh), n8);
    if (int e = launch_status("to_half_kernel (V)")) return e;
    value = w.vh;
    value_dtype = DT_F16;
  }
  if (dis_label)
    if (int e = prepare_target(w, dis_label, B, H, Wd, st)) return e;
  int nt = 0;
  const bool fast = keep != 1 && value && !sim_map && !dis_label && !env_int("ACAN_EXACT_STATS", 0);   // inference: no statistics pass
  // keep = 2: the single persistent launch (fused_fwd_kernel).  Bit-identical results, measured SLOWER than the two launches (165 vs 148 us
  // at configs[1]: both passes are bound by the same resource, see the kernel's header), so it is opt-in, not the default.
  if (fast && keep == 2 && cta_group() == 2 && dv % 256 == 0 && ctx_dtype != DT_F32 && B <= 4096)
    return run_fused_forward(w, fh, value, ctx, ctx_dtype == DT_BF16, stats, B, N, dk, dv, scale, st);
  if (keep != 1) keep = 0;
  int l_parts = 0;
  if (fast) {
    if (int e = run_affinity_fast(w, fh, stats, B, N, dk, scale, st, false, &l_parts)) return e;
  } else {
    if (int e = run_affinity(w, fh, stats, value != nullptr, sim_map, dis_label != nullptr, B, N, dk, scale, st, true, &nt,
                             (keep && dis_label) ? w.klw_part : nullptr, keep != 0)) return e;
    l_parts = nt;
  }
  if (dis_label) {
    ordered_sum_kernel<<<1, 1024, 0, st>>>(w.kl_part, static_cast<int64_t>(B) * nt * N, 1.0f / (static_cast<float>(B) * N), loss_out);
    if (int e = launch_status("ordered_sum_kernel")) return e;
    if (keep) {
      const int64_t rows = static_cast<int64_t>(B) * N;
      sum_parts_kernel<<<static_cast<int>((rows + 255) / 256), 256, 0, st>>>(w.klw_part, w.klT, nt, N, rows);
      if (int e = launch_status("sum_parts_kernel")) return e;
    }
  }
  if (value) {
    // O[Nq, dv] = P[Nq, Nk] * V[Nk, dv]: A = P (K-major), B = V in place (MN-major: channels contiguous)
    const int cg = cta_group();
    const int m_tiles = (N + BM * cg - 1) / (BM * cg);
    const int bn = env_int("ACAN_BN_PV", choose_bn(dv, B * m_tiles, cg, 64 * cg, 64 * cg));   // 64-channel boxes per CTA
    ACAN_CHECK_ARG(dv % 64 == 0 && bn % (64 * cg) == 0, "acan_attn_fwd_nhwc: dv=%d / tile width %d not a multiple of 64 per CTA", dv, bn);
    const bool vb = value_dtype == DT_BF16;
    const int ob = ctx_dtype == DT_F32 ? 4 : 2;
    CUtensorMap ta, tb, to;
    if (int e = make_tmap(&ta, w.p, 2, N, N, B, static_cast<size_t>(N) * 2, static_cast<size_t>(N) * N * 2, BM)) return e;
    if (int e = make_tmap(&tb, value, 2, dv, N, B, static_cast<size_t>(dv) * 2, static_cast<size_t>(N) * dv * 2, 64, vb)) return e;
    if (int e = make_tmap(&to, ctx, ob, dv, N, B, static_cast<size_t>(dv) * ob, static_cast<size_t>(N) * dv * ob, BM, ctx_dtype == DT_BF16)) return e;
    GemmArgs g;
    std::memset(&g, 0, sizeof(g));
    g.batch = B; g.m_tiles = m_tiles; g.n_tiles = (dv + bn - 1) / bn; g.k_steps = (N + BK - 1) / BK;
    g.m_rows = N; g.n_cols = dv; g.bn = bn; g.idesc = umma_idesc_f16(bn, BM * cg, true, false, false, vb); g.c = 0.f;
    g.skip_epilogue = env_int("ACAN_SKIP_EPI", 0);
    g.out_bf16 = ctx_dtype == DT_BF16;
    if (fast) {
      g.l_in = w.kl_part; g.l_parts = l_parts;
      g.stats_w = stats; g.stats_w2 = stats != w.stats ? w.stats : nullptr;
      g.pdl = 1;                         // directly behind the probability pass
    } else if (keep) {
      g.l_in = w.lpart; g.l_parts = l_parts;
      g.lsum_out = w.lsum;
      g.pdl = dis_label ? 0 : 1;
    }
    ProfScope ps(PROF_PV, st);
    if (ctx_dtype == DT_F32) { if (int e = launch_gemm<MODE_PVT32, 0, true>(ta, tb, to, g, st, cg)) return e; }
    else                     { if (int e = launch_gemm<MODE_PVT, 0, true>(ta, tb, to, g, st, cg)) return e; }
  }
  return 0;
}

extern "C" int acan_attn_fwd_nhwc_f16(const void* feat_h, const void* value_h, void* ctx_h, float* sim_map, float* row_stats,
                                      const float* dis_label, int H, int Wd, float* loss_out,
                                      void* workspace, size_t workspace_bytes, int B, int N, int dk, int dv, float scale, void* stream) {
  return acan_attn_fwd_nhwc(feat_h, DT_F16, value_h, DT_F16, ctx_h, DT_F16, sim_map, row_stats, dis_label, H, Wd, loss_out,
                            workspace, workspace_bytes, B, N, dk, dv, scale, 0, stream);
}

extern "C" int acan_attn_rows(const void* workspace, const void* feat_pack, const int* row_idx_dev, int n_rows, float* rows_out,
                              int B, int N, int dk, int dv, float scale, void* stream) {
  ACAN_CHECK_ARG(workspace && row_idx_dev && rows_out && n_rows > 0, "acan_attn_rows: bad argument");
  if (int e = check_attn_shape("acan_attn_rows", B, N, dk, dv)) return e;
  const Workspace w = carve(const_cast<void*>(workspace), B, N, dk, dv);
  const __half* fh = feat_pack ? static_cast<const __half*>(feat_pack) : w.fh;
  sim_rows_kernel<<<dim3(n_rows, B), 256, dk * sizeof(float), static_cast<cudaStream_t>(stream)>>>(fh, w.stats, row_idx_dev, rows_out, n_rows, N, dk,
                                                                                                  scale * kLog2e);
  return launch_status("sim_rows_kernel");
}

extern "C" int acan_attention_loss_fused(void* workspace, const void* feat_pack, const float* dis_label, float* loss_out,
                                         int B, int H, int W, int dk, int dv, float scale, void* stream) {
  ACAN_CHECK_ARG(workspace && dis_label && loss_out && H >= 8 && W >= 8 && H % 8 == 0 && W % 8 == 0, "acan_attention_loss_fused: bad argument (H, W must be multiples of 8)");
  const int N = (H / 8) * (W / 8);
  if (int e = check_attn_shape("acan_attention_loss_fused", B, N, dk, dv)) return e;
  const Workspace w = carve(workspace, B, N, dk, dv);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int e = prepare_target(w, dis_label, B, H, W, st)) return e;
  int nt = 0;
  const __half* fh = feat_pack ? static_cast<const __half*>(feat_pack) : w.fh;
  // (the in-window target mass T_i is left in the workspace for acan_attn_bwd_nhwc)
  if (int e = run_affinity(w, fh, w.stats, false, nullptr, true, B, N, dk, scale, st, false, &nt, w.klw_part)) return e;
  ordered_sum_kernel<<<1, 1024, 0, st>>>(w.kl_part, static_cast<int64_t>(B) * nt * N, 1.0f / (static_cast<float>(B) * N), loss_out);
  if (int e = launch_status("ordered_sum_kernel")) return e;
  const int64_t rows = static_cast<int64_t>(B) * N;
  sum_parts_kernel<<<static_cast<int>((rows + 255) / 256), 256, 0, st>>>(w.klw_part, w.klT, nt, N, rows);
  return launch_status("sum_parts_kernel");
}

// ------------------------------------------------------------------------------------ backward
namespace acan {
namespace {

struct BwdWorkspace {
  Workspace f;             // forward-style scratch: fh, vh, p, part, stats, bins, logbin, lnz, kl_part, ...
  __half* doh;             // [B, dv, N] fp16   grad_ctx
  __half* ds;              // [B, N, N]  fp16   dS * s  (s = power-of-two gradient scale)
  __half* fb;              // [B, dk, N] fp16   F in its NCHW layout (A operand of the dF GEMMs)
  float* gscale;           // [2]  (s, 1/s)
  int* amax_bits;          // [1]  max |dO| as int bits
  float* klw_part;         // [B, 2*ceil(N/64), N]
  float* klT;              // [B, N]
  float* delta;            // [B, N]
  float* df1;              // [B, dk, N]
  float* df2;              // [B, dk, N]
  size_t bytes;
};

BwdWorkspace carve_bwd(void* base, int B, int N, int dk, int dv) {
  BwdWorkspace w;
  w.f = carve(base, B, N, dk, dv);
  size_t off = w.f.bytes;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return reinterpret_cast<uint8_t*>(base) + o; };
  const size_t nt_max = 2 * ((N + kMinBN - 1) / kMinBN);
  w.doh = reinterpret_cast<__half*>(take(static_cast<size_t>(B) * dv * N * 2));
  w.ds = reinterpret_cast<__half*>(take(static_cast<size_t>(B) * N * N * 2));
  w.fb = reinterpret_cast<__half*>(take(static_cast<size_t>(B) * dk * N * 2));
  w.gscale = reinterpret_cast<float*>(take(16));
  w.amax_bits = reinterpret_cast<int*>(take(16));
  w.klw_part = reinterpret_cast<float*>(take(static_cast<size_t>(B) * nt_max * N * 4));
  w.klT = reinterpret_cast<float*>(take(static_cast<size_t>(B) * N * 4));
  w.delta = reinterpret_cast<float*>(take(static_cast<size_t>(B) * N * 4));
  w.df1 = reinterpret_cast<float*>(take(static_cast<size_t>(B) * dk * N * 4));
  w.df2 = reinterpret_cast<float*>(take(static_cast<size_t>(B) * dk * N * 4));
  w.bytes = off;
  return w;
}

int cast_launch(const float* src, void* dst, int64_t n, const float* gscale, cudaStream_t st, const char* what) {
  const int64_t n8 = n / 8;
  int64_t blocks = (n8 + 255) / 256;
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  pack_cast_kernel<false><<<static_cast<int>(blocks), 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint4*>(dst), n8, gscale);
  return launch_status(what);
}

GemmArgs gemm_args(int B, int m_rows, int n_cols, int k, int bn, int cg, uint32_t idesc) {
  GemmArgs g;
  std::memset(&g, 0, sizeof(g));
  g.batch = B; g.m_tiles = (m_rows + BM * cg - 1) / (BM * cg); g.n_tiles = (n_cols + bn - 1) / bn; g.k_steps = (k + BK - 1) / BK;
  g.m_rows = m_rows; g.n_cols = n_cols; g.bn = bn; g.idesc = idesc;
  return g;
}

}  // namespace
}  // namespace acan

extern "C" size_t acan_attn_bwd_workspace_bytes(int B, int N, int dk, int dv) {
  if (B <= 0 || N <= 0 || dk <= 0 || dv <= 0) return 0;
  return carve_bwd(nullptr, B, N, dk, dv).bytes;
}

// Gradients of ctx = aggregate(softmax(F^T F scale), V) (+ grad_loss * attention loss) w.r.t. F and V, on tcgen05.
//   dV = dO P        dP = dO^T V        dS = P o (dP - delta) [+ loss term]        dF = scale * F (dS + dS^T)
extern "C" int acan_attn_bwd(const float* feat, const float* value, const float* ctx, const float* row_stats, const float* grad_ctx,
                             const float* dis_label, int H, int Wd, const float* grad_loss,
                             float* grad_feat, float* grad_value, void* workspace, size_t workspace_bytes,
                             int B, int N, int dk, int dv, float scale, void* stream) {
  ACAN_CHECK_ARG(feat && row_stats && grad_feat && workspace, "acan_attn_bwd: null pointer");
  const bool agg = grad_ctx != nullptr;           // aggregation gradient present (else: attention-loss gradient only)
  ACAN_CHECK_ARG(!agg || (value && ctx && grad_value), "acan_attn_bwd: value, ctx and grad_value are required with grad_ctx");
  ACAN_CHECK_ARG(agg || dis_label, "acan_attn_bwd: nothing to differentiate (no grad_ctx, no label)");
  ACAN_CHECK_ARG(!dis_label || grad_loss, "acan_attn_bwd: a label needs grad_loss (device scalar)");
  if (int e = check_attn_shape("acan_attn_bwd", B, N, dk, dv)) return e;
  ACAN_CHECK_ARG(N % 8 == 0 && (static_cast<int64_t>(dk) * N) % 8 == 0, "acan_attn_bwd: N must be a multiple of 8");
  ACAN_CHECK_ARG(!dis_label || ((H / 8) * (Wd / 8) == N && H % 8 == 0 && Wd % 8 == 0), "acan_attn_bwd: label [B,1,%d,%d] does not match N=%d", H, Wd, N);
  ACAN_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 1024 == 0, "acan_attn_bwd: workspace must be 1024-byte aligned");
  const BwdWorkspace w = carve_bwd(workspace, B, N, dk, dv);
  if (workspace_bytes < w.bytes) return fail(ACAN_ERR_WORKSPACE, "acan_attn_bwd: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
  if (int e = acan_device_check()) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cg = cta_group();
  const int64_t rows = static_cast<int64_t>(B) * N;
  const int grid_rows = static_cast<int>((rows + 255) / 256);

  // operand packs
  pack_feat_kernel<<<dim3((N + 63) / 64, dk / 64, B), 256, 0, st>>>(feat, w.f.fh, dk, N);
  if (int e = launch_status("pack_feat_kernel")) return e;
  if (int e = cast_launch(feat, w.fb, static_cast<int64_t>(B) * dk * N, nullptr, st, "pack F (NCHW fp16)")) return e;
  const float inv_bn = 1.0f / (static_cast<float>(B) * static_cast<float>(N));
  if (agg) {
    // delta_i and max |dO| in one pass over dO; the power-of-two scale s keeps fp16 dO and dS in range whatever the loss scale
    ACAN_CUDA(cudaMemsetAsync(w.amax_bits, 0, 4, st));
    delta_kernel<<<dim3((N + 63) / 64, B), 256, 0, st>>>(grad_ctx, ctx, w.delta, w.amax_bits, dv, N);
    if (int e = launch_status("delta_kernel")) return e;
    gscale_kernel<<<1, 1, 0, st>>>(w.amax_bits, nullptr, inv_bn, w.gscale);
    if (int e = launch_status("gscale_kernel")) return e;
    if (int e = cast_launch(value, w.f.vh, static_cast<int64_t>(B) * dv * N, nullptr, st, "pack V")) return e;
    if (int e = cast_launch(grad_ctx, w.doh, static_cast<int64_t>(B) * dv * N, w.gscale, st, "pack dO")) return e;
  } else {
    gscale_kernel<<<1, 1, 0, st>>>(nullptr, grad_loss, inv_bn, w.gscale);
    if (int e = launch_status("gscale_kernel")) return e;
  }
  // normalised P from the saved statistics (+ in-window target mass T_i for the loss term)
  ACAN_CUDA(cudaMemcpyAsync(w.f.stats, row_stats, static_cast<size_t>(rows) * 8, cudaMemcpyDeviceToDevice, st));
  if (dis_label)
    if (int e = prepare_target(w.f, dis_label, B, H, Wd, st)) return e;
  int parts = 0;
  if (int e = run_affinity(w.f, w.f.fh, w.f.stats, true, nullptr, dis_label != nullptr, B, N, dk, scale, st, false, &parts, w.klw_part)) return e;
  if (dis_label) {
    sum_parts_kernel<<<grid_rows, 256, 0, st>>>(w.klw_part, w.klT, parts, N, rows);
    if (int e = launch_status("sum_parts_kernel")) return e;
  }
  CUtensorMap ta, tb, to;

  if (agg) {
    {  // dV[c, j] = sum_i dO[c, i] P[i, j]: A = dOh (K-major over i), B = P read MN-major (j contiguous), fp32 NCHW out
      const int m_tiles = (dv + BM * cg - 1) / (BM * cg);
      const int bn = choose_bn(N, B * m_tiles, cg, 64 * cg, 64 * cg);
      if (int e = make_tmap(&ta, w.doh, 2, N, dv, B, static_cast<size_t>(N) * 2, static_cast<size_t>(dv) * N * 2, BM)) return e;
      if (int e = make_tmap(&tb, w.f.p, 2, N, N, B, static_cast<size_t>(N) * 2, static_cast<size_t>(N) * N * 2, 64)) return e;
      if (int e = make_tmap(&to, grad_value, 4, N, dv, B, static_cast<size_t>(N) * 4, static_cast<size_t>(dv) * N * 4, BM)) return e;
      GemmArgs g = gemm_args(B, dv, N, N, bn, cg, umma_idesc_f16(bn, BM * cg, true, false, false));
      g.gscale = w.gscale;
      if (int e = launch_gemm<MODE_PV, 0, true>(ta, tb, to, g, st, cg)) return e;
    }
    {  // dP[i, j] = sum_c dO[c, i] V[c, j] (both operands MN-major from their NCHW packs) -> dS epilogue, scaled fp16 out
      const int m_tiles = (N + BM * cg - 1) / (BM * cg);
      const int bn = choose_bn(N, B * m_tiles, cg, 64 * cg, 64 * cg);
      if (int e = make_tmap(&ta, w.doh, 2, N, dv, B, static_cast<size_t>(N) * 2, static_cast<size_t>(dv) * N * 2, 64)) return e;
      if (int e = make_tmap(&tb, w.f.vh, 2, N, dv, B, static_cast<size_t>(N) * 2, static_cast<size_t>(dv) * N * 2, 64)) return e;
      if (int e = make_tmap(&to, w.ds, 2, N, N, B, static_cast<size_t>(N) * 2, static_cast<size_t>(N) * N * 2, BM)) return e;
      GemmArgs g = gemm_args(B, N, N, dv, bn, cg, umma_idesc_f16(bn, BM * cg, true, true, false));
      g.ds_delta = w.delta; g.ds_p = w.f.p; g.ds_klT = w.klT; g.ds_gl = dis_label ? grad_loss : nullptr; g.ds_inv_bn = inv_bn; g.gscale = w.gscale;
      g.logbin = dis_label ? w.f.logbin : nullptr; g.lnz = w.f.lnz;
      if (int e = launch_gemm<MODE_DS, 1, true>(ta, tb, to, g, st, cg)) return e;
    }
  } else {
    kl_ds_kernel<<<static_cast<int>(rows), 256, 0, st>>>(w.f.p, w.f.logbin, w.f.lnz, w.klT, w.ds, grad_loss, inv_bn, w.gscale, N);
    if (int e = launch_status("kl_ds_kernel")) return e;
  }
  {  // dF[c, i] = scale / s * sum_j (dS[i, j] + dS[j, i]) F[c, j]: two fp16 GEMMs over the same (scaled) dS buffer
    const int m_tiles = (dk + BM * cg - 1) / (BM * cg);
    if (int e = make_tmap(&ta, w.fb, 2, N, dk, B, static_cast<size_t>(N) * 2, static_cast<size_t>(dk) * N * 2, BM)) return e;
    {
      const int bn = choose_bn(N, B * m_tiles, cg, 32);
      if (int e = make_tmap(&tb, w.ds, 2, N, N, B, static_cast<size_t>(N) * 2, static_cast<size_t>(N) * N * 2, bn / cg)) return e;
      if (int e = make_tmap(&to, w.df1, 4, N, dk, B, static_cast<size_t>(N) * 4, static_cast<size_t>(dk) * N * 4, BM)) return e;
      GemmArgs g = gemm_args(B, dk, N, N, bn, cg, umma_idesc_f16(bn, BM * cg, false, false, false));
      if (int e = launch_gemm<MODE_PV, 0, false>(ta, tb, to, g, st, cg)) return e;       // sum_j F[c,j] dS[i,j]
    }
    {
      const int bn = choose_bn(N, B * m_tiles, cg, 64 * cg, 64 * cg);
      if (int e = make_tmap(&tb, w.ds, 2, N, N, B, static_cast<size_t>(N) * 2, static_cast<size_t>(N) * N * 2, 64)) return e;
      if (int e = make_tmap(&to, w.df2, 4, N, dk, B, static_cast<size_t>(N) * 4, static_cast<size_t>(dk) * N * 4, BM)) return e;
      GemmArgs g = gemm_args(B, dk, N, N, bn, cg, umma_idesc_f16(bn, BM * cg, true, false, false));
      if (int e = launch_gemm<MODE_PV, 0, true>(ta, tb, to, g, st, cg)) return e;        // sum_j F[c,j] dS[j,i]
    }
    const int64_t n4 = static_cast<int64_t>(B) * dk * N / 4;
    int64_t blocks = (n4 + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    scaled_sum_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(reinterpret_cast<const float4*>(w.df1), reinterpret_cast<const float4*>(w.df2),
                                                                reinterpret_cast<float4*>(grad_feat), scale, w.gscale, n4);
    if (int e = launch_status("scaled_sum_kernel")) return e;
  }
  return 0;
}

// ------------------------------------------------------------------------------------ channels-last backward
namespace acan {
namespace {

struct BwdScratch {
  __half* ds;       // [B, N, N] fp16   dS * s_ds
  __half* doh;      // [B, N, dv] fp16  grad_ctx packed (fp32 grad_ctx only)
  float* delta;     // [B, N]
  float* gscale;    // [4]
  int* amax_bits;   // [1]
  size_t bytes;
};

BwdScratch carve_scratch(void* base, int B, int N, int dv) {
  BwdScratch w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return reinterpret_cast<uint8_t*>(base) + o; };
  w.ds = reinterpret_cast<__half*>(take(static_cast<size_t>(B) * N * N * 2));
  w.doh = reinterpret_cast<__half*>(take(static_cast<size_t>(B) * N * dv * 2));
  w.delta = reinterpret_cast<float*>(take(static_cast<size_t>(B) * N * 4));
  w.gscale = reinterpret_cast<float*>(take(32));
  w.amax_bits = reinterpret_cast<int*>(take(16));
  w.bytes = off;
  return w;
}

// accumulator rows = rows of A, row-major [rows, cols] output in fp32 / fp16 / bf16
template <int AM>
int launch_rowmajor(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, GemmArgs& g, cudaStream_t st, int cg, int out_dtype,
                    const CUtensorMap* ta2 = nullptr) {
  g.out_bf16 = out_dtype == DT_BF16;
  if (out_dtype == DT_F32) return launch_gemm<MODE_PVT32, AM, true>(ta, tb, to, g, st, cg, ta2);
  return launch_gemm<MODE_PVT, AM, true>(ta, tb, to, g, st, cg, ta2);
}

}  // namespace
}  // namespace acan

extern "C" size_t acan_attn_bwd_nhwc_scratch_bytes(int B, int N, int dk, int dv) {
  if (B <= 0 || N <= 0 || dk <= 0 || dv <= 0) return 0;
  return carve_scratch(nullptr, B, N, dv).bytes;
}

// Backward of acan_attn_fwd_nhwc(keep = 1) (+ grad_loss * attention loss when have_loss) w.r.t. F and V, channels-last, on tcgen05:
//   dV = P^T dO        dP = dO V^T        dS = p o (dP - delta) [+ loss term], p = P^ / lsum        dF = scale (dS + dS^T) F
// `workspace` is the forward's: it still holds P^ (fp16(P 2^10)), lsum, the row statistics and -- after acan_attention_loss_fused or a
// forward with a label -- ln(bin), ln Z and T_i.  V and an fp16 dO are read in place as MMA operands; a bf16 / fp32 grad_ctx is packed to
// fp16 times a power of two chosen on the device.  Three tensor-core launches (dV, dP -> dS, symmetrised dF), no P recompute.
extern "C" int acan_attn_bwd_nhwc(void* workspace, size_t workspace_bytes, void* scratch, size_t scratch_bytes,
                                  const void* feat_h, const void* value, int value_dtype, const float* ctx32,
                                  const void* grad_ctx, int go_dtype, int have_loss, const float* grad_loss,
                                  void* grad_feat, int gf_dtype, void* grad_value, int gv_dtype,
                                  int B, int N, int dk, int dv, float scale, void* stream) {
  ACAN_CHECK_ARG(workspace && scratch && grad_feat, "acan_attn_bwd_nhwc: null pointer");
  const bool agg = grad_ctx != nullptr;
  ACAN_CHECK_ARG(!agg || (ctx32 && grad_value), "acan_attn_bwd_nhwc: ctx and grad_value are required with grad_ctx");
  ACAN_CHECK_ARG(agg || have_loss, "acan_attn_bwd_nhwc: nothing to differentiate (no grad_ctx, no loss)");
  ACAN_CHECK_ARG(!have_loss || grad_loss, "acan_attn_bwd_nhwc: the loss term needs grad_loss (device scalar)");
  ACAN_CHECK_ARG(!agg || (value_dtype == DT_F16 && (go_dtype == DT_F16 || go_dtype == DT_BF16 || go_dtype == DT_F32)),
                 "acan_attn_bwd_nhwc: value must be fp16 (as given to the forward; NULL = the workspace's fp16 copy of a bf16 V), grad_ctx fp16 / bf16 / fp32");
  auto dt_ok = [](int d) { return d == DT_F32 || d == DT_F16 || d == DT_BF16; };
  ACAN_CHECK_ARG(dt_ok(gf_dtype) && (!agg || dt_ok(gv_dtype)), "acan_attn_bwd_nhwc: output dtype");
  if (int e = check_attn_shape("acan_attn_bwd_nhwc", B, N, dk, dv)) return e;
  ACAN_CHECK_ARG(dv % 64 == 0 && dk % 64 == 0, "acan_attn_bwd_nhwc: dk, dv must be multiples of 64");
  ACAN_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 1024 == 0 && reinterpret_cast<uintptr_t>(scratch) % 1024 == 0,
                 "acan_attn_bwd_nhwc: workspace and scratch must be 1024-byte aligned");
  const Workspace w = carve(workspace, B, N, dk, dv);
  const BwdScratch x = carve_scratch(scratch, B, N, dv);
  if (workspace_bytes < w.bytes) return fail(ACAN_ERR_WORKSPACE, "acan_attn_bwd_nhwc: workspace %zu < %zu bytes", workspace_bytes, w.bytes);
  if (scratch_bytes < x.bytes) return fail(ACAN_ERR_WORKSPACE, "acan_attn_bwd_nhwc: scratch %zu < %zu bytes", scratch_bytes, x.bytes);
  if (int e = acan_device_check()) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cg = cta_group();
  const int64_t rows = static_cast<int64_t>(B) * N;
  const __half* fh = feat_h ? static_cast<const __half*>(feat_h) : w.fh;
  if (value == nullptr) value = w.vh;        // the forward converted a bf16 V into the workspace
  const float inv_bn = 1.0f / (static_cast<float>(B) * static_cast<float>(N));
  const size_t pitchP = static_cast<size_t>(N) * 2, batchP = static_cast<size_t>(N) * N * 2;
  const size_t pitchV = static_cast<size_t>(dv) * 2, batchV = static_cast<size_t>(N) * dv * 2;
  const int m_tiles = (N + BM * cg - 1) / (BM * cg);
  CUtensorMap ta, ta2, tb, to;

  if (agg) {
    // dO becomes the fp16 operand of two GEMMs (against the fp16 probabilities and the fp16 values: both operands of a kind::f16 MMA must
    // share one format).  fp16 dO is read in place; bf16 / fp32 dO is packed to fp16 times the power of two s_op that brings max |dO| to
    // 2^-2 (exact for bf16).  delta_i = sum_c dO_ic O_ic: from the values the GEMMs read -- for bf16 that is the bf16 tensor itself (the
    // pack is exact), for fp32 the packed tensor (the pack rounds).
    const void* go = grad_ctx;
    ACAN_CUDA(cudaMemsetAsync(x.amax_bits, 0, 4, st));
    const int delta_grid = static_cast<int>((rows + 7) / 8);
    const int64_t n = rows * dv;
    int64_t pblocks = (n / 8 + 255) / 256;
    if (pblocks > kNumSMs * 16) pblocks = kNumSMs * 16;
    const float4* o4 = reinterpret_cast<const float4*>(ctx32);
    if (go_dtype == DT_F32) {
      amax_f32_kernel<<<static_cast<int>(pblocks), 256, 0, st>>>(static_cast<const float4*>(grad_ctx), n / 4, x.amax_bits);
      if (int e = launch_status("amax_f32_kernel")) return e;
      gscale4_kernel<<<1, 1, 0, st>>>(x.amax_bits, 1, 1, nullptr, inv_bn, x.gscale);
      if (int e = launch_status("gscale4_kernel")) return e;
      to_half_kernel<false><<<static_cast<int>(pblocks), 256, 0, st>>>(grad_ctx, reinterpret_cast<uint4*>(x.doh), n / 8, x.gscale);
      if (int e = launch_status("pack dO (fp32 -> fp16)")) return e;
      go = x.doh;
      delta_nhwc_kernel<false><<<delta_grid, 256, 0, st>>>(static_cast<const uint4*>(go), o4, x.delta, nullptr, rows, dv);
      if (int e = launch_status("delta_nhwc_kernel")) return e;
    } else if (go_dtype == DT_BF16) {
      delta_nhwc_kernel<true><<<delta_grid, 256, 0, st>>>(static_cast<const uint4*>(grad_ctx), o4, x.delta, x.amax_bits, rows, dv);
      if (int e = launch_status("delta_nhwc_kernel")) return e;
      gscale4_kernel<<<1, 1, 0, st>>>(x.amax_bits, 1, 0, nullptr, inv_bn, x.gscale);
      if (int e = launch_status("gscale4_kernel")) return e;
      to_half_kernel<true><<<static_cast<int>(pblocks), 256, 0, st>>>(grad_ctx, reinterpret_cast<uint4*>(x.doh), n / 8, x.gscale);
      if (int e = launch_status("pack dO (bf16 -> fp16)")) return e;
      go = x.doh;
    } else {
      delta_nhwc_kernel<false><<<delta_grid, 256, 0, st>>>(static_cast<const uint4*>(go), o4, x.delta, x.amax_bits, rows, dv);
      if (int e = launch_status("delta_nhwc_kernel")) return e;
      gscale4_kernel<<<1, 1, 0, st>>>(x.amax_bits, 0, 0, nullptr, inv_bn, x.gscale);
      if (int e = launch_status("gscale4_kernel")) return e;
    }
    const bool go_bf16 = false;
    const bool vb = false;
    {  // dV[j, c] = 2^-10 / s_op * sum_i P^[i, j] dO[i, c]: A = P^ read MN-major (j contiguous), B = dO in place MN-major (c contiguous)
      const int bn = choose_bn(dv, B * m_tiles, cg, 64 * cg, 64 * cg);
      const int ob = gv_dtype == DT_F32 ? 4 : 2;
      if (int e = make_tmap(&ta, w.p, 2, N, N, B, pitchP, batchP, 64)) return e;
      if (int e = make_tmap(&tb, go, 2, dv, N, B, pitchV, batchV, 64, go_bf16)) return e;
      if (int e = make_tmap(&to, grad_value, ob, dv, N, B, static_cast<size_t>(dv) * ob, static_cast<size_t>(N) * dv * ob, BM, gv_dtype == DT_BF16)) return e;
      GemmArgs g = gemm_args(B, N, dv, N, bn, cg, umma_idesc_f16(bn, BM * cg, true, true, false, go_bf16));
      g.out_scale = 1.0f / kKeepScale; g.out_gscale = x.gscale + 3;
      g.pdl = 1;
      if (int e = launch_rowmajor<1>(ta, tb, to, g, st, cg, gv_dtype)) return e;
    }
    {  // dP[i, j] = sum_c dO[i, c] V[j, c]: both operands K-major in place -> dS epilogue, fp16(dS * s_ds) out
      const int bn = choose_bn(N, B * m_tiles, cg, 64);
      if (int e = make_tmap(&ta, go, 2, dv, N, B, pitchV, batchV, BM, go_bf16)) return e;
      if (int e = make_tmap(&tb, value, 2, dv, N, B, pitchV, batchV, bn / cg, vb)) return e;
      if (int e = make_tmap(&to, x.ds, 2, N, N, B, pitchP, batchP, BM)) return e;
      GemmArgs g = gemm_args(B, N, N, dv, bn, cg, umma_idesc_f16(bn, BM * cg, false, false, go_bf16, vb));
      g.ds_new = 1; g.ds_delta = x.delta; g.ds_p = w.p; g.ds_lsum = w.lsum; g.gscale = x.gscale;
      g.ds_klT = w.klT; g.ds_gl = have_loss ? grad_loss : nullptr; g.ds_inv_bn = inv_bn;
      g.logbin = have_loss ? w.logbin : nullptr; g.lnz = w.lnz;
      g.pdl = 1;
      if (int e = launch_gemm<MODE_DS, 0, false>(ta, tb, to, g, st, cg)) return e;
    }
  } else {
    gscale4_kernel<<<1, 1, 0, st>>>(nullptr, 0, 0, grad_loss, inv_bn, x.gscale);
    if (int e = launch_status("gscale4_kernel")) return e;
    kl_ds_kernel<<<static_cast<int>(rows), 256, 0, st>>>(w.p, w.logbin, w.lnz, w.klT, x.ds, grad_loss, inv_bn, x.gscale, N, w.lsum);
    if (int e = launch_status("kl_ds_kernel")) return e;
  }
  {  // dF[i, c] = scale / s_ds * sum_j (dS[i, j] + dS[j, i]) F[j, c]: one launch, the k-loop walks dS K-major then MN-major into the same
     // accumulator; B = F in place, MN-major (channels contiguous)
    const int bn = choose_bn(dk, B * m_tiles, cg, 64 * cg, 64 * cg);
    const int ob = gf_dtype == DT_F32 ? 4 : 2;
    if (int e = make_tmap(&ta, x.ds, 2, N, N, B, pitchP, batchP, BM)) return e;
    if (int e = make_tmap(&ta2, x.ds, 2, N, N, B, pitchP, batchP, 64)) return e;
    if (int e = make_tmap(&tb, fh, 2, dk, N, B, static_cast<size_t>(dk) * 2, static_cast<size_t>(N) * dk * 2, 64)) return e;
    if (int e = make_tmap(&to, grad_feat, ob, dk, N, B, static_cast<size_t>(dk) * ob, static_cast<size_t>(N) * dk * ob, BM, gf_dtype == DT_BF16)) return e;
    GemmArgs g = gemm_args(B, N, dk, N, bn, cg, umma_idesc_f16(bn, BM * cg, true, false));
    g.k_half = g.k_steps;
    g.k_steps *= 2;
    g.idesc2 = umma_idesc_f16(bn, BM * cg, true, true);
    g.out_scale = scale; g.out_gscale = x.gscale + 1;
    g.pdl = 1;
    if (int e = launch_rowmajor<2>(ta, tb, to, g, st, cg, gf_dtype, &ta2)) return e;
  }
  return 0;
}
