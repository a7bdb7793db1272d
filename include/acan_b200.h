/* acan_b200.h -- C ABI of libacan_b200.so: sm_100a kernels for the ACAN hot path
 * (context aggregation module, attention loss, ordinal / cross-entropy depth head).
 *
 * The reference (miraiaroha/ACAN) has no FFI layer: its hot path is nn.Module code over ATen.
 * Each entry point below names the reference function it replaces (path:line under
 * /root/reference/code/models).  The Python drop-in modules in acan_b200/ bind these with
 * ctypes; INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer unless its name ends in _host; tensors are dense,
 *     row-major ("NCHW" with H*W flattened to N / HW); floats are IEEE fp32 unless stated;
 *   - return value: 0 on success, otherwise a cudaError_t value or one of ACAN_ERR_*;
 *     acan_last_error() returns a static/thread-local message for the last non-zero return;
 *   - nothing allocates, frees or synchronises: outputs and workspaces are caller-owned,
 *     work is enqueued on `stream` (a cudaStream_t passed as void*), inputs must stay alive
 *     until the stream reaches the end of the call;
 *   - sizes are validated before any launch; a failing call launches nothing.
 */
#ifndef ACAN_B200_H
#define ACAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACAN_ABI_VERSION 3

#define ACAN_ERR_BAD_ARG     10001   /* null pointer, non-positive size, unsupported shape */
#define ACAN_ERR_WORKSPACE   10002   /* workspace smaller than *_workspace_bytes() */
#define ACAN_ERR_DRIVER      10003   /* cuTensorMapEncodeTiled unavailable / failed */
#define ACAN_ERR_ARCH        10004   /* device is not compute capability 10.x */

/* head `kind` (model.py:47-78) */
#define ACAN_HEAD_OR_SOFT 0   /* soft_ordinal_regression  model.py:69-78 */
#define ACAN_HEAD_OR_HARD 1   /* hard_ordinal_regression  model.py:61-67 */
#define ACAN_HEAD_CE_SOFT 2   /* soft_cross_entropy       model.py:52-59 */
#define ACAN_HEAD_CE_HARD 3   /* hard_cross_entropy       model.py:47-50 */

int         acan_abi_version(void);
const char* acan_last_error(void);
/* 0 if the current device can run the library (sm_100), else ACAN_ERR_ARCH / a cudaError_t. */
int         acan_device_check(void);

/* Kernels launched by this library since it was loaded (bench.py's gpu_launches). */
long long   acan_launch_count(void);

/* Measurement hooks: between acan_profile_begin and acan_profile_end every kernel group is bracketed by a pair of
 * CUDA events recorded on the caller's stream.  acan_profile_end waits for them and returns, per group,
 * the summed device time in ms and the number of brackets.  Groups (ACAN_PROF_*): 0 operand pack, 1 affinity
 * statistics pass, 2 probability pass, 3 aggregation (P.V) pass, 4 head, 5 pooling branch, 6 materialised KL, 7 a whole
 * channels-last attention forward call.  acan_profile_begin brackets groups 0-6; acan_profile_begin_coarse brackets groups 4-7,
 * i.e. the attention call as ONE region: events between its passes would serialise them (they are chained by programmatic
 * dependent launch), so the per-pass and the whole-call views are taken in separate runs. */
#define ACAN_PROF_KINDS 8
int acan_profile_begin(int max_marks);
int acan_profile_begin_coarse(int max_marks);
int acan_profile_end(float* ms_by_kind_host, int* count_by_kind_host);

/* ---------------------------------------------------------------- depth head ---------- */

/* BaseClassificationModel_.decode_ord (model.py:40-45):
 *   prob[b,k,p] = exp(y[b,2k+1,p]) / (exp(y[b,2k,p]) + exp(y[b,2k+1,p])),  unstabilised like the reference.
 *   logits [B,2K,HW] -> prob [B,K,HW]. */
int acan_decode_ord_fwd(const float* logits, float* prob, int B, int K, int64_t HW, void* stream);

/* autograd of decode_ord: grad_logits[b,2k+1,p] = g*p*(1-p), grad_logits[b,2k,p] = -g*p*(1-p). */
int acan_decode_ord_bwd(const float* prob, const float* grad_prob, float* grad_logits,
                        int B, int K, int64_t HW, void* stream);

/* BaseClassificationModel_.inference and the four functions it dispatches to (model.py:47-99).
 *   in        OR kinds: prob [B,K,HW] (from_logits = 0) or pair logits [B,2K,HW] (from_logits = 1,
 *             decode_ord fused, the [B,K,HW] tensor never exists);  CE kinds: scores [B,K,HW].
 *   depth     [B,1,HW] metres.
 *   index     optional (may be NULL) int32 [B,1,HW]: the hard bin index (count of p > 0.5 for OR_HARD,
 *             first arg-max for CE_HARD, floor(sum p) for OR_SOFT, arg-max for CE_SOFT). */
int acan_head_fwd(const float* in, float* depth, int32_t* index, int B, int K, int64_t HW,
                  double d_min, double d_max, int kind, int from_logits, void* stream);

/* Fused classifier tail: the x8 bilinear upsampling (align_corners=True) of make_classifier
 * (model.py:123-125), decode_ord (model.py:40-45) and the inference head (model.py:47-99) in one kernel that reads the
 * STRIDE-8 logits and never materialises the [B,2K,H,W] tensor.
 *   z_nhwc  [B,h,w,C] fp32, C = 2K (OR kinds) or K (CE kinds): output of classifier.conv, channels last
 *   depth   [B,1,8h,8w];  prob: NULL, or [B,K,8h,8w] receiving decode_ord(upsample(z)) (the forward dict's 'y');
 *   index   NULL or int32 [B,1,8h,8w]. */
int acan_tail_fwd(const float* z_nhwc, float* depth, float* prob, int32_t* index, int B, int h, int w, int K,
                  double d_min, double d_max, int kind, void* stream);

/* OrdinalRegression2d in the reference's live configuration (losses.py:103-158: reduction='sum', no class weights,
 * ignore_index) fused with decode_ord (model.py:40-45) -- SURVEY.md §8f row f-2.
 *   logits [B,2K,HW] fp32 (full-resolution pair logits), label [B,HW] int64 bin indices, HW % 4 == 0
 *   loss_out / count_out: 1 float each (mean over valid pixels; number of valid pixels, kept for backward)
 *   ws: acan_ordinal_loss_ws_floats() floats.   backward: grad_logits [B,2K,HW] = grad_out * dloss/dlogits. */
int acan_ordinal_loss_ws_floats(void);
int acan_ordinal_loss_fwd(const float* logits, const int64_t* label, float* loss_out, float* count_out, float* ws,
                          int B, int K, int64_t HW, int ignore_index, void* stream);
int acan_ordinal_loss_bwd(const float* logits, const int64_t* label, const float* count, const float* grad_out, float* grad_logits,
                          int B, int K, int64_t HW, int ignore_index, void* stream);

/* CrossEntropy2d in the reference's live configuration (losses.py:103-141, 164-196: reduction='sum', no class weights) for the CE
 * classifier -- SURVEY.md §8f row f-2: a per-class binary cross-entropy on softmax(scores) against the one-hot (eps = 0) or smoothed
 * (eps > 0; prior_kind 0 = uniform 1/(K-1), 1 = gaussian sigma = K/10) target, log arguments clamped to [1e-8, 1e8], mean over the
 * pixels whose label != ignore_index.   scores [B,K,HW] fp32, label [B,HW] int64;  ws: acan_ordinal_loss_ws_floats() floats;
 * backward: grad_scores [B,K,HW] = grad_out * dloss/dscores (0 at ignored pixels). */
int acan_ce_loss_fwd(const float* scores, const int64_t* label, float* loss_out, float* count_out, float* ws,
                     int B, int K, int64_t HW, int ignore_index, float eps, int prior_kind, void* stream);
int acan_ce_loss_bwd(const float* scores, const int64_t* label, const float* count, const float* grad_out, float* grad_scores,
                     int B, int K, int64_t HW, int ignore_index, float eps, int prior_kind, void* stream);

/* Training-mode classifier tail fused with the ordinal loss (SURVEY.md §8f rows f-1 x f-2): replaces
 * nn.Upsample(x8, bilinear, align_corners=True) (model.py:123-125, 228-230) -> decode_ord (model.py:40-45) ->
 * OrdinalRegression2d (losses.py:103-158) and their three backward passes.  Works from the STRIDE-8 pair logits; the
 * [B,2K,8h,8w] tensor and its gradient never exist.
 *   z_nhwc  [B,h,w,2K] fp32 (output of classifier.conv, channels last), label [B,8h,8w] int64 bin indices
 *   loss_out / count_out: 1 float each (mean over pixels whose label != ignore_index; their number, kept for backward)
 *   ws: acan_tail_loss_ws_floats(B,h,w) floats.   backward: grad_z_nhwc [B,h,w,2K] = grad_out * dloss/dz (deterministic). */
size_t acan_tail_loss_ws_floats(int B, int h, int w);
int acan_tail_loss_fwd(const float* z_nhwc, const int64_t* label, float* loss_out, float* count_out, float* ws,
                       int B, int h, int w, int K, int ignore_index, void* stream);
int acan_tail_loss_bwd(const float* z_nhwc, const int64_t* label, const float* count, const float* grad_out, float* grad_z_nhwc,
                       int B, int h, int w, int K, int ignore_index, void* stream);

/* Epilogue of the merge convolution (sadecoder.py:107-111): y <- ReLU(y + bias[b, :]) in place, y [B,N,C] channels-last fp16
 * (dtype 1) / bf16 (dtype 2), bias [B,C] fp32 = W[:, dv:] g + b, the spatially constant share of conv1x1(cat(context, g)). */
int acan_bias_relu_nhwc(void* y, int dtype, const float* bias, int B, int N, int C, void* stream);

/* nn.MaxPool2d(kernel_size=3, stride=2, padding=1) of the stem (model.py:170, 218) on a channels-last fp16 (dtype 1) / bf16
 * (dtype 2) tensor, inference only (no indices): x [B,H,W,C] -> y [B,(H-1)/2+1,(W-1)/2+1,C], C % 8 == 0. */
int acan_maxpool3x3s2_nhwc(const void* x, void* y, int dtype, int B, int H, int W, int C, void* stream);

/* Training-mode BatchNorm2d over channels-last half-precision activations fused with the ReLU (and residual add) behind it:
 * the Q/K projection's BN + ReLU (sadecoder.py:25-30) and, same arithmetic, every
 * conv -> BN -> ReLU of the backbone (model.py:128-160).  Semantics of nn.BatchNorm2d in train(): batch statistics (biased
 * variance) normalise, running statistics take the unbiased variance with `momentum`.
 *   x, y, residual, grad_*: [M, C] (M = B*H*W, channels contiguous), dtype 0 = fp16 / 1 = bf16, C % 8 == 0, 16-byte aligned
 *   y = [relu]( (x - mean) * invstd * weight + bias [+ residual] );   save_mean / save_invstd [C] fp32 are kept for backward
 *   running_mean / running_var: NULL (track_running_stats = False) or [C] fp32, updated in place
 *   ws: acan_bn_ws_floats(C) floats, ZERO-INITIALISED ONCE by the caller (it holds self-resetting arrival counters and the generation
 *       words of the slab barrier, which keep whatever value the last call left) and not shared between streams.
 *   backward: y = the forward output when a ReLU was fused (its sign is the mask), else NULL; grad_residual: NULL, or receives the
 *       masked upstream gradient (what flows into the residual branch); grad_weight / grad_bias [C] fp32. */
size_t acan_bn_ws_floats(int C);
/* SMs the single-launch (cooperative) BN grids may cover, 32..148 (default 148); returns the previous value, n_sms <= 0 only queries.
 * A cooperative grid starts only when all its blocks fit at once: a training step that overlaps an NCCL collective with its backward
 * pass leaves NCCL's SMs out (acan_b200.training.GraphedTrainStep(overlap=True)).  Process-wide; read at launch (so at graph capture). */
int acan_bn_set_sm_limit(int n_sms);
int acan_bn_train_fwd(const void* x, int dtype, const float* weight, const float* bias, const void* residual, void* y,
                      float* save_mean, float* save_invstd, float* running_mean, float* running_var,
                      float momentum, float eps, int relu, float* ws, int64_t M, int C, void* stream);
int acan_bn_train_bwd(const void* x, const void* y, const void* grad_y, int dtype, const float* weight,
                      const float* save_mean, const float* save_invstd, void* grad_x, void* grad_residual,
                      float* grad_weight, float* grad_bias, float* ws, int64_t M, int C, void* stream);

/* The eight depth-error measures of the evaluation loop (utils/eval_utils.py:16-48 compute_errors) in one
 * pass: gt, pred [n] fp32 (n = batch_size * H * W), pixels with gt <= 0 ignored; out8 (device) receives, in the reference's
 * measure_list order, a1, a2, a3, rmse, rmse_log, log10 (a natural log in the reference, kept), abs_rel, sq_rel -- each the mean over
 * the valid pixels of the whole batch times batch_size, as the reference returns them.  ws: acan_depth_metrics_ws_floats() floats. */
int acan_depth_metrics_ws_floats(void);
int acan_depth_metrics(const float* gt, const float* pred, float* out8, float* ws, int64_t n, int batch_size, void* stream);

/* continuous2discrete (model.py:19-23, torch-0.4 semantics): metres -> bin index as float, 0 outside (d_min,d_max). */
int acan_continuous2discrete(const float* depth, float* bins, int64_t n, double d_min, double d_max, int K, void* stream);

/* ---------------------------------------------------------------- pooling branch ------ */

/* SADecoder.image_context (sadecoder.py:97-106): global mean over the N positions, optional
 * Dropout2d keep-mask, Linear(C->H1)+ReLU, Linear(H1->H2)+ReLU.  The trailing bilinear "upsample"
 * of the 1x1 map is a broadcast and is never materialised.
 *   x [B,C,N] fp32;  keep_mask NULL or [B,C] in {0,1} (kept channels are scaled by 2);
 *   w1 [H1,C], b1 [H1], w2 [H2,H1], b2 [H2];
 *   mean_out [B,C] (post-mask, saved for backward), hidden_out [B,H1] (post-ReLU), g_out [B,H2]. */
int acan_pool_branch_fwd(const float* x, const float* keep_mask,
                         const float* w1, const float* b1, const float* w2, const float* b2,
                         float* mean_out, float* hidden_out, float* g_out,
                         int B, int C, int N, int H1, int H2, void* stream);

/* The same for a channels-last half-precision feature map: x [B,N,C] fp16 (x_dtype 1) or bf16 (x_dtype 2), C % 8 == 0.
 * scratch: acan_pool_nhwc_scratch_floats(B, C, N) floats (slab partial sums, reduced in a fixed order). */
size_t acan_pool_nhwc_scratch_floats(int B, int C, int N);
int acan_pool_branch_fwd_nhwc(const void* x, int x_dtype, const float* keep_mask,
                              const float* w1, const float* b1, const float* w2, const float* b2,
                              float* scratch, float* mean_out, float* hidden_out, float* g_out,
                              int B, int C, int N, int H1, int H2, void* stream);

/* ---------------------------------------------------------------- attention ----------- */

/* Bytes of scratch the attention entry points need for (B, N, dk, dv).  The workspace must be 1024-byte
 * aligned.  After acan_attn_fwd it holds the fp16 operand pack of F and the row statistics, which
 * acan_attn_rows and acan_attention_loss_fused reuse (pass the same B, N, dk, dv). */
size_t acan_attn_workspace_bytes(int B, int N, int dk, int dv);

/* PixelAttentionBlock_.forward_att + SelfAttentionBlock_.forward (sadecoder.py:42-49, 80-90), Q = K = F:
 *   S = F^T F * scale;  P = softmax_rows(S);  O[b,c,i] = sum_j P[b,i,j] V[b,c,j].
 *   feat   F [B,dk,N] fp32 (f_key output, NCHW)
 *   value  V [B,dv,N] fp32 (f_value output, NCHW); NULL (with ctx NULL) computes the affinity only
 *   ctx    O [B,dv,N] fp32 (NCHW, what .permute(0,2,1).contiguous().view() yields in the reference)
 *   sim_map    NULL, or [B,N,N] fp32: the full affinity, materialised only on request
 *   row_stats  [B,N,2] fp32: (m_i, l_i) with P_ij = exp2(S_ij*scale*log2e - m_i) / l_i  (kept for loss / backward)
 * tcgen05 path: operands are rounded to fp16 (11-bit significand, the rounding of TF32), accumulation fp32.
 * Requires dk % 64 == 0, dv % 128 == 0, N % 8 == 0. */
int acan_attn_fwd(const float* feat, const float* value, float* ctx, float* sim_map, float* row_stats,
                  void* workspace, size_t workspace_bytes,
                  int B, int N, int dk, int dv, float scale, void* stream);

/* Same, with the attention loss (losses.py:54-62) fused into the probability pass: sim_map is compared
 * tile by tile, on chip, against the target affinity generated on the fly from dis_label [B,1,H,W]
 * (bin indices, model.py:260,270; (H/8)*(W/8) == N).  loss_out: 1 float (device). */
int acan_attn_fwd_loss(const float* feat, const float* value, float* ctx, float* sim_map, float* row_stats,
                       const float* dis_label, int H, int W, float* loss_out,
                       void* workspace, size_t workspace_bytes,
                       int B, int N, int dk, int dv, float scale, void* stream);

/* Backward of acan_attn_fwd / acan_attn_fwd_loss on the tensor cores (the reference relies on autograd through
 * sadecoder.py:42-49,85-87 and losses.py:18-33; Q and K share F, so dF collects both sides of dS):
 *   dV = dO P      dP = dO^T V      dS = P o (dP - delta_i) - g/(B N) (W o [clamp window] - P T_i)      dF = scale F (dS + dS^T)
 *   feat, value, ctx (the forward output), row_stats: as given to / returned by the forward call (fp32 NCHW)
 *   grad_ctx   [B,dv,N] fp32, or NULL when only the attention loss is differentiated (value / ctx / grad_value unused)
 *   dis_label  NULL, or the forward's label: adds *grad_loss * d(attention loss)/dF;  grad_loss is a DEVICE scalar (the upstream
 *              gradient of the loss as autograd hands it over): the call never synchronises with the host and can be captured
 *              in a CUDA graph
 *   grad_feat  [B,dk,N] fp32, grad_value [B,dv,N] fp32.   workspace: acan_attn_bwd_workspace_bytes(), 1024-byte aligned.
 * P is recomputed from row_stats; every tensor-core operand is fp16 with fp32 accumulation, dO and dS pre-multiplied by a
 * power of two chosen on the device from max|dO| (so the loss scale cannot push them out of the fp16 range). */
size_t acan_attn_bwd_workspace_bytes(int B, int N, int dk, int dv);
int acan_attn_bwd(const float* feat, const float* value, const float* ctx, const float* row_stats, const float* grad_ctx,
                  const float* dis_label, int H, int W, const float* grad_loss,
                  float* grad_feat, float* grad_value, void* workspace, size_t workspace_bytes,
                  int B, int N, int dk, int dv, float scale, void* stream);

/* Channels-last half-precision form of the two calls above, for pipelines whose convolutions already emit fp16 in
 * channels_last memory: nothing is packed or transposed.
 *   feat_h  [B,N,dk] fp16 (f_key output, NHWC)   -- is the K-major operand of S as it stands
 *   value_h [B,N,dv] fp16 (f_value output, NHWC) -- read in place as an MN-major tensor-core operand; NULL = affinity only
 *   ctx_h   [B,N,dv] fp16 (NHWC)
 *   dis_label / loss_out: NULL, or as in acan_attn_fwd_loss.   Requires dv % 128 == 0. */
int acan_attn_fwd_nhwc_f16(const void* feat_h, const void* value_h, void* ctx_h, float* sim_map, float* row_stats,
                           const float* dis_label, int H, int W, float* loss_out,
                           void* workspace, size_t workspace_bytes,
                           int B, int N, int dk, int dv, float scale, void* stream);

/* General channels-last form (the training step's entry; acan_attn_fwd_nhwc_f16 is this call with fp16 everywhere and keep = 0).
 * Element types (ACAN_DT_*): feat fp16 (read in place), bf16 or fp32 (one elementwise pass into the workspace's fp16 pack);
 * value fp16, read in place, or bf16, converted exactly into the workspace by one pass (an fp16 x bf16 kind::f16 MMA traps as an
 * illegal instruction on sm_100a although the descriptor has one format field per operand, and the probabilities need fp16's 11 bits);
 * ctx fp16, bf16 or fp32.
 *   keep = 0  inference: statistics-free probability pass when neither sim_map nor the loss is requested (diagonal, early-exit
 *             statistics, probability and aggregation launches chained by programmatic dependent launch).
 *   keep = 2  inference with the probability and aggregation tiles interleaved in ONE persistent launch (dv % 256 == 0, 16-bit ctx):
 *             bit-identical, measured slower than keep = 0 on B200 (both passes are shared-memory-bandwidth bound); kept as the
 *             documented experiment it is.
 *   keep = 1  training (autograd through sadecoder.py:42-49, 80-90): exact two-pass row statistics; the normalised probabilities stay
 *             in the workspace as fp16(P * 2^10) together with the row sums of those rounded values, which is the divisor used by
 *             the aggregation here and by acan_attn_bwd_nhwc -- the rows the tensor cores read then sum to exactly 1, which is what
 *             keeps the softmax-backward cancellation (sum_j dS_ij = 0) intact in 16-bit operands.  Ask for ctx in fp32 when a
 *             backward pass follows: delta_i = sum_c dO_ic O_ic needs the unrounded aggregation output. */
#define ACAN_DT_F32  0
#define ACAN_DT_F16  1
#define ACAN_DT_BF16 2
int acan_attn_fwd_nhwc(const void* feat, int feat_dtype, const void* value, int value_dtype, void* ctx, int ctx_dtype,
                       float* sim_map, float* row_stats, const float* dis_label, int H, int W, float* loss_out,
                       void* workspace, size_t workspace_bytes, int B, int N, int dk, int dv, float scale, int keep, void* stream);

/* Backward of acan_attn_fwd_nhwc(keep = 1), channels-last, four tensor-core launches and no recomputation of P:
 *   dV = P^T dO      dP = dO V^T      dS = p o (dP - delta_i) - g/(B N) (W o [clamp window] - p T_i)      dF = scale (dS + dS^T) F
 *   workspace   the FORWARD's workspace, untouched since: probabilities, their row sums, row statistics; with have_loss also what
 *               acan_attention_loss_fused (or a forward with a label) left there: ln(bin), ln Z, the in-window target mass T_i
 *   scratch     acan_attn_bwd_nhwc_scratch_bytes(), 1024-byte aligned (dS, an fp16 pack of an fp32 grad_ctx, delta, scales)
 *   feat_h      [B,N,dk] fp16 as given to the forward (NULL: the workspace pack, when the forward converted bf16 / fp32 features)
 *   value       [B,N,dv] fp16 as given to the forward (value_dtype ACAN_DT_F16), or NULL when the forward was given bf16 values (their fp16
 *               copy is in the workspace); ctx32 [B,N,dv] fp32 the forward's output
 *   grad_ctx    [B,N,dv] fp16 (read in place as an MMA operand), or bf16 / fp32 (packed to fp16 times a power of two chosen on the
 *               device from max |dO|: exact for bf16); NULL when only the attention loss is differentiated
 *   have_loss   adds *grad_loss (DEVICE scalar) * d(attention loss)/dF
 *   grad_feat   [B,N,dk], grad_value [B,N,dv]; fp32 / fp16 / bf16 each.
 * delta_i is computed from the very operand the dP GEMM reads and dS is formed with p = P^/lsum, so sum_j dS_ij = 0 holds for the
 * values the tensor cores see; measured gradient error vs fp64 autograd <= 5e-4 (max-norm relative) at the reference's channel counts. */
size_t acan_attn_bwd_nhwc_scratch_bytes(int B, int N, int dk, int dv);
int acan_attn_bwd_nhwc(void* workspace, size_t workspace_bytes, void* scratch, size_t scratch_bytes,
                       const void* feat_h, const void* value, int value_dtype, const float* ctx32,
                       const void* grad_ctx, int go_dtype, int have_loss, const float* grad_loss,
                       void* grad_feat, int gf_dtype, void* grad_value, int gv_dtype,
                       int B, int N, int dk, int dv, float scale, void* stream);

/* Selected rows of sim_map (vis_utils.py:107-121 reads rows N/4, N/2, 3N/4 only):
 *   rows_out [B, n_rows, N] fp32 for row indices row_idx[0..n_rows) (int32, device).
 * Needs the workspace left by a preceding forward with the same (B, N, dk, dv); feat_pack = NULL after
 * acan_attn_fwd (the pack lives in the workspace), = the feat_h pointer after acan_attn_fwd_nhwc_f16. */
int acan_attn_rows(const void* workspace, const void* feat_pack, const int* row_idx, int n_rows, float* rows_out,
                   int B, int N, int dk, int dv, float scale, void* stream);

/* ---------------------------------------------------------------- attention loss ------ */

/* AttentionLoss2d.forward (losses.py:54-62) on a materialised sim_map:
 *   bins[b,n] = dis_label[b,0,8*(n / w), 8*(n % w)]   (nearest downsample, losses.py:50)
 *   W = gt affinity (losses.py:40-46);  loss = mean_{b,i} sum_j W log(clamp(W/Q, 1e-8, 1e8)).
 *   sim_map [B,N,N] fp32, dis_label [B,1,H,W] fp32 bin indices (model.py:260,270), h = H/8, w = W/8, N = h*w.
 *   loss_out: 1 float (device).  ws: [2*B*N] floats of scratch (row losses, stride-8 bins).
 *   grad_sim: NULL, or [B,N,N] receiving grad_scale * dLoss/dsim_map = -grad_scale * W/Q / (B*N) inside the
 *             clamp window, 0 outside it (where the reference yields 0, or NaN when Q == 0 exactly). */
int acan_attention_loss(const float* sim_map, const float* dis_label, float* loss_out, float* ws,
                        float* grad_sim, float grad_scale, int B, int H, int W, void* stream);

/* AttentionLoss2d.get_gt_sim_map (losses.py:48-52): gt [B,N,N] fp32 from dis_label [B,1,H,W]; bins_ws: [B*N] floats. */
int acan_gt_sim_map(const float* dis_label, float* gt, float* bins_ws, int B, int H, int W, void* stream);

/* Fused, streamed form on its own: the loss straight from the fp16 operand pack and row statistics an
 * acan_attn_fwd call left in `workspace` -- sim_map is recomputed tile by tile on the tensor cores and
 * never touches HBM.  Same value as acan_attention_loss(sim_map) up to fp16-operand rounding of S.
 * feat_pack: as for acan_attn_rows.  Also leaves ln(bin), ln Z and the in-window target mass T_i in the workspace
 * (what acan_attn_bwd_nhwc(have_loss = 1) reads). */
int acan_attention_loss_fused(void* workspace, const void* feat_pack, const float* dis_label, float* loss_out,
                              int B, int H, int W, int dk, int dv, float scale, void* stream);

/* ---------------------------------------------------------------- optimizer step ------ */

/* torch.optim.SGD (momentum, weight decay; nesterov = False, dampening = 0: what creator/optimizer.py:8-10 builds) over flat fp32
 * buffers in ONE pass: g' = g + wd p;  m = mu m + g';  p = p - lr m;  optionally the bf16 shadow of p is written in the same pass.
 * The training driver (acan_b200/training.py) binds every parameter, gradient and momentum buffer of the network to slices of such
 * buffers, in net.parameters() order; per-group hyper-parameters come from a DEVICE table, so a learning-rate schedule needs no re-capture.
 *   seg_end[n_seg] (int64) / seg_group[n_seg] (int32): end offset (exclusive) and parameter-group index of each parameter's slice
 *   chunk_group[ceil(n / acan_sgd_flat_chunk())] (int8): the group of a chunk of elements when it lies in one group, -1 when it spans several
 *   hyper[n_groups][4] fp32: (lr, momentum, weight_decay, unused)
 *   first_chunk, n_chunks: the range of chunks to update (the multi-GPU driver updates a range as soon as its all-reduce has landed) */
int acan_sgd_flat_chunk(void);
int acan_sgd_flat(float* param, const float* grad, float* momentum_buf, void* shadow_bf16, const signed char* chunk_group,
                  const long long* seg_end, const int* seg_group, int n_seg, const float* hyper,
                  long long first_chunk, long long n_chunks, long long n, void* stream);

/* n_tensors dense 16-bit tensors -> fp32, one launch per 96 tensors: the step that moves the weight gradients autocast leaves in 16 bit
 * (one tensor per layer) into their slices of the flat fp32 gradient buffer (acan_b200/training.py; replaces one framework cast kernel
 * per layer).  src / dst / count are HOST arrays of device pointers and element counts (read during the call; the pointers travel in the
 * kernels' parameter space, so a captured CUDA graph replays with the captured addresses).  src_dtype: ACAN_DT_F16 or ACAN_DT_BF16. */
int acan_cast_to_f32_multi(const void* const* src, float* const* dst, const long long* count, int n_tensors, int src_dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACAN_B200_H */
