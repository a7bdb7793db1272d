"""torch.autograd.Function wrappers around the sm_100a kernels.

Forward passes are the hand-written kernels of ``csrc/`` (through the C ABI).  Backward:
  * decode_ord        -- CUDA kernel (``acan_decode_ord_bwd``)
  * attention loss    -- CUDA kernel (gradient w.r.t. sim_map comes out of ``acan_attention_loss``)
  * attention core    -- tcgen05 kernels (``acan_attn_bwd``): P recomputed from the saved row statistics, dV and dP as
                         TMA/UMMA GEMMs, dS (with the attention-loss term) formed in the dP epilogue, dF from two bf16 GEMMs
                         over dS and its transpose (read in place as an MN-major operand).  No CPU path anywhere.
The reference has no custom Function: gradients follow what autograd derives for
sadecoder.py:42-49,80-90 and losses.py:18-33, including Q and K sharing F (dF gets both sides of dS).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


class DecodeOrd(torch.autograd.Function):
    """BaseClassificationModel_.decode_ord (model.py:40-45)."""

    @staticmethod
    def forward(ctx, logits):
        prob = ops.decode_ord(logits)
        ctx.save_for_backward(prob)
        return prob

    @staticmethod
    def backward(ctx, grad_prob):
        (prob,) = ctx.saved_tensors
        return ops.decode_ord_backward(prob, grad_prob.contiguous())


class FusedOrdinalLoss(torch.autograd.Function):
    """OrdinalRegression2d(reduction='sum', ignore_index) o decode_ord on the full-resolution pair logits (losses.py:103-158,
    model.py:40-45): one read of the logits forward, one read + one write backward; the probabilities never exist."""

    @staticmethod
    def forward(ctx, logits, label, ignore_index):
        loss, count, y, lab = ops.ordinal_loss_forward(logits, label, ignore_index)
        ctx.save_for_backward(y, lab, count)
        ctx.ignore_index = ignore_index
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        y, lab, count = ctx.saved_tensors
        return ops.ordinal_loss_backward(y, lab, count, grad_out, ctx.ignore_index), None, None


class FusedCrossEntropyLoss(torch.autograd.Function):
    """CrossEntropy2d(reduction='sum') on the class scores of the CE classifier (losses.py:103-141, 164-196): softmax, the smoothed
    target and the per-class binary cross-entropy in one kernel forward, one backward; no [B,H,W,K] temporaries."""

    @staticmethod
    def forward(ctx, scores, label, ignore_index, eps, prior):
        loss, count, y, lab = ops.ce_loss_forward(scores, label, ignore_index, eps, prior)
        ctx.save_for_backward(y, lab, count)
        ctx.cfg = (ignore_index, eps, prior, scores.dtype)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        y, lab, count = ctx.saved_tensors
        ignore_index, eps, prior, dtype = ctx.cfg
        return ops.ce_loss_backward(y, lab, count, grad_out, ignore_index, eps, prior).to(dtype), None, None, None, None


class FusedTailLoss(torch.autograd.Function):
    """OrdinalRegression2d o decode_ord o Upsample(x8, bilinear, align_corners) on the STRIDE-8 pair logits (model.py:123-125,
    40-45; losses.py:103-158): the full-resolution logits and their gradient never exist (SURVEY §8f rows f-1 x f-2)."""

    @staticmethod
    def forward(ctx, z, label, ignore_index):
        loss, count, z_nhwc, lab = ops.tail_loss_forward(z, label, ignore_index)
        ctx.save_for_backward(z_nhwc, lab, count)
        ctx.ignore_index = ignore_index
        ctx.z_dtype = z.dtype
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        z_nhwc, lab, count = ctx.saved_tensors
        return ops.tail_loss_backward(z_nhwc, lab, count, grad_out, ctx.ignore_index).to(ctx.z_dtype), None, None


class BatchNormAct(torch.autograd.Function):
    """nn.BatchNorm2d in train() (+ residual add, + ReLU) on channels_last half-precision activations: two HBM-bound kernels
    forward, two backward (acan_bn_train_fwd / _bwd) instead of the framework's generic statistics / transform / clamp / add
    passes.  Running statistics are updated in place inside the forward kernel, like F.batch_norm does."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, momentum, eps, relu):
        y, mean, invstd = ops.bn_train_forward(x, weight, bias, running_mean, running_var, momentum, eps, relu, residual)
        ctx.save_for_backward(x, y if relu else None, weight, mean, invstd)
        ctx.has_residual = residual is not None
        ctx.res_dtype = residual.dtype if residual is not None else None
        return y

    @staticmethod
    def backward(ctx, grad_y):
        x, y, weight, mean, invstd = ctx.saved_tensors
        gx, gres, gw, gb = ops.bn_train_backward(x, y, grad_y, weight, mean, invstd, ctx.has_residual and ctx.needs_input_grad[3])
        if gres is not None and gres.dtype != ctx.res_dtype:
            gres = gres.to(ctx.res_dtype)
        return gx, gw.to(weight.dtype), gb.to(weight.dtype), gres, None, None, None, None, None


FUSED_BN = True          # A/B switch: False routes every BatchNorm through the framework's own kernels


class deferred_bn_counters:
    """``with deferred_bn_counters():`` -- the ``num_batches_tracked += 1`` of every fused BatchNorm inside the block is collected and applied
    by ONE multi-tensor add on exit instead of one 2-us kernel per layer (106 of them in a ResNet-101 step: 0.19 ms of a 13.7 ms step)."""
    active = None

    def __enter__(self):
        self._outer = deferred_bn_counters.active
        deferred_bn_counters.active = []
        return self

    def __exit__(self, *exc):
        pending, deferred_bn_counters.active = deferred_bn_counters.active, self._outer
        if pending:
            counts = {}
            for t in pending:                                  # the shared f_key / f_query BatchNorm is bumped twice per forward: one add of 2,
                counts.setdefault(id(t), [t, 0])[1] += 1       # not two racing adds of 1 in the same multi-tensor launch
            with torch.no_grad():
                for k in sorted({c for _, c in counts.values()}):
                    torch._foreach_add_([t for t, c in counts.values() if c == k], k)
        return False


def _bump(counter: torch.Tensor) -> None:
    if deferred_bn_counters.active is not None:
        deferred_bn_counters.active.append(counter)
    else:
        with torch.no_grad():
            counter += 1


def bn_act(bn, x, relu: bool, residual=None, momentum=None):
    """``[relu](bn(x) [+ residual])`` for an ``nn.BatchNorm2d``: the fused kernels when ``bn`` is in train() on channels_last
    half-precision CUDA activations (the autocast training step), the framework's own ops otherwise (fp32 / NCHW / eval)."""
    if FUSED_BN and bn.training and bn.affine and ops.bn_supported(x):
        m = bn.momentum if momentum is None else momentum
        if bn.track_running_stats:
            if m is None:                                     # cumulative moving average (momentum=None)
                m = 1.0 / float(bn.num_batches_tracked + 1)
            _bump(bn.num_batches_tracked)
            rm, rv = bn.running_mean, bn.running_var
        else:
            rm, rv, m = None, None, 0.0
        return BatchNormAct.apply(x, bn.weight.float(), bn.bias.float(), residual, rm, rv, float(m), float(bn.eps), bool(relu))
    if momentum is not None and bn.training:
        keep, bn.momentum = bn.momentum, momentum
        try:
            y = bn(x)
        finally:
            bn.momentum = keep
    else:
        y = bn(x)
    if residual is not None:
        y = y + residual
    return torch.relu(y) if relu else y


class PoolBranch(torch.autograd.Function):
    """SADecoder.image_context without the broadcast (sadecoder.py:97-106) -> g [B,H2]."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, keep_mask):
        if x.dtype in (torch.float16, torch.bfloat16) and x.dim() == 4 and x.shape[1] % 8 == 0 and x.permute(0, 2, 3, 1).is_contiguous():
            mean, hid, g = ops.pool_branch_cl(x, w1, b1, w2, b2, keep_mask)      # channels_last half activations are read in place
        else:
            mean, hid, g = ops.pool_branch(x, w1, b1, w2, b2, keep_mask)
        ctx.save_for_backward(mean, hid, g, w1, w2, keep_mask if keep_mask is not None else torch.empty(0, device=x.device))
        ctx.x_dtype = x.dtype
        ctx.x_shape = x.shape
        return g

    @staticmethod
    def backward(ctx, grad_g):
        mean, hid, g, w1, w2, keep = ctx.saved_tensors
        b, c = mean.shape
        n = 1
        for d in ctx.x_shape[2:]:
            n *= d
        dz2 = grad_g * (g > 0).to(grad_g.dtype)              # [B,H2]
        gw2 = dz2.t() @ hid
        gb2 = dz2.sum(0)
        dz1 = (dz2 @ w2) * (hid > 0).to(grad_g.dtype)        # [B,H1]
        gw1 = dz1.t() @ mean
        gb1 = dz1.sum(0)
        dmean = dz1 @ w1                                     # [B,C]
        if keep.numel():
            dmean = dmean * keep * 2.0
        gx = (dmean / n).to(ctx.x_dtype).view(b, c, *([1] * (len(ctx.x_shape) - 2))).expand(ctx.x_shape)
        return gx, gw1, gb1, gw2, gb2, None


def _softmax_backward(p: torch.Tensor, dp: torch.Tensor) -> torch.Tensor:
    return p * (dp - (dp * p).sum(-1, keepdim=True))


class Attention(torch.autograd.Function):
    """context = aggregate(softmax(F^T F / sqrt(dk)), V) on the channels-last tensor-core path (acan_attn_fwd_nhwc, keep = 1).

    forward(feat [B,dk,h,w], value [B,dv,h,w], holder) -> (context [B,dv,h,w] in value's dtype, token 0-dim)
    ``holder`` receives ``.ws`` (this forward's workspace: operand pack, statistics, probabilities) for lazy sim_map access.
    ``token`` ties the attention loss into this node: ``AttentionLossOnToken`` (built by the drop-in AttentionLoss2d from the lazy
    sim_map) runs the fused loss kernel forward and, in backward, only parks its upstream gradient in the workspace; since the token
    is an output of this node, autograd runs this node's backward afterwards, and ONE acan_attn_bwd_nhwc call differentiates the
    aggregation and the loss together (dV, dP -> dS with the loss term in its epilogue, dF).  Unused outputs arrive as None.
    """

    @staticmethod
    def forward(ctx, feat, value, holder):
        out = ops.attention_forward_train(feat, value)
        ws = out["ws"]
        holder.ws, holder.row_stats = ws, out["row_stats"]
        holder.gen = getattr(holder, "gen", 0) + 1
        ctx.ws, ctx.feat_dtype, ctx.value_dtype = ws, feat.dtype, value.dtype
        ctx.save_for_backward(out["value16"] if out["value16"] is not None else torch.empty(0, device=feat.device), out["ctx32"])
        ctx.set_materialize_grads(False)
        o = out["ctx32"]
        if value.dtype in (torch.float16, torch.bfloat16):
            o = o.to(value.dtype)
        token = torch.zeros((), device=feat.device, dtype=torch.float32)
        return o.permute(0, 3, 1, 2), token

    @staticmethod
    def backward(ctx, grad_ctx, grad_token):
        value16, ctx32 = ctx.saved_tensors
        if value16.numel() == 0:
            value16 = None                       # bf16 V: its fp16 copy lives in the forward's workspace
        ws = ctx.ws
        grad_loss = ws.pending_loss_grad if grad_token is not None else None
        ws.pending_loss_grad = None
        if grad_ctx is None and grad_loss is None:
            return None, None, None
        gf, gv = ops.attention_backward_cl(ws, value16, ctx32, grad_ctx, grad_loss, ctx.feat_dtype, ctx.value_dtype)
        return gf.permute(0, 3, 1, 2), (gv.permute(0, 3, 1, 2) if gv is not None else None), None


class AttentionLossOnToken(torch.autograd.Function):
    """AttentionLoss2d.forward (losses.py:54-62) for the lazy sim_map of an ``Attention`` node: the value comes from the fused
    tensor-core KL over that forward's workspace (which also leaves the target statistics there); the gradient is delivered by
    ``Attention.backward`` (see there)."""

    @staticmethod
    def forward(ctx, token, dis_label, ws):
        ctx.ws = ws
        loss = ops.attention_loss_fused(ws, dis_label)
        ws.has_target = True
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        ctx.ws.pending_loss_grad = grad_out
        return torch.zeros((), device=grad_out.device, dtype=torch.float32), None, None


class AttentionLegacy(torch.autograd.Function):
    """The same through the fp32-NCHW C ABI (acan_attn_fwd / acan_attn_bwd): kept for callers that hold fp32 NCHW tensors and want
    fp32 NCHW results without any layout change, and as the A/B partner of the channels-last path in the tests."""

    @staticmethod
    def forward(ctx, feat, value, dis_label, holder):
        out = ops.attention_forward(feat, value, want_sim_map=False, dis_label=dis_label, ws=getattr(holder, "ws", None))
        holder.ws = out["ws"]
        holder.row_stats = out["row_stats"]
        holder.gen = getattr(holder, "gen", 0) + 1
        ctx.save_for_backward(feat, value, out["ctx"], out["row_stats"], dis_label if dis_label is not None else torch.empty(0, device=feat.device))
        loss = out["loss"] if out["loss"] is not None else torch.zeros((), device=feat.device)
        return out["ctx"], loss

    @staticmethod
    def backward(ctx, grad_ctx, grad_loss):
        feat, value, out, row_stats, dis_label = ctx.saved_tensors
        lab = dis_label if dis_label.numel() else None
        g_feat, g_value = ops.attention_backward(feat, row_stats, value, out, grad_ctx, lab, grad_loss if lab is not None else 0.0)
        return g_feat.view_as(feat), g_value.view_as(value), None, None


class SimMap(torch.autograd.Function):
    """sim_map [B,N,N] materialised on request (forward_att, sadecoder.py:42-49).  Backward through a dense, caller-made
    gradient w.r.t. sim_map is the one place left on cuBLAS (nobody on the reference's path differentiates a materialised
    sim_map except the attention loss, which has its own fused backward)."""

    @staticmethod
    def forward(ctx, feat):
        ctx.save_for_backward(feat)
        sim = ops.attention_forward(feat.float(), None, want_sim_map=True)["sim_map"]
        ctx.sim = sim
        return sim

    @staticmethod
    def backward(ctx, grad_sim):
        (feat,) = ctx.saved_tensors
        b, dk = feat.shape[:2]
        f = feat.reshape(b, dk, -1).float()
        ds = _softmax_backward(ctx.sim, grad_sim)
        return (torch.matmul(f, ds + ds.transpose(1, 2)) * (dk ** -0.5)).view_as(feat)


class AttentionLossMaterialised(torch.autograd.Function):
    """AttentionLoss2d.forward on a real sim_map tensor (losses.py:54-62): one streaming kernel for loss and gradient."""

    @staticmethod
    def forward(ctx, sim_map, dis_label):
        loss, grad = ops.attention_loss(sim_map, dis_label, want_grad=True)
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        (grad,) = ctx.saved_tensors
        return grad * grad_out, None
