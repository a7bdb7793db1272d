"""Drop-in nn.Modules for the reference's context aggregation module and attention loss.

Same class names, constructor arguments, ``state_dict`` keys / shapes, parameter names and forward
signatures as ``code/models/sadecoder.py`` and ``code/models/losses.py:35-62`` of the reference, so
checkpoints load both ways and ``create_params`` (creator/params.py:14-31) builds the same optimizer groups.
What differs is what runs: the affinity / softmax / aggregation, the pooling branch and the loss are
the sm_100a kernels of ``csrc/`` (through ``libacan_b200.so``); the convolutions stay cuDNN calls.
CUDA tensors only -- there is no CPU path.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functions as fn
from . import ops


class LazySimMap:
    """Stand-in for ``sim_map [B,N,N]`` that costs nothing until somebody looks at it.

    The reference always returns sim_map (sadecoder.py:122, model.py:233) and ``LossFunc.forward`` always
    *reads* the dict entry (model.py:258), but only two consumers ever touch the values: the attention loss
    when beta != 0 (model.py:270) and the TensorBoard figure, which takes rows N/4, N/2, 3N/4
    (vis_utils.py:107-121).  Both are served without materialising N x N in HBM:
      * ``AttentionLoss2d`` (below) recognises this object and runs the fused tensor-core KL;
      * ``lazy.cpu()[:b][:, points]`` runs the selected-rows kernel;
    anything else (torch functions, arithmetic, ``.materialize()``) produces the real fp32 tensor through
    ``functions.SimMap`` -- with autograd back to F.
    """

    def __init__(self, feat: torch.Tensor, holder, batch_slice: Optional[slice] = None, to_cpu: bool = False, gen: Optional[int] = None,
                 token: Optional[torch.Tensor] = None, ws=None):
        self._feat = feat
        self._holder = holder
        self._gen = holder.gen if gen is None else gen
        self._slice = batch_slice
        self._cpu = to_cpu
        self._dense = None
        self._own_ws = ws          # training forwards own their workspace (it outlives later forwards of the module)
        self._token = token        # output of the fn.Attention node this sim_map belongs to (ties the attention loss into its backward)

    def _ws(self):
        """operand pack + row statistics for THIS forward: the module's workspace if it still belongs to it,
        otherwise (the module ran again since) a private re-pack."""
        if self._own_ws is not None:
            return self._own_ws
        if self._holder.gen == self._gen and self._holder.ws is not None:
            return self._holder.ws
        if self._own_ws is None:
            self._own_ws = ops.attention_forward(self._feat.detach(), None)["ws"]
        return self._own_ws

    # -- tensor-like surface -------------------------------------------------------------
    @property
    def shape(self):
        b = self._feat.shape[0]
        if self._slice is not None:
            b = len(range(*self._slice.indices(b)))
        n = self._feat.shape[2] * self._feat.shape[3]
        return torch.Size((b, n, n))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 3

    @property
    def device(self):
        return torch.device("cpu") if self._cpu else self._feat.device

    @property
    def dtype(self):
        return torch.float32

    @property
    def requires_grad(self):
        return self._feat.requires_grad

    def cpu(self):
        return LazySimMap(self._feat, self._holder, self._slice, True, self._gen, self._token, self._own_ws)

    def cuda(self, *a, **k):
        return LazySimMap(self._feat, self._holder, self._slice, False, self._gen, self._token, self._own_ws)

    def detach(self):
        return LazySimMap(self._feat.detach(), self._holder, self._slice, self._cpu, self._gen, None, self._own_ws)

    def materialize(self) -> torch.Tensor:
        if self._dense is None:
            sim = fn.SimMap.apply(self._feat)
            if self._slice is not None:
                sim = sim[self._slice]
            self._dense = sim.cpu() if self._cpu else sim
        return self._dense

    def rows(self, points) -> torch.Tensor:
        """sim_map[:, points] -> [B, len(points), N] via the selected-rows kernel (no N x N tensor)."""
        idx = torch.as_tensor(points, dtype=torch.int64).reshape(-1)
        n = int(self.shape[1])
        if idx.numel() == 0 or int(idx.min()) < -n or int(idx.max()) >= n:
            raise IndexError(f"sim_map row index out of range for N={n}")
        idx = idx % n                                       # python-style negative indices
        out = ops.attention_rows(self._ws(), idx)
        if self._slice is not None:
            out = out[self._slice]
        return out.cpu() if self._cpu else out

    def __getitem__(self, key):
        if isinstance(key, slice):
            if self._slice is None:
                return LazySimMap(self._feat, self._holder, key, self._cpu, self._gen, self._token, self._own_ws)
        elif isinstance(key, tuple) and len(key) == 2 and key[0] == slice(None) and not isinstance(key[1], slice):
            k1 = key[1]
            if isinstance(k1, int):
                return self.rows([k1])[:, 0]
            if isinstance(k1, (list, tuple)) or (torch.is_tensor(k1) and k1.dtype in (torch.int32, torch.int64)):
                return self.rows(k1)
        return self.materialize()[key]

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f"LazySimMap(shape={tuple(self.shape)}, device={self.device}, materialised={self._dense is not None})"

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def conv(a):
            if isinstance(a, LazySimMap):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(conv(x) for x in a)
            return a
        return func(*conv(args), **{k: conv(v) for k, v in (kwargs or {}).items()})

    def __getattr__(self, name):            # any other Tensor method: fall through to the dense tensor
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)


class LazyProb:
    """Stand-in for the forward dict's ``'y'`` ([B,K,H,W] bin probabilities / scores) in eval mode.

    The reference upsamples the stride-8 logits x8 and decodes them at full resolution (model.py:123-125, 228-230) only for
    ``inference(y)`` to reduce them to one depth per pixel (depthest_trainer.py:184-185).  This handle keeps the stride-8
    logits; ``ResNet.inference`` recognises it and runs the fused tail kernel (``acan_tail_fwd``: upsample + decode_ord +
    head, the full-resolution tensor never exists).  Any other use materialises the real tensor with the same kernel."""

    def __init__(self, z: torch.Tensor, classifier: str, d_min: float, d_max: float, n_c: int):
        self._z, self._cls, self._dmin, self._dmax, self._k = z, classifier, d_min, d_max, n_c
        self._dense = None

    @property
    def shape(self):
        b, _, h, w = self._z.shape
        return torch.Size((b, self._k, 8 * h, 8 * w))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 4

    @property
    def device(self):
        return self._z.device

    @property
    def dtype(self):
        return torch.float32

    def depth(self, mode: str, want_index: bool = False):
        return ops.tail_inference(self._z, self._cls, mode, self._dmin, self._dmax, self._k, want_index=want_index)

    def materialize(self) -> torch.Tensor:
        if self._dense is None:
            _, self._dense = ops.tail_inference(self._z, self._cls, "soft", self._dmin, self._dmax, self._k, want_prob=True)
        return self._dense

    def __getitem__(self, key):
        return self.materialize()[key]

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f"LazyProb(shape={tuple(self.shape)}, materialised={self._dense is not None})"

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def conv(a):
            if isinstance(a, LazyProb):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(conv(x) for x in a)
            return a
        return func(*conv(args), **{k: conv(v) for k, v in (kwargs or {}).items()})

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)


class LazyLogitsProb:
    """Training-mode stand-in for ``'y'`` = decode_ord(upsampled logits) ([B,K,H,W]).  Holds the full-resolution pair logits
    (with their autograd history); the drop-in ``OrdinalRegression2d`` consumes them through the fused loss kernel and
    ``inference`` through the from-logits head, so the probability tensor and its gradient are never materialised.  Any
    other consumer (e.g. the reference's own loss classes) gets ``decode_ord(logits)`` with the usual autograd."""

    def __init__(self, logits: Optional[torch.Tensor], n_c: int, z: Optional[torch.Tensor] = None, upsample=None):
        # either the full-resolution pair logits, or the stride-8 logits `z` plus the module that upsamples them: the drop-in
        # loss / inference then work from z directly (fused tail kernels) and the full-resolution tensor is built on demand only
        self._logits, self._k, self._dense, self._z, self._up = logits, n_c, None, z, upsample

    @property
    def z(self):
        """stride-8 pair logits [B,2K,h,w] (None when the handle was built from full-resolution logits)."""
        return self._z

    @property
    def logits(self):
        if self._logits is None:
            self._logits = self._up(self._z).float()
        return self._logits

    @property
    def shape(self):
        if self._logits is None:
            b, _, h, w = self._z.shape
            return torch.Size((b, self._k, 8 * h, 8 * w))
        b, _, h, w = self._logits.shape
        return torch.Size((b, self._k, h, w))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 4

    @property
    def device(self):
        return (self._z if self._logits is None else self._logits).device

    @property
    def dtype(self):
        return torch.float32

    @property
    def requires_grad(self):
        return (self._z if self._logits is None else self._logits).requires_grad

    def materialize(self) -> torch.Tensor:
        if self._dense is None:
            self._dense = fn.DecodeOrd.apply(self.logits.float())
        return self._dense

    def __getitem__(self, key):
        return self.materialize()[key]

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f"LazyLogitsProb(shape={tuple(self.shape)}, materialised={self._dense is not None})"

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def conv(a):
            if isinstance(a, LazyLogitsProb):
                return a.materialize()
            if isinstance(a, (list, tuple)):
                return type(a)(conv(x) for x in a)
            return a
        return func(*conv(args), **{k: conv(v) for k, v in (kwargs or {}).items()})

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)


class _Holder:
    """per-module slot for the attention workspace of the latest forward (operand pack, row statistics).  Inference workspaces are
    pooled by shape and never dropped while the module lives: a captured CUDA graph (GraphedPredictor) holds their raw addresses, and an
    eager forward with another batch size in between (the validation loop's last batch, test-time augmentation) must not free them."""

    def __init__(self):
        self.ws = None
        self.row_stats = None
        self.gen = 0
        self.pool = {}

    def pooled(self, shape, device):
        ws = self.pool.get((shape, device))
        return ws


class Reshape(nn.Module):
    """kept for module-tree parity with sadecoder.py:11-18 (no parameters)."""

    def __init__(self, *args):
        super().__init__()
        self.shape = args

    def forward(self, x):
        return x.view((x.size(0),) + self.shape)


def _init_like_reference(module: nn.Module) -> None:
    """kaiming-normal(relu) conv weights, zero conv bias, BN gamma 1 / beta 0 (sadecoder.py:32-40, 70-78)."""
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight.data, nonlinearity="relu")
            if m.bias is not None:
                nn.init.constant_(m.bias.data, 0)
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1)
            m.bias.data.zero_()


class PixelAttentionBlock_(nn.Module):
    """sadecoder.py:20-52.  ``f_query`` IS ``f_key`` (one module under two names, 7 duplicated state_dict keys)."""

    def __init__(self, in_channels, key_channels):
        super().__init__()
        self.in_channels = in_channels
        self.key_channels = key_channels
        self.f_key = nn.Sequential(OrderedDict([
            ("conv", nn.Conv2d(in_channels, key_channels, kernel_size=1, stride=1, padding=0)),
            ("bn", nn.BatchNorm2d(key_channels)),
            ("relu", nn.ReLU(True))]))
        _init_like_reference(self)
        self.f_query = self.f_key
        self._holder = _Holder()

    def key_features(self, x: torch.Tensor) -> torch.Tensor:
        """F = f_key(x), computed ONCE.  The reference evaluates the shared module twice (sadecoder.py:44-45), which
        in train() advances the BN running statistics two momentum steps with the same batch moments and bumps
        num_batches_tracked by 2.  Two steps with momentum m equal one step with momentum
        1 - (1-m)^2 = m (2 - m):  r2 = (1-m)^2 r0 + (1 - (1-m)^2) mu  -- applied by running BN once with that
        momentum (nothing is modified in place after the forward, so autograd's saved buffers stay valid)."""
        bn = self.f_key.bn
        twice = self.training and bn.track_running_stats and bn.momentum is not None
        # conv -> BN -> ReLU; on channels_last half-precision activations in train() the BN + ReLU run as the fused
        # batch-statistics kernels (fn.bn_act), otherwise as the framework's own ops
        feat = fn.bn_act(bn, self.f_key.conv(x), relu=True, momentum=bn.momentum * (2.0 - bn.momentum) if twice else None)
        if twice:
            fn._bump(bn.num_batches_tracked)
        return feat

    def forward_att(self, x):
        """dense sim_map [B,N,N] (sadecoder.py:42-49)."""
        return fn.SimMap.apply(self.key_features(x).float())

    def forward(self, x):
        raise NotImplementedError


class SelfAttentionBlock_(PixelAttentionBlock_):
    """sadecoder.py:55-90 with scale = 1 (the only value SADecoder builds; the scale > 1 branch is dead and
    shape-inconsistent in the reference)."""

    def __init__(self, in_channels, key_channels, value_channels, scale=1):
        super().__init__(in_channels, key_channels)
        if scale != 1:
            raise NotImplementedError("SelfAttentionBlock_: only scale=1 exists on the reference's live path")
        self.scale = scale
        self.value_channels = value_channels
        self.f_value = nn.Sequential(OrderedDict([
            ("conv1", nn.Conv2d(in_channels, value_channels, kernel_size=3, stride=1, padding=1)),
            ("relu1", nn.ReLU(inplace=True)),
            ("conv2", nn.Conv2d(value_channels, value_channels, kernel_size=1, stride=1)),
            ("relu2", nn.ReLU(inplace=True))]))
        _init_like_reference(self)
        self.lazy_sim_map = True      # False: always materialise sim_map like the reference does

    def forward(self, x):
        # F and V go to the channels-last tensor-core path as the projections left them: 16-bit channels_last tensors (the autocast
        # training step) are read in place, fp32 NCHW tensors (the reference's own dtype) take one cast + transpose copy each
        feat = self.key_features(x)
        value = self.f_value(x)
        context, token = fn.Attention.apply(feat, value, self._holder)
        sim_map = LazySimMap(feat, self._holder, token=token, ws=self._holder.ws)
        if not self.lazy_sim_map:
            sim_map = sim_map.materialize()
        return [context, sim_map]


class SADecoder(nn.Module):
    """sadecoder.py:92-122.  Only valid for inputs whose (H, W) equal the constructor's, as in the reference."""

    def __init__(self, in_channels=2048, key_channels=512, value_channels=2048, height=224, width=304):
        super().__init__()
        out_channels = 512
        self.saconv = SelfAttentionBlock_(in_channels, key_channels, value_channels)
        # same child names as the reference so that 'image_context.linear1.weight' etc. keep their keys
        self.image_context = nn.Sequential(OrderedDict([
            ("avgpool", nn.AvgPool2d((height // 8, width // 8), padding=0)),
            ("dropout", nn.Dropout2d(0.5, inplace=True)),
            ("reshape1", Reshape(2048)),
            ("linear1", nn.Linear(2048, 512)),
            ("relu1", nn.ReLU(inplace=True)),
            ("linear2", nn.Linear(512, 512)),
            ("relu2", nn.ReLU(inplace=True)),
            ("reshape2", Reshape(512, 1, 1)),
            ("upsample", nn.Upsample(size=(height // 8, width // 8), mode="bilinear", align_corners=True))]))
        self.merge = nn.Sequential(OrderedDict([
            ("dropout1", nn.Dropout2d(0.5, inplace=True)),
            ("conv1", nn.Conv2d(value_channels + out_channels, value_channels, kernel_size=1, stride=1)),
            ("relu", nn.ReLU(inplace=True)),
            ("dropout2", nn.Dropout2d(0.5, inplace=False))]))
        self.hw = (height // 8, width // 8)
        self.value_channels = value_channels
        self.sim_map = None

    def get_sim_map(self):
        return self.sim_map

    @staticmethod
    def _channel_keep(drop: nn.Dropout2d, b: int, c: int, device) -> Optional[torch.Tensor]:
        """Dropout2d keep decisions [B,C] in {0,1}, drawn with torch's own feature-dropout so the RNG stream
        matches the reference's Dropout2d call on a [B,C,h,w] tensor (noise is [B,C,1,1] there too)."""
        if not (drop.training and drop.p > 0):
            return None
        ones = torch.ones(b, c, 1, 1, device=device)
        return (F.dropout2d(ones, drop.p, True) > 0).view(b, c).float()

    def _forward_half_inference(self, x):
        """Inference on a half-precision channels_last feature map (autocast pipelines): F, V and x are consumed in place
        by the channels-last kernels (no pack, no fp32 copies), ctx comes back fp16 channels_last for the merge GEMM."""
        b = x.shape[0]
        sa = self.saconv
        if getattr(sa, "fused_key", None) is not None:          # BN-folded, fused-epilogue convolutions (inference.py)
            feat, value = sa.fused_key(x), sa.fused_value(x)
        else:
            with torch.autocast("cuda", dtype=torch.float16):
                feat = sa.key_features(x)
                value = sa.f_value(x)
        shape_key = ((feat.shape[0], feat.shape[2] * feat.shape[3], feat.shape[1], value.shape[1]), feat.device)
        out = ops.attention_forward_cl(feat, value, ws=sa._holder.pool.get(shape_key))
        sa._holder.pool[shape_key] = out["ws"]
        sa._holder.ws = out["ws"]
        sa._holder.row_stats = out["row_stats"]
        sa._holder.gen += 1
        self.sim_map = LazySimMap(feat, sa._holder)
        ic, mg, dv = self.image_context, self.merge, self.value_channels
        _, _, g = ops.pool_branch_cl(x, ic.linear1.weight, ic.linear1.bias, ic.linear2.weight, ic.linear2.bias, None)
        w = mg.conv1.weight.view(mg.conv1.out_channels, -1)
        bias = F.linear(g, w[:, dv:].float(), mg.conv1.bias.float())
        y = F.conv2d(out["ctx"], self._ctx_weight_half(w, dv))
        if y.shape[1] % 8 == 0 and y.permute(0, 2, 3, 1).is_contiguous():
            return ops.bias_relu_cl_(y, bias), self.sim_map                  # per-image bias + ReLU in one in-place pass
        return torch.relu_(y + bias.to(y.dtype).view(b, -1, 1, 1)), self.sim_map

    def _ctx_weight_half(self, w2d: torch.Tensor, dv: int) -> torch.Tensor:
        """fp16 channels_last copy of W[:, :dv], cached until the parameter changes."""
        key = (self.merge.conv1.weight._version, self.merge.conv1.weight.data_ptr())
        if getattr(self, "_wctx_cache", (None, None))[0] != key:
            wh = w2d[:, :dv].detach().to(torch.float16).reshape(w2d.shape[0], dv, 1, 1).contiguous(memory_format=torch.channels_last)
            self._wctx_cache = (key, wh)
        return self._wctx_cache[1]

    def forward(self, x):
        b = x.shape[0]
        if tuple(x.shape[2:]) != self.hw:
            raise RuntimeError(f"SADecoder built for a {self.hw} feature map, got {tuple(x.shape[2:])} (sadecoder.py:98,106)")
        if (not self.training and not torch.is_grad_enabled() and x.dtype in (torch.float16, torch.bfloat16)
                and self.saconv.lazy_sim_map and self.value_channels % 128 == 0):
            return self._forward_half_inference(x)
        context, sim_map = self.saconv(x)
        self.sim_map = sim_map
        # image-level pooling branch: mean -> dropout -> MLP, never broadcast (sadecoder.py:97-106)
        ic = self.image_context
        keep_pool = self._channel_keep(ic.dropout, b, x.shape[1], x.device)
        half_cl = x.dtype in (torch.float16, torch.bfloat16) and x.permute(0, 2, 3, 1).is_contiguous()
        g = fn.PoolBranch.apply(x if half_cl else x.float(), ic.linear1.weight, ic.linear1.bias, ic.linear2.weight, ic.linear2.bias, keep_pool)
        # merge = ReLU(W [context ; g] + b): GEMM over the dv context channels, the spatially constant pooled
        # branch folds into a per-image bias W[:, dv:] g  (sadecoder.py:107-111,120-121); no cat, no broadcast
        mg = self.merge
        dv = self.value_channels
        w = mg.conv1.weight.view(mg.conv1.out_channels, -1)
        keep_in = self._channel_keep(mg.dropout1, b, w.shape[1], x.device)
        if keep_in is not None:
            context = context * (2.0 * keep_in[:, :dv]).view(b, dv, 1, 1)
            g = g * (2.0 * keep_in[:, dv:])
        bias = F.linear(g, w[:, dv:], mg.conv1.bias)                                  # [B, dv]
        out = F.conv2d(context, self._ctx_weight(w, dv)) + bias.view(b, -1, 1, 1)
        out = mg.dropout2(torch.relu_(out))
        return out, self.sim_map

    def _ctx_weight(self, w2d: torch.Tensor, dv: int) -> torch.Tensor:
        """W[:, :dv] as a contiguous 1x1 conv weight (cuDNN wants it dense); differentiable slice."""
        return w2d[:, :dv].reshape(w2d.shape[0], dv, 1, 1)


class AttentionLoss2d(nn.Module):
    """losses.py:35-62.  ``forward(sim_map, label)`` with ``label`` = bin indices [B,1,H,W] (model.py:260,270).
    A ``LazySimMap`` runs the fused tensor-core KL (sim_map never in HBM); a dense tensor runs the
    single-pass streaming kernel."""

    def __init__(self, scale=1):
        super().__init__()
        if scale != 1:
            raise NotImplementedError("AttentionLoss2d: the reference always builds scale=1 (get_lossfunc.py:35)")
        self.scale = scale

    def get_gt_sim_map(self, label):
        return ops.gt_sim_map(label)

    def forward(self, sim_map, label):
        if isinstance(sim_map, LazySimMap):
            if sim_map._token is not None and sim_map._own_ws is not None and sim_map._slice is None:
                # training: value from the fused KL over that forward's workspace, gradient through the fn.Attention node it belongs to
                return fn.AttentionLossOnToken.apply(sim_map._token, label, sim_map._own_ws)
            return FusedAttentionLoss.apply(sim_map._feat, label, sim_map)
        return fn.AttentionLossMaterialised.apply(sim_map, label)


class FusedAttentionLoss(torch.autograd.Function):
    """loss from the operand pack of the preceding attention forward: S is recomputed tile by tile on the tensor
    cores and compared on chip with the target generated from the labels (acan_attention_loss_fused)."""

    @staticmethod
    def forward(ctx, feat, dis_label, lazy):
        ctx.save_for_backward(feat, dis_label)
        ws = lazy._ws()
        ctx.row_stats = lazy._holder.row_stats if ws is lazy._holder.ws else ops.attention_forward(feat.float(), None)["row_stats"]
        return ops.attention_loss_fused(ws, dis_label)

    @staticmethod
    def backward(ctx, grad_out):
        feat, dis_label = ctx.saved_tensors
        g, _ = ops.attention_backward(feat, ctx.row_stats, dis_label=dis_label, grad_loss=grad_out)   # loss-only branch of acan_attn_bwd
        return g.view_as(feat), None, None
