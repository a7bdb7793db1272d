"""Tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw device pointers and the
current CUDA stream to ``libacan_b200.so``.  No autograd here (see ``functions.py``), no fallback.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib

HEAD_KINDS = {("OR", "soft"): 0, ("OR", "hard"): 1, ("CE", "soft"): 2, ("CE", "hard"): 3}


def _cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.AcanLibraryError(f"{name} must be a CUDA tensor: acan_b200 has no CPU path")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ------------------------------------------------------------------------------------ head

def decode_ord(logits: torch.Tensor) -> torch.Tensor:
    """[B,2K,H,W] -> [B,K,H,W]   (model.py:40-45)"""
    y = _cuda_f32(logits, "logits")
    b, c2, h, w = y.shape
    out = torch.empty((b, c2 // 2, h, w), device=y.device, dtype=torch.float32)
    with torch.cuda.device(y.device):
        _lib.check(_lib.load().acan_decode_ord_fwd(y.data_ptr(), out.data_ptr(), b, c2 // 2, h * w, _lib.current_stream()), "acan_decode_ord_fwd")
    return out


def decode_ord_backward(prob: torch.Tensor, grad_prob: torch.Tensor) -> torch.Tensor:
    p = _cuda_f32(prob, "prob")
    g = _cuda_f32(grad_prob, "grad_prob")
    b, k, h, w = p.shape
    out = torch.empty((b, 2 * k, h, w), device=p.device, dtype=torch.float32)
    with torch.cuda.device(p.device):
        _lib.check(_lib.load().acan_decode_ord_bwd(p.data_ptr(), g.data_ptr(), out.data_ptr(), b, k, h * w, _lib.current_stream()), "acan_decode_ord_bwd")
    return out


def head_inference(x: torch.Tensor, classifier: str, mode: str, d_min: float, d_max: float, n_c: int,
                   from_logits: bool = False, want_index: bool = False):
    """depth [B,1,H,W] (and optionally the int32 bin index) from probabilities / scores / pair logits (model.py:47-99)."""
    kind = HEAD_KINDS[("OR" if classifier == "OR" else "CE", "soft" if mode == "soft" else "hard")]
    t = _cuda_f32(x, "head input")
    b, c, h, w = t.shape
    want_c = 2 * n_c if (from_logits and classifier == "OR") else n_c
    if c != want_c:
        raise _lib.AcanLibraryError(f"head input has {c} channels, expected {want_c}")
    depth = torch.empty((b, 1, h, w), device=t.device, dtype=torch.float32)
    index = torch.empty((b, 1, h, w), device=t.device, dtype=torch.int32) if want_index else None
    with torch.cuda.device(t.device):
        _lib.check(_lib.load().acan_head_fwd(t.data_ptr(), depth.data_ptr(), _lib.ptr(index), b, n_c, h * w, float(d_min), float(d_max),
                                             kind, int(bool(from_logits)), _lib.current_stream()), "acan_head_fwd")
    return (depth, index) if want_index else depth


def tail_inference(z: torch.Tensor, classifier: str, mode: str, d_min: float, d_max: float, n_c: int,
                   want_prob: bool = False, want_index: bool = False):
    """Fused classifier tail: stride-8 logits z [B,C,h,w] (any float dtype / layout) -> depth [B,1,8h,8w]
    (+ 'y' = decode_ord(upsample(z)) [B,K,8h,8w] and / or the int32 bin index)."""
    kind = HEAD_KINDS[("OR" if classifier == "OR" else "CE", "soft" if mode == "soft" else "hard")]
    if not z.is_cuda:
        raise _lib.AcanLibraryError("z must be a CUDA tensor: acan_b200 has no CPU path")
    zz = z.permute(0, 2, 3, 1).float().contiguous()        # [B,h,w,C] fp32: 0.9 MB per NYU image
    b, h, w, c = zz.shape
    if c != (2 * n_c if classifier == "OR" else n_c):
        raise _lib.AcanLibraryError(f"tail input has {c} channels")
    depth = torch.empty((b, 1, 8 * h, 8 * w), device=zz.device, dtype=torch.float32)
    prob = torch.empty((b, n_c, 8 * h, 8 * w), device=zz.device, dtype=torch.float32) if want_prob else None
    index = torch.empty((b, 1, 8 * h, 8 * w), device=zz.device, dtype=torch.int32) if want_index else None
    with torch.cuda.device(zz.device):
        _lib.check(_lib.load().acan_tail_fwd(zz.data_ptr(), depth.data_ptr(), _lib.ptr(prob), _lib.ptr(index), b, h, w, n_c,
                                             float(d_min), float(d_max), kind, _lib.current_stream()), "acan_tail_fwd")
    out = (depth,) + ((prob,) if want_prob else ()) + ((index,) if want_index else ())
    return out[0] if len(out) == 1 else out


def ordinal_loss_forward(logits: torch.Tensor, label: torch.Tensor, ignore_index: int = 0):
    """(loss 0-dim, count 0-dim): OrdinalRegression2d fused with decode_ord on full-resolution pair logits [B,2K,H,W]."""
    y = _cuda_f32(logits, "logits")
    b, c2, h, w = y.shape
    lab = label.to(device=y.device, dtype=torch.int64).contiguous()
    if lab.numel() != b * h * w:
        raise _lib.AcanLibraryError(f"label {tuple(lab.shape)} does not match logits {tuple(y.shape)}")
    lib = _lib.load()
    loss = torch.empty((), device=y.device, dtype=torch.float32)
    count = torch.empty((), device=y.device, dtype=torch.float32)
    ws = torch.empty(int(lib.acan_ordinal_loss_ws_floats()), device=y.device, dtype=torch.float32)
    with torch.cuda.device(y.device):
        _lib.check(lib.acan_ordinal_loss_fwd(y.data_ptr(), lab.data_ptr(), loss.data_ptr(), count.data_ptr(), ws.data_ptr(), b, c2 // 2, h * w,
                                             int(ignore_index), _lib.current_stream()), "acan_ordinal_loss_fwd")
    return loss, count, y, lab


def ordinal_loss_backward(logits: torch.Tensor, label: torch.Tensor, count: torch.Tensor, grad_out: torch.Tensor, ignore_index: int = 0):
    b, c2, h, w = logits.shape
    g = grad_out.detach().to(device=logits.device, dtype=torch.float32).reshape(1).contiguous()
    out = torch.empty_like(logits)
    with torch.cuda.device(logits.device):
        _lib.check(_lib.load().acan_ordinal_loss_bwd(logits.data_ptr(), label.data_ptr(), count.data_ptr(), g.data_ptr(), out.data_ptr(), b, c2 // 2, h * w,
                                                     int(ignore_index), _lib.current_stream()), "acan_ordinal_loss_bwd")
    return out


def ce_loss_forward(scores: torch.Tensor, label: torch.Tensor, ignore_index: int, eps: float = 0.0, prior: str = "uniform"):
    """(loss 0-dim, count 0-dim, scores fp32, label int64): CrossEntropy2d(reduction='sum') on class scores [B,K,H,W] (losses.py:164-196)."""
    y = _cuda_f32(scores, "scores")
    b, k, h, w = y.shape
    lab = label.to(device=y.device, dtype=torch.int64).contiguous()
    if lab.numel() != b * h * w:
        raise _lib.AcanLibraryError(f"label {tuple(lab.shape)} does not match scores {tuple(y.shape)}")
    lib = _lib.load()
    loss = torch.empty((), device=y.device, dtype=torch.float32)
    count = torch.empty((), device=y.device, dtype=torch.float32)
    ws = torch.empty(int(lib.acan_ordinal_loss_ws_floats()), device=y.device, dtype=torch.float32)
    with torch.cuda.device(y.device):
        _lib.check(lib.acan_ce_loss_fwd(y.data_ptr(), lab.data_ptr(), loss.data_ptr(), count.data_ptr(), ws.data_ptr(), b, k, h * w, int(ignore_index),
                                        float(eps), 1 if prior == "gaussian" else 0, _lib.current_stream()), "acan_ce_loss_fwd")
    return loss, count, y, lab


def ce_loss_backward(scores, label, count, grad_out, ignore_index: int, eps: float = 0.0, prior: str = "uniform"):
    b, k, h, w = scores.shape
    g = grad_out.detach().to(device=scores.device, dtype=torch.float32).reshape(1).contiguous()
    out = torch.empty_like(scores)
    with torch.cuda.device(scores.device):
        _lib.check(_lib.load().acan_ce_loss_bwd(scores.data_ptr(), label.data_ptr(), count.data_ptr(), g.data_ptr(), out.data_ptr(), b, k, h * w,
                                                int(ignore_index), float(eps), 1 if prior == "gaussian" else 0, _lib.current_stream()), "acan_ce_loss_bwd")
    return out


def tail_loss_forward(z: torch.Tensor, label: torch.Tensor, ignore_index: int = 0):
    """(loss, count, z_nhwc, label): OrdinalRegression2d(decode_ord(upsample_x8(z))) from the stride-8 pair logits z [B,2K,h,w]
    (any float dtype / layout); label [B,8h,8w] bin indices."""
    if not z.is_cuda:
        raise _lib.AcanLibraryError("z must be a CUDA tensor: acan_b200 has no CPU path")
    zz = z.detach().permute(0, 2, 3, 1).float().contiguous()        # [B,h,w,2K] fp32: 0.9 MB per NYU image
    b, h, w, c2 = zz.shape
    lab = label.to(device=zz.device, dtype=torch.int64).contiguous()
    if lab.numel() != b * 64 * h * w:
        raise _lib.AcanLibraryError(f"label {tuple(lab.shape)} does not match stride-8 logits {tuple(z.shape)}")
    lib = _lib.load()
    loss = torch.empty((), device=zz.device, dtype=torch.float32)
    count = torch.empty((), device=zz.device, dtype=torch.float32)
    ws = torch.empty(int(lib.acan_tail_loss_ws_floats(b, h, w)), device=zz.device, dtype=torch.float32)
    with torch.cuda.device(zz.device):
        _lib.check(lib.acan_tail_loss_fwd(zz.data_ptr(), lab.data_ptr(), loss.data_ptr(), count.data_ptr(), ws.data_ptr(), b, h, w, c2 // 2,
                                          int(ignore_index), _lib.current_stream()), "acan_tail_loss_fwd")
    return loss, count, zz, lab


def tail_loss_backward(z_nhwc: torch.Tensor, label: torch.Tensor, count: torch.Tensor, grad_out: torch.Tensor, ignore_index: int = 0):
    """grad of the fused tail loss w.r.t. z, returned as a [B,2K,h,w] view of a channels-last buffer."""
    b, h, w, c2 = z_nhwc.shape
    g = grad_out.detach().to(device=z_nhwc.device, dtype=torch.float32).reshape(1).contiguous()
    out = torch.empty_like(z_nhwc)
    with torch.cuda.device(z_nhwc.device):
        _lib.check(_lib.load().acan_tail_loss_bwd(z_nhwc.data_ptr(), label.data_ptr(), count.data_ptr(), g.data_ptr(), out.data_ptr(), b, h, w, c2 // 2,
                                                  int(ignore_index), _lib.current_stream()), "acan_tail_loss_bwd")
    return out.permute(0, 3, 1, 2)


# ------------------------------------------------------------------------------------ training-mode BN (+ ReLU, + residual)

_BN_WS = {}          # (device index, stream) -> zero-initialised scratch (arrival counters + partials), grown on demand
_BN_WS_RETIRED = []  # outgrown buffers stay alive: a captured CUDA graph may still hold their address


def _bn_ws(device: torch.device, c: int) -> torch.Tensor:
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    need = int(_lib.load().acan_bn_ws_floats(c))
    t = _BN_WS.get(key)
    if t is None or t.numel() < need:
        if t is not None:
            _BN_WS_RETIRED.append(t)
        t = torch.zeros(max(need, int(_lib.load().acan_bn_ws_floats(2048))), device=device, dtype=torch.float32)
        _BN_WS[key] = t
    return t


def bn_supported(x: torch.Tensor) -> bool:
    """channels-last (or 1x1-spatial) CUDA fp16 / bf16 activations with C % 8 == 0: what the fused BN kernels take."""
    return (x.is_cuda and x.dim() == 4 and x.dtype in (torch.float16, torch.bfloat16) and x.shape[1] % 8 == 0
            and x.numel() > 0 and x.permute(0, 2, 3, 1).is_contiguous())


def bn_train_forward(x, weight, bias, running_mean, running_var, momentum, eps, relu, residual=None):
    """(y, save_mean, save_invstd).  x [B,C,H,W] channels_last half; running statistics updated in place (may be None)."""
    b, c, h, w = x.shape
    y = torch.empty_like(x)                                  # preserves the channels_last strides
    mean = torch.empty(c, device=x.device, dtype=torch.float32)
    invstd = torch.empty(c, device=x.device, dtype=torch.float32)
    if residual is not None and not (residual.dtype == x.dtype and residual.shape == x.shape and residual.permute(0, 2, 3, 1).is_contiguous()):
        residual = residual.to(x.dtype).contiguous(memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().acan_bn_train_fwd(x.data_ptr(), 1 if x.dtype == torch.bfloat16 else 0, weight.data_ptr(), bias.data_ptr(), _lib.ptr(residual),
                                                 y.data_ptr(), mean.data_ptr(), invstd.data_ptr(), _lib.ptr(running_mean), _lib.ptr(running_var),
                                                 float(momentum), float(eps), int(bool(relu)), _bn_ws(x.device, c).data_ptr(), b * h * w, c,
                                                 _lib.current_stream()), "acan_bn_train_fwd")
    return y, mean, invstd


def bn_train_backward(x, y, grad_y, weight, mean, invstd, want_residual_grad):
    """(grad_x, grad_residual | None, grad_weight, grad_bias); y = forward output when a ReLU was fused, else None."""
    b, c, h, w = x.shape
    if not (grad_y.dtype == x.dtype and grad_y.permute(0, 2, 3, 1).is_contiguous()):
        grad_y = grad_y.to(x.dtype).contiguous(memory_format=torch.channels_last)
    gx = torch.empty_like(x)
    gres = torch.empty_like(x) if (want_residual_grad and y is not None) else None
    gw = torch.empty(c, device=x.device, dtype=torch.float32)
    gb = torch.empty(c, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().acan_bn_train_bwd(x.data_ptr(), _lib.ptr(y), grad_y.data_ptr(), 1 if x.dtype == torch.bfloat16 else 0, weight.data_ptr(),
                                                 mean.data_ptr(), invstd.data_ptr(), gx.data_ptr(), _lib.ptr(gres), gw.data_ptr(), gb.data_ptr(),
                                                 _bn_ws(x.device, c).data_ptr(), b * h * w, c, _lib.current_stream()), "acan_bn_train_bwd")
    if want_residual_grad and gres is None:
        gres = grad_y                                        # no ReLU in between: the upstream gradient passes through unchanged
    return gx, gres, gw, gb


def continuous2discrete(depth: torch.Tensor, d_min: float, d_max: float, n_c: int) -> torch.Tensor:
    d = _cuda_f32(depth, "depth")
    out = torch.empty_like(d)
    with torch.cuda.device(d.device):
        _lib.check(_lib.load().acan_continuous2discrete(d.data_ptr(), out.data_ptr(), d.numel(), float(d_min), float(d_max), n_c, _lib.current_stream()),
                   "acan_continuous2discrete")
    return out


# ------------------------------------------------------------------------------------ pooling branch

def pool_branch(x: torch.Tensor, w1, b1, w2, b2, keep_mask: Optional[torch.Tensor] = None):
    """(mean [B,C], hidden [B,H1], g [B,H2])   (sadecoder.py:97-106)"""
    xx = _cuda_f32(x, "x")
    b, c = xx.shape[:2]
    n = xx.numel() // (b * c)
    w1, b1, w2, b2 = (_cuda_f32(t, "pool weight") for t in (w1, b1, w2, b2))
    h1, h2 = w1.shape[0], w2.shape[0]
    mean = torch.empty((b, c), device=xx.device, dtype=torch.float32)
    hid = torch.empty((b, h1), device=xx.device, dtype=torch.float32)
    g = torch.empty((b, h2), device=xx.device, dtype=torch.float32)
    km = _cuda_f32(keep_mask, "keep_mask") if keep_mask is not None else None
    with torch.cuda.device(xx.device):
        _lib.check(_lib.load().acan_pool_branch_fwd(xx.data_ptr(), _lib.ptr(km), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                                    mean.data_ptr(), hid.data_ptr(), g.data_ptr(), b, c, n, h1, h2, _lib.current_stream()),
                   "acan_pool_branch_fwd")
    return mean, hid, g


def pool_branch_cl(x: torch.Tensor, w1, b1, w2, b2, keep_mask: Optional[torch.Tensor] = None):
    """pool_branch for a half-precision (fp16 / bf16) feature map held in channels_last memory: read in place."""
    if not x.is_cuda or x.dtype not in (torch.float16, torch.bfloat16):
        raise _lib.AcanLibraryError("pool_branch_cl wants a CUDA fp16 / bf16 tensor")
    xx = x.permute(0, 2, 3, 1).contiguous()          # no-op for channels_last input
    b, h, w, c = xx.shape
    w1, b1, w2, b2 = (_cuda_f32(t, "pool weight") for t in (w1, b1, w2, b2))
    h1, h2 = w1.shape[0], w2.shape[0]
    lib = _lib.load()
    scratch = torch.empty(int(lib.acan_pool_nhwc_scratch_floats(b, c, h * w)), device=xx.device, dtype=torch.float32)
    mean = torch.empty((b, c), device=xx.device, dtype=torch.float32)
    hid = torch.empty((b, h1), device=xx.device, dtype=torch.float32)
    g = torch.empty((b, h2), device=xx.device, dtype=torch.float32)
    km = _cuda_f32(keep_mask, "keep_mask") if keep_mask is not None else None
    with torch.cuda.device(xx.device):
        _lib.check(lib.acan_pool_branch_fwd_nhwc(xx.data_ptr(), 1 if xx.dtype == torch.float16 else 2, _lib.ptr(km), w1.data_ptr(), b1.data_ptr(),
                                                 w2.data_ptr(), b2.data_ptr(), scratch.data_ptr(), mean.data_ptr(), hid.data_ptr(), g.data_ptr(),
                                                 b, c, h * w, h1, h2, _lib.current_stream()), "acan_pool_branch_fwd_nhwc")
    return mean, hid, g


def maxpool3x3s2_cl(x: torch.Tensor) -> torch.Tensor:
    """nn.MaxPool2d(3, 2, 1) on a CUDA fp16 / bf16 channels_last tensor [B,C,H,W] (inference: no indices, no autograd)."""
    if not x.is_cuda or x.dtype not in (torch.float16, torch.bfloat16) or not x.permute(0, 2, 3, 1).is_contiguous() or x.shape[1] % 8:
        raise _lib.AcanLibraryError("maxpool3x3s2_cl wants a CUDA fp16 / bf16 channels_last tensor with C % 8 == 0")
    b, c, h, w = x.shape
    y = torch.empty((b, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().acan_maxpool3x3s2_nhwc(x.data_ptr(), y.data_ptr(), 1 if x.dtype == torch.float16 else 2, b, h, w, c, _lib.current_stream()),
                   "acan_maxpool3x3s2_nhwc")
    return y


def bias_relu_cl_(y: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """y <- relu(y + bias[:, :, None, None]) in place; y [B,C,h,w] fp16 / bf16 in channels_last memory, bias [B,C]."""
    if not y.is_cuda or y.dtype not in (torch.float16, torch.bfloat16) or not y.permute(0, 2, 3, 1).is_contiguous():
        raise _lib.AcanLibraryError("bias_relu_cl_ wants a CUDA fp16 / bf16 channels_last tensor")
    b, c, h, w = y.shape
    bb = _cuda_f32(bias, "bias").reshape(b, c)
    with torch.cuda.device(y.device):
        _lib.check(_lib.load().acan_bias_relu_nhwc(y.data_ptr(), 1 if y.dtype == torch.float16 else 2, bb.data_ptr(), b, h * w, c, _lib.current_stream()),
                   "acan_bias_relu_nhwc")
    return y


# ------------------------------------------------------------------------------------ attention

class AttnWorkspace:
    """Scratch for the attention passes; keeps the fp16 operand pack + row statistics of the last forward so
    selected sim_map rows / the fused loss can be produced later without re-packing."""

    def __init__(self, b: int, n: int, dk: int, dv: int, device):
        self.shape = (b, n, dk, dv)
        self.bytes = int(_lib.load().acan_attn_workspace_bytes(b, n, dk, dv))
        # torch's caching allocator returns >= 512 B aligned blocks; over-allocate and align to 1024 ourselves
        self._buf = torch.empty(self.bytes + 1024, device=device, dtype=torch.uint8)
        base = self._buf.data_ptr()
        self.ptr = (base + 1023) // 1024 * 1024
        self.scale = None
        self.feat_pack = None     # channels-last fp16 F of the last forward when it was consumed in place (kept alive here)
        self.has_target = False   # ln(bin) / ln Z / T_i of an attention-loss forward are in the workspace
        self.pending_loss_grad = None   # upstream gradient of the attention loss, parked here by its autograd node for the fused backward


def attention_forward(feat: torch.Tensor, value: Optional[torch.Tensor], want_sim_map: bool = False,
                      dis_label: Optional[torch.Tensor] = None, ws: Optional[AttnWorkspace] = None):
    """Returns dict(ctx [B,dv,h,w] | None, sim_map [B,N,N] | None, row_stats [B,N,2], loss 0-dim | None, ws)."""
    f = _cuda_f32(feat, "feat")
    b, dk, h, w = f.shape
    n = h * w
    v = _cuda_f32(value, "value") if value is not None else None
    dv = v.shape[1] if v is not None else 0
    if ws is None or ws.shape != (b, n, dk, dv) or ws._buf.device != f.device:
        ws = AttnWorkspace(b, n, dk, dv, f.device)
    scale = float(dk) ** -0.5
    ws.scale = scale
    ctx = torch.empty((b, dv, h, w), device=f.device, dtype=torch.float32) if v is not None else None
    sim = torch.empty((b, n, n), device=f.device, dtype=torch.float32) if want_sim_map else None
    stats = torch.empty((b, n, 2), device=f.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(f.device):
        st = _lib.current_stream()
        if dis_label is None:
            loss = None
            _lib.check(lib.acan_attn_fwd(f.data_ptr(), _lib.ptr(v), _lib.ptr(ctx), _lib.ptr(sim), stats.data_ptr(), ws.ptr, ws.bytes,
                                         b, n, dk, dv, scale, st), "acan_attn_fwd")
        else:
            lab = _cuda_f32(dis_label, "dis_label")
            hh, ww = lab.shape[-2:]
            loss = torch.empty((), device=f.device, dtype=torch.float32)
            _lib.check(lib.acan_attn_fwd_loss(f.data_ptr(), _lib.ptr(v), _lib.ptr(ctx), _lib.ptr(sim), stats.data_ptr(), lab.data_ptr(), hh, ww,
                                              loss.data_ptr(), ws.ptr, ws.bytes, b, n, dk, dv, scale, st), "acan_attn_fwd_loss")
    ws.feat_pack = None
    return {"ctx": ctx, "sim_map": sim, "row_stats": stats, "loss": loss, "ws": ws}


def _nhwc_half(t: torch.Tensor, name: str) -> torch.Tensor:
    """[B,C,h,w] (any layout / float dtype) -> dense [B,h,w,C] fp16; a no-op view for fp16 channels_last input."""
    if not t.is_cuda:
        raise _lib.AcanLibraryError(f"{name} must be a CUDA tensor: acan_b200 has no CPU path")
    t = t.permute(0, 2, 3, 1)
    if t.dtype != torch.float16:
        t = t.to(torch.float16)
    return t.contiguous()


def attention_forward_cl(feat: torch.Tensor, value: Optional[torch.Tensor], want_sim_map: bool = False,
                         dis_label: Optional[torch.Tensor] = None, ws: Optional[AttnWorkspace] = None, fused: bool = False):
    """Channels-last half-precision attention: feat [B,dk,h,w] / value [B,dv,h,w] as the convolutions left them
    (fp16, channels_last memory -> consumed in place, no pack); ctx comes back [B,dv,h,w] fp16 channels_last.
    fused (inference, no sim_map / loss): probability and aggregation tiles in ONE persistent launch instead of two -- same arithmetic,
    bit-identical results, measured slower (both passes are shared-memory-bandwidth bound: csrc/attention.cu), hence off by default."""
    f = _nhwc_half(feat, "feat")
    b, h, w, dk = f.shape
    n = h * w
    v = _nhwc_half(value, "value") if value is not None else None
    dv = v.shape[3] if v is not None else 0
    if ws is None or ws.shape != (b, n, dk, dv) or ws._buf.device != f.device:
        ws = AttnWorkspace(b, n, dk, dv, f.device)
    scale = float(dk) ** -0.5
    ws.scale = scale
    ctx = torch.empty((b, h, w, dv), device=f.device, dtype=torch.float16) if v is not None else None
    sim = torch.empty((b, n, n), device=f.device, dtype=torch.float32) if want_sim_map else None
    stats = torch.empty((b, n, 2), device=f.device, dtype=torch.float32)
    lab = _cuda_f32(dis_label, "dis_label") if dis_label is not None else None
    hh, ww = (lab.shape[-2:] if lab is not None else (0, 0))
    loss = torch.empty((), device=f.device, dtype=torch.float32) if lab is not None else None
    with torch.cuda.device(f.device):
        _lib.check(_lib.load().acan_attn_fwd_nhwc(f.data_ptr(), 1, _lib.ptr(v), 1, _lib.ptr(ctx), 1, _lib.ptr(sim), stats.data_ptr(), _lib.ptr(lab), hh, ww,
                                                  _lib.ptr(loss), ws.ptr, ws.bytes, b, n, dk, dv, scale, 2 if fused else 0, _lib.current_stream()), "acan_attn_fwd_nhwc")
    ws.feat_pack = f
    return {"ctx": ctx.permute(0, 3, 1, 2) if ctx is not None else None, "sim_map": sim, "row_stats": stats, "loss": loss, "ws": ws}


_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _nhwc_dense(t: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """[B,C,h,w] (any layout) -> dense [B,h,w,C] in `dtype` (default: its own); a view when it already is, else ONE cast + transpose copy."""
    v = t.permute(0, 2, 3, 1)
    if (dtype is None or v.dtype == dtype) and v.is_contiguous():
        return v
    out = torch.empty(v.shape, device=t.device, dtype=dtype or t.dtype)
    out.copy_(v)
    return out


def attention_forward_train(feat: torch.Tensor, value: torch.Tensor, dis_label: Optional[torch.Tensor] = None):
    """Training-mode attention (acan_attn_fwd_nhwc, keep = 1) on whatever the projections produced:
      feat  [B,dk,h,w]  fp16 channels_last is read in place; bf16 / fp32 take one pass into the workspace's fp16 pack
      value [B,dv,h,w]  fp16 channels_last is read in place; bf16 (the autocast step: exact) / fp32 take one cast (+ transpose) copy to
                        fp16 -- both operands of a kind::f16 MMA must share one format, and the probabilities need fp16's 11 bits
    Returns dict(ctx32 [B,h,w,dv] fp32, row_stats, loss | None, ws, feat_pack [B,h,w,dk] fp16 | None, value16 [B,h,w,dv]).
    The workspace is NEW for every call and belongs to the caller's autograd node: it keeps the probabilities for the backward pass."""
    if not (feat.is_cuda and value.is_cuda):
        raise _lib.AcanLibraryError("attention inputs must be CUDA tensors: acan_b200 has no CPU path")
    b, dk, h, w = feat.shape
    dv = value.shape[1]
    n = h * w
    f = _nhwc_dense(feat.detach(), None if feat.dtype in (torch.float16, torch.bfloat16) else torch.float16)
    # bf16 values are converted by the library into the workspace (14 us at configs[3] against 60 us for the framework's strided copy)
    v = _nhwc_dense(value.detach(), value.dtype if value.dtype in (torch.float16, torch.bfloat16) else torch.float16)
    ws = AttnWorkspace(b, n, dk, dv, f.device)
    ws.scale = float(dk) ** -0.5
    ctx32 = torch.empty((b, h, w, dv), device=f.device, dtype=torch.float32)
    stats = torch.empty((b, n, 2), device=f.device, dtype=torch.float32)
    lab = _cuda_f32(dis_label, "dis_label") if dis_label is not None else None
    hh, ww = (lab.shape[-2:] if lab is not None else (0, 0))
    loss = torch.empty((), device=f.device, dtype=torch.float32) if lab is not None else None
    with torch.cuda.device(f.device):
        _lib.check(_lib.load().acan_attn_fwd_nhwc(f.data_ptr(), _DT[f.dtype], v.data_ptr(), _DT[v.dtype], ctx32.data_ptr(), 0, 0, stats.data_ptr(),
                                                  _lib.ptr(lab), hh, ww, _lib.ptr(loss), ws.ptr, ws.bytes, b, n, dk, dv, ws.scale, 1, _lib.current_stream()),
                   "acan_attn_fwd_nhwc")
    ws.feat_pack = f if f.dtype == torch.float16 else None      # None: the fp16 pack lives in the workspace
    ws.has_target = lab is not None
    return {"ctx32": ctx32, "row_stats": stats, "loss": loss, "ws": ws, "feat_pack": ws.feat_pack, "value16": v if v.dtype == torch.float16 else None}


_bwd_scratch = {}          # (device, bytes) -> buffer; entries are never dropped (captured CUDA graphs hold their addresses)


def attention_backward_cl(ws: "AttnWorkspace", value16: torch.Tensor, ctx32: torch.Tensor, grad_ctx: Optional[torch.Tensor],
                          grad_loss: Optional[torch.Tensor], gf_dtype: torch.dtype, gv_dtype: torch.dtype):
    """(grad_feat [B,h,w,dk], grad_value [B,h,w,dv] | None) from acan_attn_bwd_nhwc.  `ws`: the forward's workspace (probabilities, row sums;
    target statistics if the attention loss ran on it).  grad_ctx [B,dv,h,w] any layout: 16-bit channels_last is read in place.
    grad_loss: 0-dim CUDA tensor (upstream gradient of the attention loss) or None."""
    b, n, dk, dv = ws.shape
    dev = ws._buf.device
    hh, wd = ctx32.shape[1], ctx32.shape[2]
    agg = grad_ctx is not None
    go = None
    if agg:
        go = _nhwc_dense(grad_ctx.detach(), grad_ctx.dtype if grad_ctx.dtype in _DT else torch.float32)
    lib = _lib.load()
    nbytes = int(lib.acan_attn_bwd_nhwc_scratch_bytes(b, n, dk, dv))
    key = (dev, nbytes)
    buf = _bwd_scratch.get(key)
    if buf is None:
        buf = _bwd_scratch[key] = torch.empty(nbytes + 1024, device=dev, dtype=torch.uint8)
    sptr = (buf.data_ptr() + 1023) // 1024 * 1024
    gf = torch.empty((b, hh, wd, dk), device=dev, dtype=gf_dtype)
    gv = torch.empty((b, hh, wd, dv), device=dev, dtype=gv_dtype) if agg else None
    gl = grad_loss.detach().to(device=dev, dtype=torch.float32).reshape(1) if grad_loss is not None else None
    with torch.cuda.device(dev):
        _lib.check(lib.acan_attn_bwd_nhwc(ws.ptr, ws.bytes, sptr, nbytes, _lib.ptr(ws.feat_pack), _lib.ptr(value16) if agg else 0,
                                          1, _lib.ptr(ctx32) if agg else 0, _lib.ptr(go), _DT[go.dtype] if agg else 1,
                                          int(gl is not None), _lib.ptr(gl), gf.data_ptr(), _DT[gf_dtype], _lib.ptr(gv), _DT[gv_dtype],
                                          b, n, dk, dv, ws.scale, _lib.current_stream()), "acan_attn_bwd_nhwc")
    return gf, gv


def attention_backward(feat: torch.Tensor, row_stats: torch.Tensor, value: Optional[torch.Tensor] = None,
                       ctx: Optional[torch.Tensor] = None, grad_ctx: Optional[torch.Tensor] = None,
                       dis_label: Optional[torch.Tensor] = None, grad_loss=0.0):
    """(grad_feat, grad_value | None) from the tcgen05 backward (acan_attn_bwd).  ``grad_loss``: the upstream gradient of the
    attention loss, a 0-dim CUDA tensor (no host synchronisation: the C ABI takes a device scalar) or a python float."""
    f = _cuda_f32(feat, "feat")
    b, dk, h, w = f.shape
    n = h * w
    st = _cuda_f32(row_stats, "row_stats")
    agg = grad_ctx is not None
    v = _cuda_f32(value, "value") if agg else None
    o = _cuda_f32(ctx, "ctx") if agg else None
    go = _cuda_f32(grad_ctx, "grad_ctx") if agg else None
    dv = v.shape[1] if agg else 128
    lab = _cuda_f32(dis_label, "dis_label") if dis_label is not None else None
    hh, ww = (lab.shape[-2:] if lab is not None else (0, 0))
    lib = _lib.load()
    nbytes = int(lib.acan_attn_bwd_workspace_bytes(b, n, dk, dv))
    key = (f.device, nbytes)
    if key not in _bwd_scratch:                      # one scratch buffer per (device, size), reused across steps, never dropped
        _bwd_scratch[key] = torch.empty(nbytes + 1024, device=f.device, dtype=torch.uint8)
    buf = _bwd_scratch[key]
    ptr = (buf.data_ptr() + 1023) // 1024 * 1024
    gf = torch.empty_like(f)
    gv = torch.empty_like(v) if agg else None
    if lab is None:
        gl = None
    elif isinstance(grad_loss, torch.Tensor):
        gl = grad_loss.detach().to(device=f.device, dtype=torch.float32).reshape(1)
    else:
        gl = torch.full((1,), float(grad_loss), device=f.device, dtype=torch.float32)
    with torch.cuda.device(f.device):
        _lib.check(lib.acan_attn_bwd(f.data_ptr(), _lib.ptr(v), _lib.ptr(o), st.data_ptr(), _lib.ptr(go), _lib.ptr(lab), hh, ww, _lib.ptr(gl),
                                     gf.data_ptr(), _lib.ptr(gv), ptr, nbytes, b, n, dk, dv, float(dk) ** -0.5, _lib.current_stream()), "acan_attn_bwd")
    return gf, gv


def attention_rows(ws: AttnWorkspace, rows: torch.Tensor) -> torch.Tensor:
    """sim_map[:, rows, :] -> [B, len(rows), N] from the operand pack of the last forward (vis_utils.py:107-121)."""
    b, n, dk, dv = ws.shape
    idx = rows.to(device=ws._buf.device, dtype=torch.int32).contiguous()
    out = torch.empty((b, idx.numel(), n), device=ws._buf.device, dtype=torch.float32)
    with torch.cuda.device(ws._buf.device):
        _lib.check(_lib.load().acan_attn_rows(ws.ptr, _lib.ptr(ws.feat_pack), idx.data_ptr(), idx.numel(), out.data_ptr(), b, n, dk, dv, ws.scale, _lib.current_stream()),
                   "acan_attn_rows")
    return out


def attention_loss_fused(ws: AttnWorkspace, dis_label: torch.Tensor) -> torch.Tensor:
    b, n, dk, dv = ws.shape
    lab = _cuda_f32(dis_label, "dis_label")
    hh, ww = lab.shape[-2:]
    loss = torch.empty((), device=lab.device, dtype=torch.float32)
    with torch.cuda.device(lab.device):
        _lib.check(_lib.load().acan_attention_loss_fused(ws.ptr, _lib.ptr(ws.feat_pack), lab.data_ptr(), loss.data_ptr(), b, hh, ww, dk, dv, ws.scale, _lib.current_stream()),
                   "acan_attention_loss_fused")
    return loss


# ------------------------------------------------------------------------------------ attention loss (materialised)

def attention_loss(sim_map: torch.Tensor, dis_label: torch.Tensor, want_grad: bool = False, grad_scale: float = 1.0):
    q = _cuda_f32(sim_map, "sim_map")
    lab = _cuda_f32(dis_label, "dis_label")
    b, n, _ = q.shape
    hh, ww = lab.shape[-2:]
    if (hh // 8) * (ww // 8) != n:
        raise _lib.AcanLibraryError(f"label {tuple(lab.shape)} does not match sim_map {tuple(q.shape)}")
    loss = torch.empty((), device=q.device, dtype=torch.float32)
    ws = torch.empty(2 * b * n, device=q.device, dtype=torch.float32)
    grad = torch.empty_like(q) if want_grad else None
    with torch.cuda.device(q.device):
        _lib.check(_lib.load().acan_attention_loss(q.data_ptr(), lab.data_ptr(), loss.data_ptr(), ws.data_ptr(), _lib.ptr(grad), float(grad_scale),
                                                   b, hh, ww, _lib.current_stream()), "acan_attention_loss")
    return (loss, grad) if want_grad else loss


def gt_sim_map(dis_label: torch.Tensor) -> torch.Tensor:
    lab = _cuda_f32(dis_label, "dis_label")
    b = lab.shape[0]
    hh, ww = lab.shape[-2:]
    n = (hh // 8) * (ww // 8)
    gt = torch.empty((b, n, n), device=lab.device, dtype=torch.float32)
    ws = torch.empty(b * n, device=lab.device, dtype=torch.float32)
    with torch.cuda.device(lab.device):
        _lib.check(_lib.load().acan_gt_sim_map(lab.data_ptr(), gt.data_ptr(), ws.data_ptr(), b, hh, ww, _lib.current_stream()), "acan_gt_sim_map")
    return gt


def cast_to_f32_multi(dst, src) -> bool:
    """dst[i] (fp32) <- src[i] (fp16 / bf16), all pairs in ONE launch per 96 tensors (acan_cast_to_f32_multi).  Every pair must be dense
    and share one memory order (equal strides); returns False -- nothing done -- when a pair does not qualify, so that the caller can
    fall back to the framework's copies."""
    import ctypes
    if not src:
        return True
    dt = src[0].dtype
    if dt not in (torch.float16, torch.bfloat16):
        return False
    for d, s_ in zip(dst, src):
        same_order = all(a == b_ for a, b_, sz in zip(s_.stride(), d.stride(), s_.shape) if sz > 1)     # strides of size-1 dims are arbitrary
        if not (s_.is_cuda and d.is_cuda and s_.dtype == dt and d.dtype == torch.float32 and s_.shape == d.shape and same_order
                and s_.device == d.device and _is_dense(s_)):
            return False
    n = len(src)
    sp = (ctypes.c_void_p * n)(*[t.data_ptr() for t in src])
    dp = (ctypes.c_void_p * n)(*[t.data_ptr() for t in dst])
    cn = (ctypes.c_longlong * n)(*[t.numel() for t in src])
    with torch.cuda.device(src[0].device):
        _lib.check(_lib.load().acan_cast_to_f32_multi(sp, dp, cn, n, _DT[dt], _lib.current_stream()), "acan_cast_to_f32_multi")
    return True


def _is_dense(t: torch.Tensor) -> bool:
    """every element of the storage span is an element of the tensor (some permutation of a contiguous tensor)"""
    if t.numel() == 0:
        return True
    span = 1 + sum((sz - 1) * st for sz, st in zip(t.shape, t.stride()))
    return span == t.numel() and all(st > 0 or sz == 1 for sz, st in zip(t.shape, t.stride()))
