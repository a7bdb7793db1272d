"""CPU oracle for the ACAN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  Nothing under
``acan_b200/`` imports it; the product path fails loudly when the CUDA library is
missing instead of falling back to anything in here.

What it is: a functional restatement (plain ``torch`` CPU ops, fp32 or fp64, no
``nn.Module``) of the algorithm the reference implements in
``code/models/sadecoder.py``, ``code/models/losses.py:12-62,103-158``,
``code/models/model.py:19-99,113-126,216-272`` and (row f-4 of the scope table) ``code/utils/eval_utils.py:16-209``.  Every function cites the reference
lines it follows.  The reference is Python over ATen, so the oracle is Python over
ATen too; it was written from the maths, not transcribed.

Parity pin: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, run in the build
container by ``tests/golden/make_golden.py`` (imports ``/root/reference/code/models``
unmodified, with the two out-of-tree shims of SURVEY.md §8c) and committed as
``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks every function here
against those files, plus the closed-form known-answer cases of SURVEY.md §8c.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# depth <-> bin quantisation                                          (model.py:19-27)
# --------------------------------------------------------------------------------------


def continuous2discrete(depth: Tensor, d_min: float, d_max: float, n_c: int) -> Tensor:
    """metres -> bin index (as float), 0 for depths outside (d_min, d_max).

    model.py:19-23.  The reference line ``1 - bool*bool`` only runs on torch<1.2; the
    semantics kept here are the torch-0.4 ones (SURVEY.md §8c shim 1): the mask is the
    complement of ``d_min < depth < d_max``.  The log-range constant is a Python float
    (float64) folded into the tensor dtype, as ``np.log`` does in the reference.
    """
    outside = ~((depth > d_min) & (depth < d_max))
    k = torch.round(torch.log(depth / d_min) / math.log(d_max / d_min) * (n_c - 1))
    k = k.clone()
    k[outside] = 0
    return k


def discrete2continuous(k: Tensor, d_min: float, d_max: float, n_c: int) -> Tensor:
    """bin index -> metres, log-space bins (model.py:25-27)."""
    return torch.exp(k / (n_c - 1) * math.log(d_max / d_min) + math.log(d_min))


# --------------------------------------------------------------------------------------
# ordinal / cross-entropy depth head                                  (model.py:40-99)
# --------------------------------------------------------------------------------------


def decode_ord(y: Tensor) -> Tensor:
    """[B,2K,H,W] pair logits -> [B,K,H,W] P(bin > k); unstabilised, as model.py:40-45."""
    b, c2, h, w = y.shape
    pair = y.reshape(b, c2 // 2, 2, h, w)
    e = torch.exp(pair)
    return e[:, :, 1] / (e[:, :, 0] + e[:, :, 1])


def hard_ordinal_index(prob: Tensor) -> Tensor:
    """count of bins with p > 0.5 -> int64 [B,1,H,W] (model.py:62-63)."""
    return (prob > 0.5).sum(dim=1, keepdim=True)


def hard_ordinal_regression(prob: Tensor, d_min: float, d_max: float, n_c: int) -> Tensor:
    """model.py:61-67: midpoint of bins ``label`` and ``label+1``."""
    lab = hard_ordinal_index(prob).to(prob.dtype)
    return 0.5 * (discrete2continuous(lab, d_min, d_max, n_c) + discrete2continuous(lab + 1, d_min, d_max, n_c))


def soft_ordinal_regression(prob: Tensor, d_min: float, d_max: float, n_c: int) -> Tensor:
    """model.py:69-78: linear interpolation between neighbouring bin midpoints at s = sum_k p_k."""
    s = prob.sum(dim=1, keepdim=True)
    i = torch.floor(s)
    f = s - i
    d0 = discrete2continuous(i, d_min, d_max, n_c)
    d1 = discrete2continuous(i + 1, d_min, d_max, n_c)
    d2 = discrete2continuous(i + 2, d_min, d_max, n_c)
    return 0.5 * (d0 + d1) * (1 - f) + 0.5 * (d1 + d2) * f


def hard_cross_entropy_index(score: Tensor) -> Tensor:
    """argmax over raw scores, first index on ties (model.py:48)."""
    return torch.argmax(score, dim=1, keepdim=True)


def hard_cross_entropy(score: Tensor, d_min: float, d_max: float, n_c: int) -> Tensor:
    """model.py:47-50."""
    return discrete2continuous(hard_cross_entropy_index(score).to(score.dtype), d_min, d_max, n_c)


def soft_cross_entropy(score: Tensor, d_min: float, d_max: float, n_c: int) -> Tensor:
    """model.py:52-59: exp of the softmax-expected log-depth."""
    p = torch.softmax(score, dim=1)
    w = torch.arange(n_c, dtype=score.dtype) * math.log(d_max / d_min) / (n_c - 1) + math.log(d_min)
    return torch.exp((p * w.view(1, n_c, 1, 1)).sum(dim=1, keepdim=True))


def inference(y: Tensor, classifier: str, mode: str, d_min: float, d_max: float, n_c: int) -> Tensor:
    """dispatcher of model.py:80-99 (``y`` already a tensor)."""
    if classifier == "OR":
        fn = soft_ordinal_regression if mode == "soft" else hard_ordinal_regression
    else:
        fn = soft_cross_entropy if mode == "soft" else hard_cross_entropy
    return fn(y, d_min, d_max, n_c)


# --------------------------------------------------------------------------------------
# context aggregation module                                     (sadecoder.py:20-122)
# --------------------------------------------------------------------------------------


def batchnorm_relu(
    z: Tensor,
    weight: Tensor,
    bias: Tensor,
    running_mean: Tensor,
    running_var: Tensor,
    training: bool,
    eps: float = 1e-5,
) -> Tuple[Tensor, Optional[Tuple[Tensor, Tensor]]]:
    """BatchNorm2d + ReLU of f_key (sadecoder.py:27-28).  Returns (out, batch moments or None)."""
    if training:
        mu = z.mean(dim=(0, 2, 3))
        var_b = z.var(dim=(0, 2, 3), unbiased=False)
        out = (z - mu.view(1, -1, 1, 1)) / torch.sqrt(var_b.view(1, -1, 1, 1) + eps)
        n = z.numel() // z.shape[1]
        moments = (mu, var_b * n / max(n - 1, 1))
    else:
        out = (z - running_mean.view(1, -1, 1, 1)) / torch.sqrt(running_var.view(1, -1, 1, 1) + eps)
        moments = None
    out = out * weight.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    return torch.relu(out), moments


def bn_running_update_twice(
    running_mean: Tensor, running_var: Tensor, moments: Tuple[Tensor, Tensor], momentum: float = 0.1
) -> Tuple[Tensor, Tensor]:
    """f_query is f_key and is *called twice* per forward (sadecoder.py:30,44-45): in train
    mode the running stats take two momentum steps with the same batch moments."""
    mu, var_u = moments
    rm, rv = running_mean, running_var
    for _ in range(2):
        rm = (1 - momentum) * rm + momentum * mu
        rv = (1 - momentum) * rv + momentum * var_u
    return rm, rv


def key_projection(x: Tensor, sd: Dict[str, Tensor], prefix: str, training: bool) -> Tuple[Tensor, Optional[Tuple[Tensor, Tensor]]]:
    """F = ReLU(BN(conv1x1(x)))  [B,dk,h,w]   (sadecoder.py:25-28)."""
    z = F.conv2d(x, sd[prefix + "conv.weight"], sd[prefix + "conv.bias"])
    return batchnorm_relu(
        z, sd[prefix + "bn.weight"], sd[prefix + "bn.bias"], sd[prefix + "bn.running_mean"], sd[prefix + "bn.running_var"], training
    )


def affinity(feat: Tensor) -> Tensor:
    """sim_map = softmax_rows(F^T F / sqrt(dk))   [B,N,N]   (sadecoder.py:42-49; Q = K = F)."""
    b, dk = feat.shape[:2]
    f = feat.reshape(b, dk, -1)
    s = torch.matmul(f.transpose(1, 2), f) * (dk ** -0.5)
    return torch.softmax(s, dim=-1)


def aggregate(sim_map: Tensor, value: Tensor) -> Tensor:
    """context[b,c,i] = sum_j sim_map[b,i,j] value[b,c,j], reshaped NCHW (sadecoder.py:85-87)."""
    b, dv, h, w = value.shape
    v = value.reshape(b, dv, -1)
    return torch.matmul(v, sim_map.transpose(1, 2)).reshape(b, dv, h, w)


def value_projection(x: Tensor, sd: Dict[str, Tensor], prefix: str) -> Tensor:
    """V = ReLU(conv1x1(ReLU(conv3x3(x))))   (sadecoder.py:63-67)."""
    v = torch.relu(F.conv2d(x, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"], padding=1))
    return torch.relu(F.conv2d(v, sd[prefix + "conv2.weight"], sd[prefix + "conv2.bias"]))


def self_attention_block(x: Tensor, sd: Dict[str, Tensor], prefix: str = "", training: bool = False):
    """SelfAttentionBlock_.forward with scale=1 (sadecoder.py:80-90) -> (context, sim_map, bn moments)."""
    feat, moments = key_projection(x, sd, prefix + "f_key.", training)
    sim = affinity(feat)
    val = value_projection(x, sd, prefix + "f_value.")
    return aggregate(sim, val), sim, moments


def image_context(x: Tensor, sd: Dict[str, Tensor], prefix: str, keep_mask: Optional[Tensor] = None) -> Tensor:
    """global mean over (h,w) -> [Dropout2d] -> Linear+ReLU -> Linear+ReLU  -> [B,512]
    (sadecoder.py:97-106; the bilinear 'upsample' of a 1x1 map is a broadcast).
    ``keep_mask`` [B,C] in {0,1} is the Dropout2d(0.5) keep decision (scaled by 2 when given)."""
    g = x.mean(dim=(2, 3))
    if keep_mask is not None:
        g = g * keep_mask * 2.0
    g = torch.relu(F.linear(g, sd[prefix + "linear1.weight"], sd[prefix + "linear1.bias"]))
    return torch.relu(F.linear(g, sd[prefix + "linear2.weight"], sd[prefix + "linear2.bias"]))


def merge(context: Tensor, g: Tensor, sd: Dict[str, Tensor], prefix: str,
          keep_in: Optional[Tensor] = None, keep_out: Optional[Tensor] = None) -> Tensor:
    """ReLU(conv1x1(cat(context, broadcast(g)))) with the two optional Dropout2d(0.5) masks
    (sadecoder.py:107-111,120-121).  ``keep_in`` [B,dv+512], ``keep_out`` [B,dv]."""
    b, dv, h, w = context.shape
    cat = torch.cat([context, g.view(b, -1, 1, 1).expand(b, g.shape[1], h, w)], dim=1)
    if keep_in is not None:
        cat = cat * (keep_in * 2.0).view(b, -1, 1, 1)
    out = torch.relu(F.conv2d(cat, sd[prefix + "conv1.weight"], sd[prefix + "conv1.bias"]))
    if keep_out is not None:
        out = out * (keep_out * 2.0).view(b, -1, 1, 1)
    return out


def sadecoder_forward(x: Tensor, sd: Dict[str, Tensor], training: bool = False):
    """SADecoder.forward (sadecoder.py:116-122), dropout disabled -> (out, sim_map, bn moments)."""
    context, sim, moments = self_attention_block(x, sd, "saconv.", training)
    g = image_context(x, sd, "image_context.")
    return merge(context, g, sd, "merge."), sim, moments


# --------------------------------------------------------------------------------------
# attention loss                                                      (losses.py:12-62)
# --------------------------------------------------------------------------------------


def nearest_downsample8(label: Tensor) -> Tensor:
    """F.interpolate(mode='nearest') to (H//8, W//8) picks pixel (8i, 8j) (losses.py:50)."""
    h, w = label.shape[-2:]
    return label[..., : (h // 8) * 8 : 8, : (w // 8) * 8 : 8]


def gt_affinity(bins: Tensor) -> Tensor:
    """ground-truth affinity from stride-8 bin indices ``bins`` [B,N] (float).

    losses.py:40-46 computes softmax_j(-|log a_i - log a_j|) and zeroes NaNs.  In closed
    form: W_ij = min(a_i/a_j, a_j/a_i) / sum_j(...), rows and columns with
    a = 0 all zero."""
    a = bins
    ai, aj = a.unsqueeze(2), a.unsqueeze(1)
    lo, hi = torch.minimum(ai, aj), torch.maximum(ai, aj)
    r = torch.where(lo > 0, lo / hi.clamp_min(1e-30), torch.zeros_like(lo))
    z = r.sum(dim=2, keepdim=True)
    return torch.where(z > 0, r / z.clamp_min(1e-30), torch.zeros_like(r))


def gt_affinity_literal(bins: Tensor) -> Tensor:
    """the same thing evaluated literally as the reference does (log / abs / softmax / NaN->0)."""
    la = torch.log(bins)
    w = torch.softmax(-torch.abs(la.unsqueeze(2) - la.unsqueeze(1)), dim=-1)
    return torch.where(torch.isnan(w), torch.zeros_like(w), w)


def kl_attention(sim_map: Tensor, gt: Tensor) -> Tensor:
    """mean over (b, i) of sum_j P log(clamp(P/Q, 1e-8, 1e8)); P = gt, Q = sim_map (losses.py:12,18-33)."""
    ratio = torch.clamp(gt / sim_map, 1e-8, 1e8)
    return (gt * torch.log(ratio)).sum(dim=-1).mean()


def attention_loss(sim_map: Tensor, dis_label: Tensor, literal: bool = False) -> Tensor:
    """AttentionLoss2d.forward (losses.py:54-62); ``dis_label`` [B,1,H,W] holds bin indices
    (model.py:260,270), not metres."""
    small = nearest_downsample8(dis_label)
    bins = small.reshape(small.shape[0], -1).to(sim_map.dtype)
    gt = gt_affinity_literal(bins) if literal else gt_affinity(bins)
    return kl_attention(sim_map, gt)


def attention_loss_grad_sim(sim_map: Tensor, gt: Tensor) -> Tensor:
    """dL/dQ of ``kl_attention``: -P/Q / (B*N) inside the clamp window, 0 outside."""
    ratio = gt / sim_map
    inside = (ratio > 1e-8) & (ratio < 1e8)
    b, n = sim_map.shape[:2]
    return torch.where(inside, -gt / sim_map, torch.zeros_like(gt)) / (b * n)


# --------------------------------------------------------------------------------------
# classifier tail + whole network                          (model.py:113-126,162-233)
# --------------------------------------------------------------------------------------


def classifier_tail(x: Tensor, sd: Dict[str, Tensor], prefix: str = "classifier.") -> Tensor:
    """1x1 conv to 2K (OR) / K (CE) channels, then x8 bilinear, align_corners=True (model.py:123-125)."""
    y = F.conv2d(x, sd[prefix + "conv.weight"], sd[prefix + "conv.bias"])
    return F.interpolate(y, scale_factor=8, mode="bilinear", align_corners=True)


def _bn(x: Tensor, sd: Dict[str, Tensor], p: str, training: bool) -> Tensor:
    return F.batch_norm(x, sd[p + "running_mean"].clone(), sd[p + "running_var"].clone(), sd[p + "weight"], sd[p + "bias"],
                        training=training, momentum=0.0, eps=1e-5)


def _bottleneck(x: Tensor, sd: Dict[str, Tensor], p: str, stride: int, dilation: int, training: bool) -> Tensor:
    """1x1 -> 3x3(dilated) -> 1x1 residual block (model.py:128-160)."""
    out = torch.relu(_bn(F.conv2d(x, sd[p + "conv1.weight"]), sd, p + "bn1.", training))
    out = torch.relu(_bn(F.conv2d(out, sd[p + "conv2.weight"], stride=stride, padding=dilation, dilation=dilation), sd, p + "bn2.", training))
    out = _bn(F.conv2d(out, sd[p + "conv3.weight"]), sd, p + "bn3.", training)
    if (p + "downsample.0.weight") in sd:
        x = _bn(F.conv2d(x, sd[p + "downsample.0.weight"], stride=stride), sd, p + "downsample.1.", training)
    return torch.relu(out + x)


def backbone_forward(x: Tensor, sd: Dict[str, Tensor], layers: Sequence[int], training: bool = False) -> Tensor:
    """dilated ResNet encoder, output stride 8 (model.py:177-180,216-225)."""
    x = torch.relu(_bn(F.conv2d(x, sd["conv1.weight"], stride=2, padding=3), sd, "bn1.", training))
    x = F.max_pool2d(x, 3, stride=2, padding=1)
    for li, (nblk, stride, dil) in enumerate(zip(layers, (1, 2, 1, 1), (1, 1, 2, 4)), start=1):
        for bi in range(nblk):
            x = _bottleneck(x, sd, f"layer{li}.{bi}.", stride if bi == 0 else 1, dil, training)
    return x


def acan_forward(x: Tensor, sd: Dict[str, Tensor], layers: Sequence[int], classifier: str = "OR", training: bool = False):
    """ResNet.forward (model.py:216-233) with alpha = 0 -> dict(y, sim_map); dropout disabled."""
    feat = backbone_forward(x, sd, layers, training)
    dsd = {k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}
    out, sim, _ = sadecoder_forward(feat, dsd, training)
    y = classifier_tail(out, sd)
    if classifier == "OR":
        y = decode_ord(y)
    return {"inter_y": None, "sim_map": sim, "y": y}


def ordinal_regression_loss(prob: Tensor, label: Tensor, ignore_index: int = 0) -> Tensor:
    """OrdinalRegression2d with reduction='sum' (losses.py:103-158): per valid pixel
    -(sum_{k<l} log p_k + sum_{k>=l} log(1-p_k)), mean over pixels with label != ignore_index.
    Only used to restate the reference training step around the hot path."""
    b, c, h, w = prob.shape
    k = torch.arange(c).view(1, c, 1, 1)
    lab = label.view(b, 1, h, w)
    lt = (k < lab).to(prob.dtype)
    ent = -(torch.log(prob.clamp(1e-8, 1e8)) * lt + torch.log((1 - prob).clamp(1e-8, 1e8)) * (1 - lt))
    return ent.sum(dim=1)[label.view(b, h, w) != ignore_index].mean()


def cross_entropy_loss(score: Tensor, label: Tensor, ignore_index: int, eps: float = 0.0, prior: str = "uniform") -> Tensor:
    """CrossEntropy2d with reduction='sum', no class weights (losses.py:103-141, 164-196): per valid pixel
    -sum_k [ s_k log clamp(p_k) + (1 - s_k) log clamp(1 - p_k) ],  p = softmax over classes, s = one-hot smoothed by eps with a uniform
    1/(K-1) or gaussian (sigma = K/10) prior on the other classes; mean over pixels with label != ignore_index."""
    b, c, h, w = score.shape
    p = torch.softmax(score, dim=1)
    k = torch.arange(c, dtype=score.dtype).view(1, c, 1, 1)
    lab = label.view(b, 1, h, w).to(score.dtype)
    hot = (k == lab).to(score.dtype)
    if eps == 0:
        pri = 0.0
    elif prior == "gaussian":
        sig = c / 10.0
        pri = torch.exp(-(k - lab) ** 2 / (2 * sig * sig)) / math.sqrt(2 * math.pi * sig * sig)
    else:
        pri = 1.0 / (c - 1)
    s_ = (1 - eps) * hot + eps * pri * (1 - hot)
    ent = -(s_ * torch.log(p.clamp(1e-8, 1e8)) + (1 - s_) * torch.log((1 - p).clamp(1e-8, 1e8)))
    return ent.sum(dim=1)[label.view(b, h, w) != ignore_index].mean()


# ---------------------------------------------------------------- evaluation helpers

def compute_errors(gt: Tensor, pred: Tensor) -> Dict[str, Tensor]:
    """utils/eval_utils.py:16-48.  Means over the pixels with gt > 0 of the whole batch, each multiplied by the batch size; the
    reference's 'log10' measure uses the same NATURAL log as 'rmse_log' (eval_utils.py:30-31)."""
    n_batch = pred.shape[0]
    keep = gt > 0
    g, p = gt[keep], pred[keep]
    slog = lambda t: torch.log(t.clamp(1e-6, 1e6))        # noqa: E731
    ratio = torch.maximum(g / p, p / g)
    dlog = slog(g) - slog(p)
    vals = [(ratio < 1.25 ** k).to(g.dtype).mean() for k in (1, 2, 3)]
    vals += [((g - p) ** 2).mean().sqrt(), (dlog ** 2).mean().sqrt(), dlog.abs().mean(), ((g - p).abs() / g).mean(), ((g - p) ** 2 / g).mean()]
    return {name: v * n_batch for name, v in zip(("a1", "a2", "a3", "rmse", "rmse_log", "log10", "abs_rel", "sq_rel"), vals)}


def _rescale(image: Tensor, scale: float) -> Tensor:
    return image if scale == 1 else torch.nn.functional.interpolate(image, scale_factor=scale, mode="bilinear", align_corners=True)


def _net_depth(net, x: Tensor) -> Tensor:
    out = net(x)
    return net.inference(out[-1] if isinstance(out, list) else out)


def predict_whole_img(net, image: Tensor, scale: float = 1, flip: bool = False) -> Tensor:
    """utils/eval_utils.py:131-176: one forward on the (rescaled) image, depth * scale, nearest resize back; with flip the
    mirrored image's prediction is mirrored back and averaged in."""
    size = image.shape[2:]

    def once(img):
        return torch.nn.functional.interpolate(_net_depth(net, _rescale(img, scale)) * scale, size=tuple(size), mode="nearest")
    d = once(image)
    return 0.5 * (d + once(image.flip([3])).flip([3])) if flip else d


def predict_sliding(net, image: Tensor, tile_size, classes: int, scale: float = 1, flip: bool = False) -> Tensor:
    """utils/eval_utils.py:68-129, 160-165: non-overlapping tiles of ``tile_size`` over the rescaled image (the last row / column of
    tiles is shifted back inside), per-tile depth * scale, count-normalised, nearest resize by 1/scale."""
    th, tw = tile_size

    def once(img):
        big = _rescale(img, scale)
        n, _, hh, ww = big.shape
        acc = torch.zeros((n, 1, hh, ww), dtype=big.dtype, device=big.device)
        cnt = torch.zeros_like(acc)
        rows = int(math.ceil((hh - th) / th) + 1)
        cols = int(math.ceil((ww - tw) / tw) + 1)
        for r in range(rows):
            for c in range(cols):
                x2, y2 = min(c * tw + tw, ww), min(r * th + th, hh)
                x1, y1 = max(x2 - tw, 0), max(y2 - th, 0)
                tile = big[:, :, y1:y2, x1:x2]
                padded = torch.nn.functional.pad(tile, (0, tw - tile.shape[3], 0, th - tile.shape[2]))
                d = torch.nn.functional.interpolate(_net_depth(net, padded) * scale, size=(th, tw), mode="nearest")
                acc[:, :, y1:y2, x1:x2] += d[:, :, :tile.shape[2], :tile.shape[3]]
                cnt[:, :, y1:y2, x1:x2] += 1
        return torch.nn.functional.interpolate(acc / cnt, scale_factor=(1.0 / scale, 1.0 / scale), mode="nearest")
    d = once(image)
    return 0.5 * (d + once(image.flip([3])).flip([3])) if flip else d


def predict_multi_scale(net, image: Tensor, scales, classes: int, flip: bool) -> Tensor:
    """utils/eval_utils.py:178-209: mean over scales; scales <= 1 see the whole image, larger ones are tiled at the original size."""
    total = torch.zeros((image.shape[0], 1) + tuple(image.shape[2:]), dtype=image.dtype, device=image.device)
    for s_ in scales:
        s_ = float(s_)
        total += predict_whole_img(net, image, s_, flip) if s_ <= 1.0 else predict_sliding(net, image, tuple(image.shape[2:]), classes, s_, flip)
    return total / len(scales)
