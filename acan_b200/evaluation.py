"""Evaluation helpers of the reference's validation / test loop (utils/eval_utils.py), same names and
arguments: ``compute_errors`` as one fused reduction kernel, and the test-time-augmentation drivers with every forward pass of
equal input size batched into ONE call of the network (the reference runs the image, its mirror image, and each tile of the
1.25x view one after the other: 10 forwards per test batch for scales = [1, 1.25] with flips; here it is one forward on a
10x larger batch).  Eval-mode only: batching changes nothing when BatchNorm uses running statistics.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import _lib

measure_list = ["a1", "a2", "a3", "rmse", "rmse_log", "log10", "abs_rel", "sq_rel"]        # utils/eval_utils.py:14


def compute_errors(gt: torch.Tensor, pred: torch.Tensor) -> Dict[str, torch.Tensor]:
    """utils/eval_utils.py:16-48: dict of the eight measures (0-dim CUDA tensors), each the mean over the pixels with gt > 0 of
    the whole batch times the batch size.  One pass over gt / pred, no masked copies, no host synchronisation."""
    if not (gt.is_cuda and pred.is_cuda):
        raise _lib.AcanLibraryError("compute_errors needs CUDA tensors: acan_b200 has no CPU path")
    if gt.numel() != pred.numel():
        raise _lib.AcanLibraryError(f"gt {tuple(gt.shape)} and pred {tuple(pred.shape)} differ in size")
    g = gt.detach().float().contiguous()
    p = pred.detach().float().contiguous()
    lib = _lib.load()
    out = torch.empty(8, device=g.device, dtype=torch.float32)
    ws = torch.empty(int(lib.acan_depth_metrics_ws_floats()), device=g.device, dtype=torch.float32)
    with torch.cuda.device(g.device):
        _lib.check(lib.acan_depth_metrics(g.data_ptr(), p.data_ptr(), out.data_ptr(), ws.data_ptr(), g.numel(), int(pred.shape[0]), _lib.current_stream()),
                   "acan_depth_metrics")
    return dict(zip(measure_list, out.unbind(0)))


def pad_image(image: torch.Tensor, target_size) -> torch.Tensor:
    """utils/eval_utils.py:51-66."""
    return F.pad(image, (0, target_size[1] - image.shape[3], 0, target_size[0] - image.shape[2]), "constant")


def _rescale(image: torch.Tensor, scale: float) -> torch.Tensor:
    return image if scale == 1 else F.interpolate(image, scale_factor=scale, mode="bilinear", align_corners=True)


class _Job:
    """one forward pass the reference would run on its own: an input batch and where its depth goes."""
    __slots__ = ("x", "scale", "view", "box", "crop")

    def __init__(self, x, scale, view, box=None, crop=None):
        self.x, self.scale, self.view, self.box, self.crop = x, scale, view, box, crop


def _tile_boxes(hh: int, ww: int, th: int, tw: int) -> List[Tuple[int, int, int, int]]:
    rows = int(math.ceil((hh - th) / th) + 1)
    cols = int(math.ceil((ww - tw) / tw) + 1)
    boxes = []
    for r in range(rows):
        for c in range(cols):
            x2, y2 = min(c * tw + tw, ww), min(r * th + th, hh)
            boxes.append((max(y2 - th, 0), y2, max(x2 - tw, 0), x2))
    return boxes


def _run_batched(net, jobs: Sequence[_Job], max_batch: int) -> List[torch.Tensor]:
    """depth of every job; jobs with the same input size share forward passes of at most ``max_batch`` images."""
    out: List[torch.Tensor] = [None] * len(jobs)
    by_shape: Dict[Tuple[int, ...], List[int]] = {}
    for i, j in enumerate(jobs):
        by_shape.setdefault(tuple(j.x.shape[1:]), []).append(i)
    for idxs in by_shape.values():
        n = jobs[idxs[0]].x.shape[0]
        per_call = max(1, max_batch // n)
        for k in range(0, len(idxs), per_call):
            chunk = idxs[k:k + per_call]
            pred = net(torch.cat([jobs[i].x for i in chunk], dim=0))
            depth = net.inference(pred[-1] if isinstance(pred, list) else pred)
            for i, d in zip(chunk, depth.split(n, dim=0)):
                out[i] = d
    return out


@torch.no_grad()
def predict_multi_scale(net, image: torch.Tensor, scales, classes: int, flip: bool, max_batch: int = 64) -> torch.Tensor:
    """utils/eval_utils.py:178-209 (with predict_whole_img :167-176 and predict_sliding :160-165 inlined): mean over ``scales`` of the
    depth predicted on the rescaled image -- whole for scale <= 1, in tiles of the original size for scale > 1 -- each optionally
    averaged with the prediction on the mirrored image.  Same result as the reference's loops; all equally-sized inputs go through
    the network together."""
    if getattr(net, "training", False):
        raise RuntimeError("predict_multi_scale batches forward passes: the network must be in eval() mode")
    n, _, hh0, ww0 = image.shape
    views = [False, True] if flip else [False]
    jobs: List[_Job] = []
    for s_ in scales:
        s_ = float(s_)
        for mirrored in views:
            big = _rescale(image.flip([3]) if mirrored else image, s_)
            if s_ <= 1.0:
                jobs.append(_Job(big, s_, mirrored))
            else:
                for (y1, y2, x1, x2) in _tile_boxes(big.shape[2], big.shape[3], hh0, ww0):
                    tile = big[:, :, y1:y2, x1:x2]
                    jobs.append(_Job(pad_image(tile, (hh0, ww0)), s_, mirrored, (y1, y2, x1, x2, big.shape[2], big.shape[3]), tile.shape[2:]))
    depths = _run_batched(net, jobs, max_batch)
    total = torch.zeros((n, 1, hh0, ww0), device=image.device, dtype=torch.float32)
    for s_ in scales:
        s_ = float(s_)
        per_view = []
        for mirrored in views:
            mine = [(j, d) for j, d in zip(jobs, depths) if j.scale == s_ and j.view == mirrored]
            if s_ <= 1.0:
                d = F.interpolate(mine[0][1].float() * s_, size=(hh0, ww0), mode="nearest")
            else:
                hh, ww = mine[0][0].box[4], mine[0][0].box[5]
                acc = torch.zeros((n, 1, hh, ww), device=image.device, dtype=torch.float32)
                cnt = torch.zeros_like(acc)
                for j, dj in mine:
                    y1, y2, x1, x2 = j.box[:4]
                    dj = F.interpolate(dj.float() * s_, size=(hh0, ww0), mode="nearest")[:, :, :j.crop[0], :j.crop[1]]
                    acc[:, :, y1:y2, x1:x2] += dj
                    cnt[:, :, y1:y2, x1:x2] += 1
                d = F.interpolate(acc / cnt, scale_factor=(1.0 / s_, 1.0 / s_), mode="nearest")
            per_view.append(d.flip([3]) if mirrored else d)
        total += per_view[0] if not flip else 0.5 * (per_view[0] + per_view[1])
    return total / len(scales)


@torch.no_grad()
def predict_whole_img(net, image: torch.Tensor, scale: float = 1, flip: bool = False) -> torch.Tensor:
    """utils/eval_utils.py:167-176; the image and its mirror image share one forward pass."""
    if float(scale) > 1.0:
        d = [F.interpolate(x.float() * scale, size=tuple(image.shape[2:]), mode="nearest")
             for x in _run_batched(net, [_Job(_rescale(v, scale), scale, i) for i, v in enumerate([image, image.flip([3])] if flip else [image])], 64)]
        return 0.5 * (d[0] + d[1].flip([3])) if flip else d[0]
    return predict_multi_scale(net, image, [scale], 0, flip)


@torch.no_grad()
def predict_sliding(net, image: torch.Tensor, tile_size, classes: int, scale: float = 1, flip: bool = False) -> torch.Tensor:
    """utils/eval_utils.py:160-165 for tiles of the image's own size (what predict_multi_scale passes)."""
    if tuple(tile_size) != tuple(image.shape[2:]) or float(scale) <= 1.0:
        raise NotImplementedError("predict_sliding mirrors the reference's only call pattern: tile_size == image size, scale > 1")
    return predict_multi_scale(net, image, [scale], classes, flip)
