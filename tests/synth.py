"""Deterministic synthetic inputs / weights shared by the golden generator and the tests.

Weights are filled per state_dict key from a generator seeded by a hash of the key, so the
reference model (in ``tests/golden/make_golden.py``) and the drop-in model (in the GPU tests)
receive bit-identical parameters without shipping a 200 MB checkpoint.  Protocol follows
SURVEY.md §8d (seed 2019, images U[0,1), labels with ~10 % zeros, BN never run in eval mode
on raw statistics)."""
from __future__ import annotations

import zlib
from typing import Dict

import torch

SEED = 2019


def _gen(tag: str, seed: int = SEED) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(tag.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def randn(tag: str, *shape, seed: int = SEED) -> torch.Tensor:
    return torch.randn(*shape, generator=_gen(tag, seed), dtype=torch.float32)


def rand(tag: str, *shape, seed: int = SEED) -> torch.Tensor:
    return torch.rand(*shape, generator=_gen(tag, seed), dtype=torch.float32)


def fill_state_dict(sd: Dict[str, torch.Tensor], seed: int = SEED) -> Dict[str, torch.Tensor]:
    """Same-shaped replacement for every entry of ``sd``: He-normal conv/linear weights, small biases,
    BN gamma ~ 1, beta ~ 0, running_mean ~ 0, running_var ~ 1 (all perturbed so nothing is degenerate)."""
    out = {}
    for key, v in sd.items():
        shp = tuple(v.shape)
        k = key.replace("f_query.", "f_key.")             # f_query IS f_key (sadecoder.py:30): one tensor, two names
        if k.endswith("num_batches_tracked"):
            out[key] = torch.zeros(shp, dtype=v.dtype)
        elif k.endswith("running_mean"):
            out[key] = 0.1 * randn(k, *shp, seed=seed)
        elif k.endswith("running_var"):
            out[key] = 0.5 + rand(k, *shp, seed=seed)
        elif v.dim() == 1 and k.endswith("weight"):       # BN gamma
            out[key] = 0.75 + 0.5 * rand(k, *shp, seed=seed)
        elif k.endswith("bias"):
            out[key] = 0.05 * randn(k, *shp, seed=seed)
        else:                                            # conv / linear weight
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            out[key] = randn(k, *shp, seed=seed) * (2.0 / fan_in) ** 0.5
    return out


def feature_map(tag: str, b: int, c: int, h: int, w: int, seed: int = SEED) -> torch.Tensor:
    """relu(N(0,1)) stride-8 backbone features fed directly to the CAM."""
    return torch.relu(randn(tag, b, c, h, w, seed=seed))


def depth_label(tag: str, b: int, h: int, w: int, lo: float, hi: float, seed: int = SEED) -> torch.Tensor:
    """metres in U[lo, hi) with ~10 % of pixels set to 0 (out of range -> bin 0)."""
    d = lo + (hi - lo) * rand(tag + ".d", b, 1, h, w, seed=seed)
    d[rand(tag + ".z", b, 1, h, w, seed=seed) < 0.1] = 0.0
    return d


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max-norm relative error max|a-b| / max|b| (element-wise relative is meaningless near 0)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


class ToyDepthNet(torch.nn.Module):
    """Deterministic stand-in for the network in the test-time-augmentation tests: its output depends on the absolute position
    inside the tile it is given and is not flip-symmetric, so a tiling, flipping or re-assembly mistake changes the result.
    forward(x [N,3,h,w]) -> scores [N,classes,h,w];  inference(scores) -> depth [N,1,h,w]  (the two calls the reference's
    predict_* helpers make, utils/eval_utils.py:112-118, 156-162)."""

    def __init__(self, classes: int = 4):
        super().__init__()
        self.conv = torch.nn.Conv2d(3, classes, 3, padding=1)
        with torch.no_grad():
            self.conv.weight.copy_(0.3 * randn("toy.w", classes, 3, 3, 3))
            self.conv.bias.copy_(0.1 * randn("toy.b", classes))

    def forward(self, x):
        h, w = x.shape[2:]
        ramp = 0.1 * torch.linspace(0, 1, w, device=x.device).view(1, 1, 1, w) + 0.05 * torch.linspace(0, 1, h, device=x.device).view(1, 1, h, 1) ** 2
        return self.conv(x) + ramp

    def inference(self, y):
        return y.abs().sum(dim=1, keepdim=True) + 0.5
