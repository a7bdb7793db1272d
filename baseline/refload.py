"""Import the UNMODIFIED reference (miraiaroha/ACAN) from ``baseline/_ref/code`` (what ``baseline/fetch_ref.sh`` copied; travels to the
GPU box) or, failing that, from the read-only checkout ``/root/reference/code``.  Test / benchmark infrastructure: nothing under
``acan_b200/`` imports this.

Two out-of-tree shims, both required by today's PyTorch and neither touching the reference's files:
  1. ``models.model.continuous2discrete`` computes ``1 - bool*bool`` (model.py:20), which raises on torch >= 1.2; it is replaced by
     the torch-0.4 meaning, mask = not (d_min < depth < d_max).
  2. on a host without CUDA, ``Tensor.cuda`` is made a no-op (model.py:54 and losses.py hard-code ``.cuda()``).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.path.join(HERE, "_ref", "code"), "/root/reference/code")


def reference_root():
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "models", "model.py")):
            return c
    return None


_cache = None


def load_reference():
    """-> namespace(models, model, losses, sadecoder, creator, root).  Raises RuntimeError when no copy of the reference is present."""
    global _cache
    if _cache is not None:
        return _cache
    root = reference_root()
    if root is None:
        raise RuntimeError("reference not found: run baseline/fetch_ref.sh where /root/reference exists")
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self                     # shim 2
    for name in ("models", "creator"):
        if name in sys.modules and not getattr(sys.modules[name], "__file__", "").startswith(root):
            raise RuntimeError(f"a different top-level package named {name!r} is already imported")
    sys.path.insert(0, root)
    try:
        models = importlib.import_module("models")
        model = importlib.import_module("models.model")
        losses = importlib.import_module("models.losses")
        sadecoder = importlib.import_module("models.sadecoder")
        creator = importlib.import_module("creator") if os.path.isfile(os.path.join(root, "creator", "params.py")) else None
    finally:
        sys.path.remove(root)

    def continuous2discrete(depth, d_min, d_max, n_c):                      # shim 1
        mask = ~((depth > d_min) & (depth < d_max))
        depth = torch.round(torch.log(depth / d_min) / np.log(d_max / d_min) * (n_c - 1))
        depth[mask] = 0
        return depth

    model.continuous2discrete = continuous2discrete
    _cache = types.SimpleNamespace(models=models, model=model, losses=losses, sadecoder=sadecoder, creator=creator, root=root)
    return _cache


def reference_args(**kw):
    """the argparse namespace the reference's factories read (config/__init__.py defaults that matter on the hot path)."""
    base = dict(encoder="resnet101", min_depth=0.72, max_depth=10.0, classes=80, classifier="OR", inference="soft", decoder="attention",
                height=256, width=352, alpha=0.0, beta=0.0, use_weights=False, dataset="nyu", eps=0.0, prior="uniform",
                optimizer="sgd", lr=1e-4, momentum=0.9, weight_decay=5e-4)
    base.update(kw)
    return types.SimpleNamespace(**base)
