"""Structure-level check of ``acan_b200.network.install`` on a network built by the REFERENCE's own factories.
Needs the reference checkout (``/root/reference``: present in the build container, absent on the GPU box -> skipped there); no CUDA
call is made: the forward pass itself is covered on the GPU by the golden-based network tests."""
import importlib
import os
import sys
import types

import pytest
import torch

REF = "/root/reference/code"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference checkout not available")


def _reference():
    sys.path.insert(0, REF)
    try:
        ref_models = importlib.import_module("models")
        ref_params = importlib.import_module("creator.params") if os.path.isfile(os.path.join(REF, "creator", "params.py")) else None
    finally:
        sys.path.remove(REF)
    return ref_models, ref_params


def _args(**kw):
    base = dict(encoder="resnet50", min_depth=0.72, max_depth=10.0, classes=80, classifier="OR", inference="soft", decoder="attention",
                height=64, width=96, alpha=0.5, beta=1.0, use_weights=False, dataset="nyu", eps=0.0, prior="uniform", optimizer="sgd", lr=1e-4)
    base.update(kw)
    return types.SimpleNamespace(**base)


def test_install_on_a_reference_built_network():
    ref_models, _ = _reference()
    from acan_b200 import network as b200
    from acan_b200.modules import AttentionLoss2d, SADecoder
    args = _args()
    net = ref_models.create_network(args)
    criterion = ref_models.create_lossfunc(args, net)
    before_keys = list(net.state_dict().keys())
    before_sd = {k: v.clone() for k, v in net.state_dict().items()}
    before_children = [n for n, _ in net.named_children()]
    before_params = [n for n, _ in net.named_parameters()]
    out = b200.install(net, criterion)
    assert out is net
    # same checkpoint surface: keys, order, values (incl. the aliased saconv.f_query.* duplicates), children and parameter names
    assert list(net.state_dict().keys()) == before_keys
    assert all(torch.equal(net.state_dict()[k], v) for k, v in before_sd.items())
    assert [n for n, _ in net.named_children()] == before_children and [n for n, _ in net.named_parameters()] == before_params
    assert isinstance(net.decoder, SADecoder) and net.decoder.saconv.f_query is net.decoder.saconv.f_key
    # hot-path entry points rebound to the B200 implementations, criterion losses swapped for the drop-ins
    for name in ("decode_ord", "inference", "soft_ordinal_regression", "hard_ordinal_regression", "forward", "encode", "_inter_logits"):
        assert getattr(net, name).__func__ is getattr(b200.ResNet if hasattr(b200.ResNet, name) else b200.BaseClassificationModel_, name), name
    assert isinstance(criterion.AttentionLoss, AttentionLoss2d)
    assert isinstance(criterion.AppearanceLoss, b200.OrdinalRegression2d) and isinstance(criterion.AuxiliaryLoss, b200.OrdinalRegression2d)
    assert net.lazy_head is True


def test_optimizer_groups_unchanged_by_install():
    ref_models, ref_params = _reference()
    if ref_params is None:
        pytest.skip("creator/params.py not in this checkout")
    from acan_b200 import network as b200
    args = _args(alpha=0.0)
    net = ref_models.create_network(args)
    groups_before = [[id(p) for p in g["params"]] for g in ref_params.create_params(args, net)]
    sizes_before = [len(g) for g in groups_before]
    b200.install(net, ref_models.create_lossfunc(args, net))
    groups_after = ref_params.create_params(args, net)
    sizes_after = [len(list(g["params"])) for g in groups_after]
    assert sizes_after == sizes_before                       # same three groups (lr, 10 lr, 20 lr) with the same number of tensors
    assert [g["lr"] for g in groups_after] == [args.lr, args.lr * 10, args.lr * 20]
