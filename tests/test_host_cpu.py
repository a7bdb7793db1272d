"""CPU-side checks (-m "not gpu"): the C-ABI library loads and exports what include/acan_b200.h declares, argument
validation fails without touching a GPU, the drop-in modules keep the reference's checkpoint / optimizer surface,
and the product path refuses to run without CUDA (no fallback)."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "acan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(acan_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from acan_b200 import _lib
    lib = _lib.load()
    declared = _header_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/acan_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature in acan_b200/_lib.py"
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.acan_abi_version() == _lib.ABI_VERSION


def test_argument_validation_needs_no_gpu():
    from acan_b200 import _lib
    lib = _lib.load()
    BAD = 10001
    assert lib.acan_decode_ord_fwd(None, None, 1, 80, 64, None) == BAD
    assert b"null" in lib.acan_last_error()
    assert lib.acan_head_fwd(1, 1, None, 1, 80, 64, 10.0, 0.72, 0, 0, None) == BAD      # d_max < d_min
    assert lib.acan_head_fwd(1, 1, None, 1, 80, 64, 0.72, 10.0, 7, 0, None) == BAD      # unknown kind
    assert lib.acan_head_fwd(1, 1, None, 1, 80, 64, 0.72, 10.0, 2, 1, None) == BAD      # from_logits on a CE head
    assert lib.acan_attn_fwd(1, 1, 1, None, 1, 1024, 1 << 30, 1, 1001, 512, 2048, 0.044, None) == BAD   # N % 8 != 0
    assert lib.acan_attn_fwd(1, 1, 1, None, 1, 1024, 1 << 30, 1, 1408, 500, 2048, 0.044, None) == BAD   # dk % 64 != 0
    assert lib.acan_attn_fwd(1, 1, 1, None, 1, 1024, 16, 1, 1408, 512, 2048, 0.044, None) == 10002      # workspace too small
    assert lib.acan_attention_loss(1, 1, 1, 1, None, 1.0, 1, 4, 4, None) == BAD                         # H, W < 8
    assert lib.acan_attn_workspace_bytes(16, 1408, 512, 2048) > 16 * 1408 * 1408 * 2
    assert lib.acan_attn_workspace_bytes(0, 1408, 512, 2048) == 0


def test_no_cpu_fallback():
    from acan_b200 import _lib, ops
    with pytest.raises(_lib.AcanLibraryError):
        ops.decode_ord(torch.zeros(1, 160, 8, 8))
    with pytest.raises(_lib.AcanLibraryError):
        ops.attention_forward(torch.zeros(1, 512, 4, 6), torch.zeros(1, 2048, 4, 6))
    with pytest.raises(_lib.AcanLibraryError):
        ops.head_inference(torch.zeros(1, 80, 8, 8), "OR", "soft", 0.72, 10.0, 80)


def test_missing_library_is_loud(monkeypatch):
    from acan_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libacan_b200.so")
    with pytest.raises(_lib.AcanLibraryError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_state_dict_and_optimizer_surface_match_reference(golden):
    from acan_b200.modules import SADecoder
    from acan_b200.network import ResNet
    g = golden("sadecoder_24")
    dec = SADecoder(height=32, width=48)
    sd = dec.state_dict()
    assert list(sd.keys()) == list(g["keys"])
    assert [str(tuple(v.shape)) for v in sd.values()] == list(g["shapes"])
    assert [n for n, _ in dec.named_parameters()] == list(g["param_names"])
    assert dec.saconv.f_query is dec.saconv.f_key
    g = golden("resnet_tiny")
    net = ResNet(0.72, 10.0, 80, "OR", "soft", "attention", 64, 96, alpha=0, beta=1.0, layers=[1, 1, 1, 1])
    sd = net.state_dict()
    assert list(sd.keys()) == list(g["keys"])
    assert [str(tuple(v.shape)) for v in sd.values()] == list(g["shapes"])
    assert [n for n, _ in net.named_children()] == list(g["children"])
    # creator/params.py:14-31: backbone = children()[:8], new modules = [8:], split by 'weight' / 'bias' in the name
    kids = list(net.children())
    base = [p for m in kids[:8] for _, p in m.named_parameters()]
    add_w = [p for m in kids[8:] for n, p in m.named_parameters() if "weight" in n]
    add_b = [p for m in kids[8:] for n, p in m.named_parameters() if "bias" in n]
    assert [len(base), len(add_w), len(add_b)] == list(g["group_sizes"])


def test_double_bn_update_closed_form():
    """f_query is f_key and runs twice in the reference: two momentum steps per forward."""
    from acan_b200.modules import PixelAttentionBlock_
    blk = PixelAttentionBlock_(16, 8).train()
    ref_bn = torch.nn.BatchNorm2d(8).train()
    ref_bn.load_state_dict(blk.f_key.bn.state_dict())
    x = torch.randn(3, 16, 5, 7)
    z = blk.f_key.conv(x)
    ref_bn(z), ref_bn(z)
    blk.key_features(x)
    assert torch.allclose(blk.f_key.bn.running_mean, ref_bn.running_mean, atol=1e-6)
    assert torch.allclose(blk.f_key.bn.running_var, ref_bn.running_var, atol=1e-6)
    assert int(blk.f_key.bn.num_batches_tracked) == 2


def test_shard_bounds_cover_batch():
    from acan_b200.parallel import shard_bounds
    for n in (1, 7, 8, 16, 33):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_bn_act_host_semantics_off_gpu():
    """fn.bn_act on tensors the fused kernels do not take (CPU / fp32 / NCHW) is exactly the framework's
    relu(bn(x) + residual), including the running-statistics update and an overridden momentum (f_key's double step)."""
    from acan_b200 import functions as fn
    torch.manual_seed(3)
    x, r = torch.randn(2, 8, 5, 7), torch.randn(2, 8, 5, 7)
    a, b = torch.nn.BatchNorm2d(8).train(), torch.nn.BatchNorm2d(8).train()
    b.load_state_dict(a.state_dict())
    got = fn.bn_act(a, x, relu=True, residual=r)
    want = torch.relu(b(x) + r)
    assert torch.equal(got, want) and torch.equal(a.running_var, b.running_var) and int(a.num_batches_tracked) == 1
    fn.bn_act(a, x, relu=False, momentum=0.19)
    b.momentum = 0.19
    b(x)
    assert torch.allclose(a.running_mean, b.running_mean) and a.momentum == 0.1
    a.eval(), b.eval()
    assert torch.equal(fn.bn_act(a, x, relu=True), torch.relu(b(x)))


def test_lazy_training_handle_shapes_and_materialisation():
    """the training-mode 'y' handle built from stride-8 logits reports the reference's [B,K,H,W] shape and, for any consumer
    other than the drop-in loss / inference, upsamples and decodes lazily (CPU: shape logic only, decode_ord needs the GPU)."""
    from acan_b200.modules import LazyLogitsProb
    z = torch.randn(2, 160, 4, 6, requires_grad=True)
    up = torch.nn.Upsample(scale_factor=8, mode="bilinear", align_corners=True)
    y = LazyLogitsProb(None, 80, z=z, upsample=up)
    assert tuple(y.shape) == (2, 80, 32, 48) and y.size(1) == 80 and y.dim() == 4 and y.requires_grad and y.z is z
    assert tuple(y.logits.shape) == (2, 160, 32, 48) and y.logits.requires_grad
    full = LazyLogitsProb(torch.randn(2, 160, 32, 48), 80)
    assert full.z is None and tuple(full.shape) == (2, 80, 32, 48)


def test_graphed_train_step_refuses_cpu():
    from acan_b200.training import GraphedTrainStep
    with pytest.raises(RuntimeError, match="no CPU path"):
        GraphedTrainStep(None, None, None, torch.zeros(1, 3, 8, 8), torch.zeros(1, 1, 8, 8))


def test_new_entry_points_validate_arguments_without_gpu():
    from acan_b200 import _lib
    lib = _lib.load()
    BAD = 10001
    assert lib.acan_bn_train_fwd(16, 1, 16, 16, None, 16, 16, 16, None, None, 0.1, 1e-5, 1, 16, 128, 12, None) == BAD      # C % 8 != 0
    assert lib.acan_bn_train_fwd(16, 3, 16, 16, None, 16, 16, 16, None, None, 0.1, 1e-5, 1, 16, 128, 16, None) == BAD      # unknown dtype
    assert lib.acan_bn_train_fwd(16, 1, 16, 16, None, 16, 16, 16, 16, None, 0.1, 1e-5, 1, 16, 128, 16, None) == BAD        # running_var missing
    assert lib.acan_bn_train_bwd(16, None, 16, 1, 16, 16, 16, 16, 16, 16, 16, 16, 128, 16, None) == BAD                    # grad_residual without y
    assert lib.acan_tail_loss_fwd(8, 8, 8, 8, 8, 0, 4, 6, 80, 0, None) == BAD                                             # B = 0
    assert lib.acan_tail_loss_bwd(8, 8, 8, 8, None, 1, 4, 6, 80, 0, None) == BAD                                          # null output
    assert lib.acan_attn_bwd(1024, None, None, 1024, None, 1024, 64, 96, None, 1024, None, 1024, 1 << 30, 1, 96, 512, 2048, 0.044, None) == BAD   # label without grad_loss
    assert lib.acan_bn_ws_floats(2048) > 128 * 2048 * 2 and lib.acan_tail_loss_ws_floats(8, 32, 44) >= 2 * 8 * 32 * 11


def test_batched_tta_drivers_match_reference_golden(golden):
    """acan_b200.evaluation.predict_multi_scale / predict_whole_img (every equally-sized forward of the reference's loops in ONE
    batched call) reproduce the reference's own outputs on the toy network -- host logic, no GPU involved."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import synth
    from acan_b200 import evaluation as E
    g = golden("eval_utils")
    net = synth.ToyDepthNet(4).eval()
    img = synth.rand("eval.img", 2, 3, 16, 24)
    calls = []
    fwd = net.forward
    net.forward = lambda x: (calls.append(tuple(x.shape)), fwd(x))[1]
    assert torch.allclose(E.predict_multi_scale(net, img, [1, 1.25], 4, True), torch.as_tensor(g["tta"]), rtol=1e-6, atol=1e-6)
    assert calls == [(20, 3, 16, 24)]                                   # 2 whole views + 2 x 4 tiles, 2 images each: one forward
    assert torch.allclose(E.predict_multi_scale(net, img, [1.25], 4, False), torch.as_tensor(g["tta_noflip"]), rtol=1e-6, atol=1e-6)
    assert torch.allclose(E.predict_whole_img(net, img, 0.75, True), torch.as_tensor(g["whole_075"]), rtol=1e-6, atol=1e-6)
    assert torch.allclose(E.predict_multi_scale(net, img, [1, 1.25], 4, True, max_batch=6), torch.as_tensor(g["tta"]), rtol=1e-6, atol=1e-6)
    with pytest.raises(RuntimeError, match="eval"):
        E.predict_multi_scale(net.train(), img, [1], 4, False)
    from acan_b200 import _lib
    with pytest.raises(_lib.AcanLibraryError, match="no CPU path"):
        E.compute_errors(torch.ones(1, 4, 4), torch.ones(1, 4, 4))
    assert _lib.load().acan_depth_metrics(8, 8, 8, 8, 0, 1, None) == 10001


def test_piece_ranges_partition_the_flat_gradient_buffer():
    """host logic of the multi-GPU training step (acan_b200.training.piece_ranges): the runs handed to the all-reduces / the SGD kernel
    cover every element exactly once, consist of whole optimizer chunks, and in the split schedule come in the order the stages' gradients
    complete, each run starting at or after its stage's first element (never before: those gradients would not be complete yet)."""
    from acan_b200.training import piece_ranges
    for n, ch, pieces in ((92_000_123, 2048, 4), (2048 * 7, 2048, 4), (5000, 2048, 8), (1, 2048, 4), (2048 * 3 + 1, 2048, 1)):
        (runs,) = piece_ranges(n, ch, None, pieces)
        assert runs[0][0][0] == 0 and runs[-1][0][1] == n and len(runs) <= pieces
        for ((lo, hi), (c0, nc)), nxt in zip(runs, runs[1:] + [None]):
            assert lo == c0 * ch and hi == min((c0 + nc) * ch, n) and hi > lo
            assert nxt is None or nxt[0][0] == hi
    n, ch = 92_000_123, 2048
    for s1, s0 in ((1_500_000, 28_000_000), (2048 * 10, 2048 * 20), (100, 200), (0, 0), (n - 5, n - 5), (4096, n)):
        stages = piece_ranges(n, ch, {2: 0, 1: s1, 0: s0}, 4)
        assert len(stages) == 3
        flat = [r for st in stages for r in st]
        covered = sorted(r[0] for r in flat)
        pos = 0
        for lo, hi in covered:                                   # a partition of [0, n)
            assert lo == pos and hi > lo
            pos = hi
        assert pos == n
        for (lo, hi), (c0, nc) in flat:
            assert lo == c0 * ch and hi == min((c0 + nc) * ch, n)
        for k, start in ((0, s0), (1, s1)):                      # a stage's run never begins inside a stage that completes LATER
            for (lo, hi), _ in stages[k]:
                assert lo >= start
