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
