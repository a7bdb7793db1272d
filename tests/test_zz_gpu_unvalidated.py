"""GPU tests (-m gpu) of code that has not run on hardware yet (written after the round's GPU budget was spent).  Nothing in the default
training / inference paths calls this code (GraphedTrainStep(fused_grad_cast=True) opts in).  The tests run only on request
(RUN_UNVALIDATED_GPU_TESTS=1 python -m pytest tests/test_zz_gpu_unvalidated.py -m gpu): a kernel that has never executed must not be able
to poison the CUDA context of, or stop, the validated suite in a `pytest -x` run.  The file sorts last and the tests are xfail(strict=False)
on top of that."""
import os

import pytest
import torch

import synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("RUN_UNVALIDATED_GPU_TESTS") != "1", reason="opt-in: code that has not run on hardware yet"),
              pytest.mark.xfail(strict=False, reason="first run on hardware: acan_cast_to_f32_multi was written after the GPU budget was spent")]
DEV = "cuda"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_multi_tensor_gradient_cast_is_exact(dtype):
    """acan_cast_to_f32_multi: 16-bit tensors -> fp32 slices of one flat buffer in one launch per 96 tensors == .float(), bit for bit;
    odd sizes, destinations that are only 4-byte aligned, channels_last weight layouts, more tensors than one launch carries."""
    from acan_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(11)
    shapes = [(64, 3, 7, 7), (1,), (7,), (4097,), (256, 64, 1, 1), (64, 64, 3, 3), (2048 * 5 + 3,), (160,)] + [(33 + i,) for i in range(100)]
    srcs = []
    for sh in shapes:
        t = torch.randn(*sh, generator=g).to(DEV).to(dtype)
        if len(sh) == 4:
            t = t.contiguous(memory_format=torch.channels_last)
        srcs.append(t)
    flat = torch.full((sum(t.numel() for t in srcs) + 1,), -7.0, device=DEV)
    dsts, off = [], 1                                            # offset 1: no destination is 16-byte aligned by construction
    for t in srcs:
        seg = flat[off:off + t.numel()]
        off += t.numel()
        dsts.append(seg.view(t.shape[0], t.shape[2], t.shape[3], t.shape[1]).permute(0, 3, 1, 2) if t.dim() == 4 else seg.view(t.shape))
    assert ops.cast_to_f32_multi(dsts, srcs)
    for d, t in zip(dsts, srcs):
        assert torch.equal(d, t.float())
    assert float(flat[0]) == -7.0
    flat2 = torch.zeros(8192 * 3, device=DEV)                    # aligned: the vector path
    big = torch.randn(8192 * 3, generator=g).to(DEV).to(dtype)
    assert ops.cast_to_f32_multi([flat2], [big]) and torch.equal(flat2, big.float())
    assert not ops.cast_to_f32_multi([flat2[:10]], [big[:20:2]])   # strided source: refused, nothing launched


