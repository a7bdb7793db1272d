"""GPU tests (-m gpu) against the UNMODIFIED reference running on the same B200 (baseline/_ref, shipped by baseline/fetch_ref.sh; the
tests skip when no copy is present) and against the train-mode golden the reference generated on CPU.

  * the drop-in SADecoder vs the reference's at the configs[1] size, eval and train() with all three Dropout2d active (same CUDA RNG seed);
  * ``acan_b200.network.install`` on a network built by the reference's own ``create_network`` / ``create_lossfunc`` (INTEGRATION.md
    section A, the primary integration path): forward dict, inference, the beta = 1 loss and every parameter gradient vs the same
    reference network left untouched, fp32, TF32 off (depthest_main.py:59-77, get_network.py:3-19, model.py:216-272);
  * the train()-mode SelfAttentionBlock_ golden: context, sim_map, and the double BatchNorm update of the shared f_key / f_query module.
"""
import copy
import os
import sys

import numpy as np
import pytest
import torch

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import refload  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"
T = torch.from_numpy


@pytest.fixture(autouse=True)
def _strict_fp32_library_ops():
    a, b = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = a, b


def _reference():
    try:
        return refload.load_reference()
    except RuntimeError as e:
        pytest.skip(str(e))


def close(a, b, tol=1e-3):
    e = synth.rel_err(a, b)
    assert e <= tol, f"max-norm relative error {e:.3e} > {tol}"


def _dropouts(module, on):
    for m in module.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.train(on)


@pytest.mark.parametrize("hh,ww,b", [(256, 352, 2), (64, 96, 3)])
def test_sadecoder_against_reference_module_on_gpu(hh, ww, b):
    ref = _reference()
    from acan_b200.modules import SADecoder
    rdec = ref.sadecoder.SADecoder(height=hh, width=ww)
    rdec.load_state_dict(synth.fill_state_dict(rdec.state_dict(), seed=5))
    rdec = rdec.to(DEV)
    dec = SADecoder(height=hh, width=ww)
    dec.load_state_dict(rdec.state_dict())
    dec = dec.to(DEV)
    x = synth.feature_map(f"refdec.x.{hh}", b, 2048, hh // 8, ww // 8).to(DEV)
    # eval
    rdec.eval(), dec.eval()
    with torch.no_grad():
        want, want_sim = rdec(x)
        got, sim = dec(x)
    close(got, want)
    close(sim.materialize(), want_sim)
    # train(): batch-statistics BN in f_key (two running-stat steps per forward) and the three Dropout2d, same RNG stream
    rdec.train(), dec.train()
    for seed in (11, 12):
        torch.manual_seed(seed)
        with torch.no_grad():
            want, want_sim = rdec(x.clone())
        torch.manual_seed(seed)
        with torch.no_grad():
            got, sim = dec(x.clone())
        close(got, want)
        assert float((want == 0).float().mean()) > 0.4           # dropout2 really dropped half of the output channels
        assert torch.equal((got == 0).all(dim=(2, 3)), (want == 0).all(dim=(2, 3)))      # ... the same ones
        close(sim.materialize(), want_sim)
    rbn, bn = rdec.saconv.f_key.bn, dec.saconv.f_key.bn
    close(bn.running_mean, rbn.running_mean, 1e-5)
    close(bn.running_var, rbn.running_var, 1e-5)
    assert int(bn.num_batches_tracked) == int(rbn.num_batches_tracked) == 4      # 2 forwards x 2 calls of the shared module


def test_install_on_reference_built_network_forward_backward():
    ref = _reference()
    from acan_b200 import network as b200
    args = refload.reference_args(encoder="resnet50", height=96, width=128, alpha=0.0, beta=1.0)
    torch.manual_seed(3)
    rnet = ref.models.create_network(args)
    rnet.load_state_dict(synth.fill_state_dict(rnet.state_dict(), seed=9))
    rnet = rnet.to(DEV).train()
    rcrit = ref.models.create_lossfunc(args, rnet)
    net = copy.deepcopy(rnet)
    crit = ref.models.create_lossfunc(args, net)
    b200.install(net, crit)
    for m in (rnet, net):
        _dropouts(m, False)
    b = 2
    x = synth.rand("refnet.x", b, 3, args.height, args.width).to(DEV)
    label = synth.depth_label("refnet.label", b, args.height, args.width, 0.5, 10.5).to(DEV)
    want = rnet(x)
    got = net(x)
    close(torch.as_tensor(got["y"].materialize()), want["y"])
    close(got["sim_map"].materialize(), want["sim_map"])
    close(net.inference(got), rnet.inference(want))
    wl = rcrit(want, label.clone(), 1)
    gl = crit(got, label.clone(), 1)
    for g_, w_ in zip(gl, wl):
        if torch.is_tensor(w_):
            assert abs(float(g_) - float(w_)) <= 1e-3 * abs(float(w_)), (float(g_), float(w_))
    wl[3].backward()
    gl[3].backward()
    scale = max(float(p.grad.abs().max()) for p in rnet.parameters() if p.grad is not None)
    worst = 0.0
    for (name, p), (_, q) in zip(net.named_parameters(), rnet.named_parameters()):
        assert (p.grad is None) == (q.grad is None), name
        if q.grad is None:
            continue
        if float(q.grad.abs().max()) < 1e-6 * scale:            # analytically zero (conv bias in front of train-mode BN)
            assert float(p.grad.abs().max()) < 1e-4 * scale, name
            continue
        e = synth.rel_err(p.grad, q.grad)
        worst = max(worst, e)
        # decoder / classifier parameters sit right behind the hot path; backbone gradients have crossed up to 50 train()-mode BN layers,
        # which amplify the attention core's fp16-operand rounding (measured worst case: conv1.weight 9.4e-3)
        assert e <= (5e-3 if name.startswith(("decoder", "classifier")) else 2e-2), (name, e)
    print(f"\ninstall() on a reference-built ResNet-50: worst parameter-gradient error vs the reference on the same GPU {worst:.2e}")
    # eval: the lazy head / fused tail against the reference's full-resolution path -- on CALIBRATED running statistics (SURVEY section 8d:
    # with the synthetic ones the logits exceed 88 and the reference's unstabilised decode_ord returns NaN)
    for m in list(rnet.modules()) + list(net.modules()):
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = 1.0
    with torch.no_grad():
        rnet(x), net(x)
    rnet.eval(), net.eval()
    with torch.no_grad():
        want_d = rnet.inference(rnet(x))
        assert bool(torch.isfinite(want_d).all())
        close(net.inference(net(x)), want_d)


def _load(module, g):
    sd = synth.fill_state_dict({k: torch.empty(eval(s)) for k, s in zip(g["keys"], g["shapes"])})
    module.load_state_dict(sd)


@pytest.mark.parametrize("fused_bn", [False, True])
def test_self_attention_block_train_mode_golden(golden, fused_bn):
    """train()-mode golden generated by the reference (tests/golden/make_golden.py: golden_sab_train): context, sim_map, and the shared
    f_key / f_query BatchNorm after one forward (two momentum steps, num_batches_tracked == 2).  fused_bn: the half-precision channels_last
    step, where the statistics come from acan_bn_train_fwd on the 16-bit convolution output (tolerances of 16-bit activations)."""
    from acan_b200.modules import SelfAttentionBlock_
    g = golden("sab_train")
    blk = SelfAttentionBlock_(64, 64, 128)
    _load(blk, g)
    blk = blk.to(DEV).train()
    x = synth.feature_map("sabt.x", 2, 64, 4, 6).to(DEV)
    if fused_bn:
        blk = blk.to(memory_format=torch.channels_last)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
            ctx, sim = blk(x.half().contiguous(memory_format=torch.channels_last))
        tol, tol_stats = 6e-3, 2e-3
    else:
        with torch.no_grad():
            ctx, sim = blk(x)
        tol, tol_stats = 1e-3, 1e-5
    close(ctx.float(), T(g["ctx_train"]), tol)
    close(sim.materialize(), T(g["sim_train"]), tol)
    bn = blk.f_key.bn
    close(bn.running_mean, T(g["rm_after"]), tol_stats)
    close(bn.running_var, T(g["rv_after"]), tol_stats)
    assert int(bn.num_batches_tracked) == int(g["nbt_after"]) == 2
    assert blk.f_query.bn is bn and np.array_equal(g["rm_after"], g["rm_q_after"])


def test_graph_replay_survives_eager_forwards_of_another_batch_size():
    """A captured CUDA graph bakes the attention workspace's address in; an eager forward of the same network with another batch size in
    between (validation loop's last batch, test-time augmentation) must not free or move it."""
    from acan_b200.inference import GraphedPredictor, optimize_for_inference
    from acan_b200.network import ResNet
    hh, ww = 64, 96
    net = ResNet(0.72, 10.0, 80, "OR", "soft", "attention", hh, ww, layers=[1, 1, 1, 1])
    net.load_state_dict(synth.fill_state_dict(net.state_dict()))
    net = net.to(DEV)
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.momentum = 1.0
    net.train()
    _dropouts(net, False)
    x4 = synth.rand("gr.x4", 4, 3, hh, ww).to(DEV)
    with torch.no_grad():
        net(x4)
    fast = optimize_for_inference(net.eval())
    pred = GraphedPredictor(fast, x4)
    first = pred.predict(x4).clone()
    second = pred.predict(x4).clone()
    x3 = synth.rand("gr.x3", 3, 3, hh, ww).to(DEV)
    with torch.no_grad():
        for _ in range(2):
            fast.inference(fast(x3))                       # eager, another batch size: allocates a second workspace
            fast.inference(fast(torch.cat([x3, x3])))
    torch.cuda.empty_cache()
    junk = [torch.full((64 << 20,), 7, dtype=torch.uint8, device=DEV) for _ in range(4)]      # would land in a freed workspace
    again = pred.predict(x4).clone()
    again2 = pred.predict(x4).clone()
    torch.cuda.synchronize()
    del junk
    assert torch.equal(first, again) and torch.equal(second, again2)
