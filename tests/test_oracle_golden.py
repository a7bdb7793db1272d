"""Pins oracle/acan_oracle.py against outputs of the reference itself (tests/golden/*.npz, made by
tests/golden/make_golden.py from /root/reference) and the closed-form KATs of SURVEY.md §8c."""
import math

import numpy as np
import pytest
import torch

import synth
from oracle import acan_oracle as O

T = torch.from_numpy


def close(a, b, tol=1e-6):
    assert synth.rel_err(a, b) <= tol, synth.rel_err(a, b)


@pytest.mark.parametrize("tag", ["nyu", "kitti"])
def test_head_matches_reference(golden, tag):
    g = golden("head_" + tag)
    dmin, dmax, K = g["params"]
    K = int(K)
    prob = T(g["prob"])
    close(O.decode_ord(T(g["y"])), prob)
    close(O.soft_ordinal_regression(prob, dmin, dmax, K), T(g["or_soft"]))
    close(O.hard_ordinal_regression(prob, dmin, dmax, K), T(g["or_hard"]))
    assert torch.equal(O.hard_ordinal_index(prob).int(), T(g["or_hard_idx"]))
    close(O.soft_ordinal_regression(T(g["sat"]), dmin, dmax, K), T(g["or_soft_sat"]))
    close(O.hard_ordinal_regression(T(g["sat"]), dmin, dmax, K), T(g["or_hard_sat"]))
    close(O.soft_cross_entropy(T(g["score"]), dmin, dmax, K), T(g["ce_soft"]))
    close(O.hard_cross_entropy(T(g["score"]), dmin, dmax, K), T(g["ce_hard"]))
    assert torch.equal(O.hard_cross_entropy_index(T(g["score"])).int(), T(g["ce_hard_idx"]))
    assert torch.equal(O.continuous2discrete(T(g["label"]), dmin, dmax, K), T(g["c2d"]))
    close(O.discrete2continuous(torch.arange(K + 3).float(), dmin, dmax, K), T(g["d2c"]))
    for mode, cls, key in (("soft", "OR", "or_soft"), ("hard", "OR", "or_hard")):
        close(O.inference(prob, cls, mode, dmin, dmax, K), T(g[key]))


def test_head_known_answers():
    dmin, dmax, K = 0.72, 10.0, 80
    k = torch.tensor([0.0, K - 1.0])
    d = O.discrete2continuous(k, dmin, dmax, K)
    assert abs(float(d[0]) - dmin) < 1e-6 and abs(float(d[1]) - dmax) < 1e-5
    ones = torch.ones(1, K, 1, 1)
    assert int(O.hard_ordinal_index(ones)) == K
    s80 = O.soft_ordinal_regression(ones, dmin, dmax, K)           # s = 80 -> bins 80, 81 (extrapolated), F = 0
    want = 0.5 * (O.discrete2continuous(torch.tensor(80.0), dmin, dmax, K) + O.discrete2continuous(torch.tensor(81.0), dmin, dmax, K))
    assert abs(float(s80) - float(want)) < 1e-5
    tie = torch.zeros(1, 2 * K, 1, 1)                              # y0 == y1 -> p = 0.5 exactly -> not counted
    assert int(O.hard_ordinal_index(O.decode_ord(tie))) == 0


def test_self_attention_block_matches_reference(golden):
    g = golden("sab_small")
    sd = {k[3:]: T(g[k]) for k in g.files if k.startswith("sd.")}
    x = T(g["x"])
    ctx, sim, mom = O.self_attention_block(x, sd, "", training=False)
    assert mom is None
    close(sim, T(g["sim_eval"]), 2e-6)
    close(ctx, T(g["ctx_eval"]), 2e-6)
    ctx, sim, mom = O.self_attention_block(x, sd, "", training=True)
    close(sim, T(g["sim_train"]), 2e-6)
    close(ctx, T(g["ctx_train"]), 2e-6)
    rm, rv = O.bn_running_update_twice(sd["f_key.bn.running_mean"], sd["f_key.bn.running_var"], mom)
    close(rm, T(g["rm_after"]), 1e-6)
    close(rv, T(g["rv_after"]), 1e-6)
    assert int(g["nbt_after"]) == 2                                # f_query is f_key, called twice


@pytest.mark.parametrize("name", ["sadecoder_24", "sadecoder_1064"])
def test_sadecoder_matches_reference(golden, name):
    g = golden(name)
    hh, ww, b = (int(v) for v in g["hw"])
    shapes = {k: eval(s) for k, s in zip(g["keys"], g["shapes"])}
    sd = synth.fill_state_dict({k: torch.empty(s) for k, s in shapes.items()})
    x = synth.feature_map(name + ".x", b, 2048, hh // 8, ww // 8)
    assert abs(float(x.double().sum()) - float(g["x_sum"])) < 1e-6 * abs(float(g["x_sum"]))
    assert abs(float(sum(v.double().sum() for v in sd.values())) - float(g["w_sum"])) < 1e-6 * abs(float(g["w_sum"]))
    with torch.no_grad():
        feat, _ = O.key_projection(x, sd, "saconv.f_key.", False)
        val = O.value_projection(x, sd, "saconv.f_value.")
        out, sim, _ = O.sadecoder_forward(x, sd, training=False)
        ctx = O.aggregate(sim, val)
        gvec = O.image_context(x, sd, "image_context.")
    for t, key, tol in ((feat, "feat", 1e-5), (val, "val", 1e-5), (sim, "sim", 1e-4), (ctx, "ctx", 1e-4), (out, "out", 1e-4)):
        got = t.reshape(-1)[T(g[key + "_idx"])]
        close(got, T(g[key + "_val"]), tol)
    close(gvec, T(g["g"]), 1e-5)
    close(sim.sum(-1), T(g["sim_rowsum"]), 1e-5)


def test_gt_affinity_known_answer(golden):
    g = golden("attention_loss")
    a = torch.tensor([[3.0, 6.0, 0.0, 12.0]])
    want = torch.tensor([[4 / 7, 2 / 7, 0, 1 / 7], [0.25, 0.5, 0, 0.25], [0, 0, 0, 0], [1 / 7, 2 / 7, 0, 4 / 7]])
    close(O.gt_affinity(a)[0], want, 1e-6)
    close(O.gt_affinity_literal(a)[0], want, 1e-6)
    close(O.gt_affinity(a), T(g["toy_gt"]), 1e-6)
    bins = O.nearest_downsample8(T(g["toy_label"])).reshape(1, -1)
    assert bins.tolist() == [[3.0, 6.0, 0.0, 12.0]]


@pytest.mark.parametrize("tag", ["a", "b"])
def test_attention_loss_matches_reference(golden, tag):
    g = golden("attention_loss")
    dmin, dmax, K = g[tag + "_params"]
    dis = O.continuous2discrete(T(g[tag + "_label"]), dmin, dmax, int(K))
    assert torch.equal(dis, T(g[tag + "_dis"]))
    sim = T(g[tag + "_sim"])
    bins = O.nearest_downsample8(dis).reshape(dis.shape[0], -1)
    gt = O.gt_affinity(bins)
    close(gt, T(g[tag + "_gt"]), 1e-6)
    close(O.gt_affinity_literal(bins), T(g[tag + "_gt"]), 1e-6)
    for literal in (False, True):
        loss = O.attention_loss(sim, dis, literal=literal)
        assert abs(float(loss) - float(g[tag + "_loss"])) <= 2e-6 * abs(float(g[tag + "_loss"]))
    close(O.attention_loss_grad_sim(sim, gt), T(g[tag + "_grad"]), 1e-5)


def test_attention_loss_clamp_window(golden):
    g = golden("attention_loss")
    sim, dis = T(g["c_sim"]), T(g["c_dis"])
    loss = O.attention_loss(sim, dis)
    assert math.isfinite(float(loss))
    assert abs(float(loss) - float(g["c_loss"])) <= 2e-6 * abs(float(g["c_loss"]))
    gt = O.gt_affinity(O.nearest_downsample8(dis).reshape(1, -1))
    grad = O.attention_loss_grad_sim(sim, gt)
    want = T(g["c_grad"])
    # Q == 0 exactly with P > 0: reference autograd yields 0 * inf = NaN; the oracle (and the kernels)
    # define the out-of-window gradient as 0 there -- the one documented deviation (hazards).
    dead = sim == 0
    assert torch.isnan(want[dead & (gt > 0)]).all()
    assert torch.equal((grad == 0)[~dead], (want == 0)[~dead])     # zero gradient outside the clamp window
    assert float(grad[0, 5, 7]) == 0.0 and float(want[0, 5, 7]) == 0.0
    close(grad[~dead], want[~dead], 1e-5)


def test_full_network_matches_reference(golden):
    g = golden("resnet_tiny")
    hh, ww, b = (int(v) for v in g["hw"])
    shapes = {k: eval(s) for k, s in zip(g["keys"], g["shapes"])}
    sd = synth.fill_state_dict({k: torch.empty(s) for k, s in shapes.items()})
    x = synth.rand("resnet.x", b, 3, hh, ww)
    label = synth.depth_label("resnet.label", b, hh, ww, 0.5, 10.5)
    with torch.no_grad():
        out = O.acan_forward(x, sd, [1, 1, 1, 1], "OR", training=True)     # BN on batch statistics, dropout off
    close(out["y"].reshape(-1)[T(g["y_idx"])], T(g["y_val"]), 1e-4)
    close(out["sim_map"].reshape(-1)[T(g["sim_idx"])], T(g["sim_val"]), 1e-4)
    close(O.inference(out["y"], "OR", "soft", 0.72, 10.0, 80), T(g["depth_soft"]), 1e-4)
    hard = O.inference(out["y"], "OR", "hard", 0.72, 10.0, 80)
    assert (hard - T(g["depth_hard"])).abs().gt(1e-4).float().mean() < 1e-3   # conv rounding may flip a p~0.5 bin
    dis = O.continuous2discrete(label, 0.72, 10.0, 80)
    l3 = O.attention_loss(out["sim_map"], dis)
    assert abs(float(l3) - float(g["loss3"])) <= 1e-4 * abs(float(g["loss3"]))
    l1 = O.ordinal_regression_loss(out["y"], dis.squeeze(1).long())
    assert abs(float(l1) - float(g["loss1"])) <= 1e-4 * abs(float(g["loss1"]))


def test_evaluation_helpers_against_reference_golden(golden):
    """row f-4: compute_errors and the test-time-augmentation drivers (utils/eval_utils.py) restated in the oracle vs the
    reference's own outputs on the same inputs (tests/golden/make_golden.py: golden_eval)."""
    g = golden("eval_utils")
    gt = synth.depth_label("eval.gt", 3, 32, 48, 0.5, 10.5).squeeze(1)
    pred = gt * (1.0 + 0.3 * synth.randn("eval.noise", 3, 32, 48)).abs() + 0.05 * synth.rand("eval.off", 3, 32, 48)
    pred[0, 0, :4] = torch.tensor([0.0, 1e-8, 1e7, 5.0])
    m = O.compute_errors(gt, pred)
    assert [str(k) for k in g["measure_list"]] == list(m.keys())
    want = torch.as_tensor(g["measures"])
    got = torch.stack([m[k] for k in m])
    assert torch.allclose(got, want, rtol=1e-5, atol=0), (got, want)
    net = synth.ToyDepthNet(4).eval()
    img = synth.rand("eval.img", 2, 3, 16, 24)
    with torch.no_grad():
        assert torch.allclose(O.predict_multi_scale(net, img, [1, 1.25], 4, True), torch.as_tensor(g["tta"]), rtol=1e-6, atol=1e-6)
        assert torch.allclose(O.predict_multi_scale(net, img, [1.25], 4, False), torch.as_tensor(g["tta_noflip"]), rtol=1e-6, atol=1e-6)
        assert torch.allclose(O.predict_whole_img(net, img, 0.75, True), torch.as_tensor(g["whole_075"]), rtol=1e-6, atol=1e-6)


def test_cross_entropy_loss_against_reference_golden(golden):
    """row f-2 (CE classifier): the oracle's CrossEntropy2d restatement vs the reference's loss and gradient (plain, uniform- and
    gaussian-smoothed targets; ignored pixels; a saturated pixel that hits both clamp windows)."""
    g = golden("ce_loss")
    K = 80
    y = 2.0 * synth.randn("ce.y", 2, K, 8, 12)
    y[0, :, 0, 0] = -30.0
    y[0, 7, 0, 0] = 30.0
    lab = O.continuous2discrete(synth.depth_label("ce.lab", 2, 8, 12, 0.5, 10.5), 0.72, 10.0, 80).squeeze(1).long()
    lab[0, 0, 0], lab[0, 0, 1] = 7, 0
    for tag, eps, prior in (("plain", 0.0, "uniform"), ("uniform", 0.1, "uniform"), ("gaussian", 0.2, "gaussian")):
        yy = y.clone().requires_grad_(True)
        loss = O.cross_entropy_loss(yy, lab, 0, eps, prior)
        loss.backward()
        assert torch.allclose(loss.detach(), torch.as_tensor(g[f"loss_{tag}"]), rtol=1e-5), tag
        assert torch.allclose(yy.grad, torch.as_tensor(g[f"grad_{tag}"]), rtol=1e-4, atol=1e-7), tag
