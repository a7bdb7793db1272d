#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

The reference (miraiaroha/ACAN, ``/root/reference/code/models``) is imported as is; the only
out-of-tree changes are the two shims of SURVEY.md §8c:
  1. ``models.model.continuous2discrete``: ``1 - bool*bool`` raises on torch>=1.2; replaced by the
     torch-0.4 meaning (mask = not (d_min < depth < d_max)).
  2. ``Tensor.cuda`` is a no-op on this CPU-only host (model.py:54, losses.py:128,155 hard-code it).
Inputs and weights come from ``tests/synth.py`` so the tests can rebuild them bit-identically; the
.npz files hold reference OUTPUTS (sub-sampled when large) and the small inputs.
The reference cannot travel to the GPU box, these files do.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))           # tests/
sys.path.insert(0, "/root/reference/code")
import synth  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self      # shim 2
import models  # noqa: E402  (the reference)
import models.model as ref_model  # noqa: E402
from models.losses import AttentionLoss2d, OrdinalRegression2d  # noqa: E402
from models.sadecoder import SADecoder, SelfAttentionBlock_  # noqa: E402


def _c2d(depth, d_min, d_max, n_c):                  # shim 1
    mask = ~((depth > d_min) & (depth < d_max))
    depth = torch.round(torch.log(depth / d_min) / np.log(d_max / d_min) * (n_c - 1))
    depth[mask] = 0
    return depth


ref_model.continuous2discrete = _c2d

NYU = (0.72, 10.0, 80)
KITTI = (1.98, 80.0, 80)


def save(name, **arrs):
    out = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()}
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB  " + " ".join(f"{k}{tuple(v.shape)}" for k, v in out.items()))


def sub(t, n=4096):
    """deterministic sub-sample of a large tensor: (flat indices, values)."""
    flat = t.detach().reshape(-1)
    step = max(1, flat.numel() // n)
    idx = torch.arange(0, flat.numel(), step)
    return idx, flat[idx]


@torch.no_grad()
def golden_head():
    for tag, (dmin, dmax, K) in (("nyu", NYU), ("kitti", KITTI)):
        base = ref_model.BaseClassificationModel_(dmin, dmax, K, "OR", "soft", "attention")
        y = 3.0 * synth.randn(f"head.{tag}.y", 2, 2 * K, 8, 12)
        prob = base.decode_ord(y)
        score = 3.0 * synth.randn(f"head.{tag}.score", 2, K, 8, 12)
        score[0, 5, 0, 0] = score[0, 9, 0, 0] = score[0].max() + 1.0      # a tie: first index wins
        sat = prob.clone()
        sat[1, :, 0, 0] = 1.0                                              # s = K: needs d(K+1), d(K+2)
        sat[1, :, 0, 1] = 0.0
        lab = synth.depth_label(f"head.{tag}.label", 2, 8, 12, dmin * 0.7, dmax * 1.05)
        save(f"head_{tag}", y=y, prob=prob, score=score, sat=sat, label=lab,
             or_soft=base.soft_ordinal_regression(prob, dmin, dmax, K),
             or_hard=base.hard_ordinal_regression(prob, dmin, dmax, K),
             or_hard_idx=(prob > 0.5).float().sum(1, keepdim=True).to(torch.int32),
             or_soft_sat=base.soft_ordinal_regression(sat, dmin, dmax, K),
             or_hard_sat=base.hard_ordinal_regression(sat, dmin, dmax, K),
             ce_soft=base.soft_cross_entropy(score, dmin, dmax, K),
             ce_hard=base.hard_cross_entropy(score, dmin, dmax, K),
             ce_hard_idx=torch.argmax(score, 1, keepdim=True).to(torch.int32),
             c2d=ref_model.continuous2discrete(lab.clone(), dmin, dmax, K),
             d2c=ref_model.discrete2continuous(torch.arange(K + 3).float(), dmin, dmax, K),
             params=np.array([dmin, dmax, K], dtype=np.float64))


def golden_sab_small():
    blk = SelfAttentionBlock_(64, 32, 48)
    sd = synth.fill_state_dict(blk.state_dict())
    blk.load_state_dict(sd)
    x = synth.feature_map("sab.x", 2, 64, 5, 7)
    blk.eval()
    with torch.no_grad():
        ctx_e, sim_e = blk(x)
    blk.train()
    with torch.no_grad():
        ctx_t, sim_t = blk(x)
    after = blk.state_dict()
    save("sab_small", x=x, **{"sd." + k: v for k, v in sd.items()},
         ctx_eval=ctx_e, sim_eval=sim_e, ctx_train=ctx_t, sim_train=sim_t,
         rm_after=after["f_key.bn.running_mean"], rv_after=after["f_key.bn.running_var"],
         nbt_after=after["f_key.bn.num_batches_tracked"])


def golden_sab_train():
    """train()-mode SelfAttentionBlock_ at a size the tensor-core kernels take (dk = 64, dv = 128, N = 24): context, sim_map and the
    f_key BatchNorm buffers after ONE forward -- the reference calls the shared f_key / f_query module twice (sadecoder.py:44-45), so
    the running statistics take two momentum steps and num_batches_tracked ends at 2."""
    blk = SelfAttentionBlock_(64, 64, 128)
    sd = synth.fill_state_dict(blk.state_dict())
    blk.load_state_dict(sd)
    x = synth.feature_map("sabt.x", 2, 64, 4, 6)
    blk.train()
    with torch.no_grad():
        ctx_t, sim_t = blk(x)
    after = blk.state_dict()
    save("sab_train", keys=np.array(list(sd.keys())), shapes=np.array([str(tuple(v.shape)) for v in sd.values()]),
         ctx_train=ctx_t, sim_train=sim_t, rm_after=after["f_key.bn.running_mean"], rv_after=after["f_key.bn.running_var"],
         nbt_after=after["f_key.bn.num_batches_tracked"], rm_q_after=after["f_query.bn.running_mean"])


def golden_sadecoder():
    for name, (hh, ww, b) in (("sadecoder_24", (32, 48, 2)), ("sadecoder_1064", (224, 304, 1))):
        dec = SADecoder(height=hh, width=ww)
        ref_sd = dec.state_dict()
        sd = synth.fill_state_dict(ref_sd)
        dec.load_state_dict(sd)
        dec.eval()
        x = synth.feature_map(name + ".x", b, 2048, hh // 8, ww // 8)
        with torch.no_grad():
            out, sim = dec(x)
            g = dec.image_context(x)[:, :, 0, 0]
            val = dec.saconv.f_value(x)
            feat = dec.saconv.f_key(x)
            ctx = dec.saconv(x)[0]
        oi, ov = sub(out)
        si, sv = sub(sim)
        ci, cv = sub(ctx)
        vi, vv = sub(val)
        fi, fv = sub(feat)
        save(name, hw=np.array([hh, ww, b]), keys=np.array(list(ref_sd.keys())),
             shapes=np.array([str(tuple(v.shape)) for v in ref_sd.values()]),
             param_names=np.array([n for n, _ in dec.named_parameters()]),
             x_sum=x.double().sum(), w_sum=sum(v.double().sum() for v in sd.values()),
             out_idx=oi, out_val=ov, sim_idx=si, sim_val=sv, ctx_idx=ci, ctx_val=cv,
             val_idx=vi, val_val=vv, feat_idx=fi, feat_val=fv, g=g,
             sim_rowsum=sim.sum(-1), out_absmax=out.abs().max(), ctx_absmax=ctx.abs().max())


def golden_attention_loss():
    crit = AttentionLoss2d(scale=1)
    # known-answer toy of SURVEY §8c: a = [3, 6, 0, 12] on a 2x2 stride-8 grid
    toy = torch.zeros(1, 1, 16, 16)
    toy[0, 0, 0, 0], toy[0, 0, 0, 8], toy[0, 0, 8, 0], toy[0, 0, 8, 8] = 3, 6, 0, 12
    toy_gt = crit.get_gt_sim_map(toy)
    arrs = dict(toy_label=toy, toy_gt=toy_gt)
    for tag, (hh, ww, b, (dmin, dmax, K)) in (("a", (32, 48, 2, NYU)), ("b", (40, 56, 3, KITTI))):
        n = (hh // 8) * (ww // 8)
        logits = 2.5 * synth.randn(f"al.{tag}.logits", b, n, n)
        sim = torch.softmax(logits, -1).requires_grad_(True)
        label = synth.depth_label(f"al.{tag}.label", b, hh, ww, dmin * 0.7, dmax * 1.05)
        dis = ref_model.continuous2discrete(label.clone(), dmin, dmax, K)
        loss = crit(sim, dis)
        loss.backward()
        arrs.update({f"{tag}_sim": sim.detach(), f"{tag}_label": label, f"{tag}_dis": dis,
                     f"{tag}_gt": crit.get_gt_sim_map(dis), f"{tag}_loss": loss.detach(), f"{tag}_grad": sim.grad,
                     f"{tag}_params": np.array([dmin, dmax, K], dtype=np.float64)})
    # clamp window: a row whose prediction is (almost) zero where the target is not
    sim = torch.softmax(2.5 * synth.randn("al.c.logits", 1, 24, 24), -1)
    sim[0, 3, :] = 0.0
    sim[0, 3, 0] = 1.0
    sim[0, 5, 7] = 1e-12
    sim = sim.requires_grad_(True)
    label = synth.depth_label("al.c.label", 1, 32, 48, 1.0, 9.0)
    label[label == 0] = 2.0                      # no invalid pixels: P > 0 everywhere, so Q = 0 hits the clamp, not 0/0
    dis = ref_model.continuous2discrete(label.clone(), *NYU)
    loss = crit(sim, dis)
    loss.backward()
    arrs.update(c_sim=sim.detach(), c_dis=dis, c_loss=loss.detach(), c_grad=sim.grad)
    save("attention_loss", **arrs)


def golden_resnet_tiny():
    import argparse
    hh, ww, b = 64, 96, 2
    dmin, dmax, K = NYU
    net = ref_model.ResNet(dmin, dmax, K, "OR", "soft", "attention", hh, ww, alpha=0, beta=1.0, layers=[1, 1, 1, 1])
    ref_sd = net.state_dict()
    sd = synth.fill_state_dict(ref_sd)
    net.load_state_dict(sd)
    net.train()                                  # BN on batch statistics (never eval() on raw stats) ...
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.eval()                             # ... with the three Dropout2d sites disabled
    x = synth.rand("resnet.x", b, 3, hh, ww)
    label = synth.depth_label("resnet.label", b, hh, ww, 0.5, 10.5)
    with torch.no_grad():
        out = net(x)
        depth_soft = net.inference(out)
        net.inferenceType = "hard"
        depth_hard = net.inference(out)
        crit = net.LossFunc(dmin, dmax, K, AppearanceLoss=OrdinalRegression2d(ignore_index=0, reduction="sum"),
                            AttentionLoss=AttentionLoss2d(scale=1), alpha=0.0, beta=1.0)
        l1, _, l3, tot = crit(out, label.clone(), 1)
    args = argparse.Namespace(encoder="resnet50", optimizer="sgd", lr=1e-4, final_lr=0)
    sys.path.insert(0, "/root/reference/code")
    from creator.params import create_params
    groups = create_params(args, net)
    group_sizes = [len(list(g["params"])) for g in groups]
    yi, yv = sub(out["y"], 8192)
    si, sv = sub(out["sim_map"], 4096)
    save("resnet_tiny", hw=np.array([hh, ww, b]), keys=np.array(list(ref_sd.keys())),
         shapes=np.array([str(tuple(v.shape)) for v in ref_sd.values()]),
         children=np.array([n for n, _ in net.named_children()]), group_sizes=np.array(group_sizes),
         x_sum=x.double().sum(), w_sum=sum(v.double().sum() for v in sd.values()),
         y_idx=yi, y_val=yv, sim_idx=si, sim_val=sv, depth_soft=depth_soft, depth_hard=depth_hard,
         loss1=l1, loss3=l3, total=tot)


def golden_ce_loss():
    """models/losses.py CrossEntropy2d (reduction='sum') on class scores: plain, uniform- and gaussian-smoothed, with ignored pixels and a
    saturated pixel (clamp windows); loss and d loss / d scores."""
    from models.losses import CrossEntropy2d
    K = 80
    y = 2.0 * synth.randn("ce.y", 2, K, 8, 12)
    y[0, :, 0, 0] = -30.0
    y[0, 7, 0, 0] = 30.0                                               # p_7 -> 1, the others -> ~1e-26: both clamps
    lab = _c2d(synth.depth_label("ce.lab", 2, 8, 12, 0.5, 10.5), *NYU).squeeze(1).long()
    lab[0, 0, 0], lab[0, 0, 1] = 7, 0
    out = {}
    for tag, eps, prior in (("plain", 0.0, "uniform"), ("uniform", 0.1, "uniform"), ("gaussian", 0.2, "gaussian")):
        yy = y.clone().requires_grad_(True)
        loss = CrossEntropy2d(ignore_index=0, eps=eps, priorType=prior)(yy, lab)
        loss.backward()
        out[f"loss_{tag}"], out[f"grad_{tag}"] = loss.detach(), yy.grad
    save("ce_loss", **out)


def golden_eval():
    """utils/eval_utils.py (imported stand-alone: utils/__init__ pulls in matplotlib): compute_errors on a batch with invalid pixels,
    zeros and outliers in the prediction; predict_multi_scale (whole image at scale 1, 2x2 tiles at 1.25, with flips) on the toy net."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_eval_utils", "/root/reference/code/utils/eval_utils.py")
    ev = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ev)
    gt = synth.depth_label("eval.gt", 3, 32, 48, 0.5, 10.5).squeeze(1)
    pred = gt * (1.0 + 0.3 * synth.randn("eval.noise", 3, 32, 48)).abs() + 0.05 * synth.rand("eval.off", 3, 32, 48)
    pred[0, 0, :4] = torch.tensor([0.0, 1e-8, 1e7, 5.0])                  # clamp windows of the safe log, division by zero in thresh
    m = ev.compute_errors(gt.clone(), pred.clone())
    net = synth.ToyDepthNet(4).eval()
    img = synth.rand("eval.img", 2, 3, 16, 24)
    with torch.no_grad():
        tta = ev.predict_multi_scale(net, img, [1, 1.25], 4, True)
        tta_noflip = ev.predict_multi_scale(net, img, [1.25], 4, False)
        whole = ev.predict_whole_img(net, img, scale=0.75, flip=True)
    save("eval_utils", measures=torch.stack([m[k] for k in ev.measure_list]), measure_list=np.array(ev.measure_list),
         tta=tta, tta_noflip=tta_noflip, whole_075=whole)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "eval":          # python tests/golden/make_golden.py eval  -> only the row f-4 fixtures
        golden_eval()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "sabt":          # ... sabt -> only the train-mode SelfAttentionBlock_ fixture
        golden_sab_train()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "ce":            # ... ce -> only the CrossEntropy2d fixtures (row f-2)
        golden_ce_loss()
        sys.exit(0)
    torch.manual_seed(synth.SEED)
    torch.set_num_threads(os.cpu_count() or 1)
    golden_head()
    golden_sab_small()
    golden_sab_train()
    golden_attention_loss()
    golden_sadecoder()
    golden_resnet_tiny()
    golden_eval()
    golden_ce_loss()
