#!/usr/bin/env python
"""Bring-up probe for the B200 box: runs each kernel family in its own subprocess (a trapped kernel must not
take the others down), checks against torch-on-GPU fp64 maths, prints diagnostics and rough timings.

    python scripts/gpu_probe.py [stage ...]      stages: head pool kl attn_small attn_shapes attn_full
Output goes to stdout; run under gpurun with `> gpurun_out/probe.log`.
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timeit(fn, iters=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def relerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def stage_head():
    import torch
    from acan_b200 import ops
    from oracle import acan_oracle as O
    torch.manual_seed(0)
    dev = "cuda"
    dmin, dmax, K = 0.72, 10.0, 80
    y = (3 * torch.randn(2, 2 * K, 32, 48)).to(dev)
    prob = ops.decode_ord(y)
    ref = O.decode_ord(y.cpu())
    print("decode_ord relerr", relerr(prob.cpu(), ref))
    for cls, mode in (("OR", "soft"), ("OR", "hard"), ("CE", "soft"), ("CE", "hard")):
        x = prob if cls == "OR" else y[:, :K].contiguous()
        d, idx = ops.head_inference(x, cls, mode, dmin, dmax, K, want_index=True)
        r = O.inference(x.cpu(), cls, mode, dmin, dmax, K)
        print(f"head {cls}-{mode}: relerr {relerr(d.cpu(), r):.3e}", end="")
        if mode == "hard":
            ri = O.hard_ordinal_index(x.cpu()) if cls == "OR" else O.hard_cross_entropy_index(x.cpu())
            print(" idx mismatches", int((idx.cpu().long() != ri).sum()), end="")
        print()
        if cls == "OR":
            d2, idx2 = ops.head_inference(y, cls, mode, dmin, dmax, K, from_logits=True, want_index=True)
            print(f"   fused-from-logits relerr {relerr(d2.cpu(), r):.3e}  idx mismatches vs prob path {int((idx2 != idx).sum())}")
    g = torch.randn_like(prob)
    gy = ops.decode_ord_backward(prob, g)
    yy = y.clone().requires_grad_(True)
    e = torch.exp(yy.reshape(2, K, 2, 32, 48))
    (e[:, :, 1] / e.sum(2)).backward(g)
    print("decode_ord bwd relerr", relerr(gy, yy.grad))
    lab = (torch.rand(2, 1, 32, 48) * 12).to(dev)
    print("c2d mismatches", int((ops.continuous2discrete(lab, dmin, dmax, K).cpu() != O.continuous2discrete(lab.cpu(), dmin, dmax, K)).sum()))
    # bandwidth at C2 / C3 sizes
    for name, (B, H, W) in (("C2", (16, 256, 352)), ("C3", (32, 160, 512))):
        HW = H * W
        yb = 3 * torch.randn(B, 2 * K, H, W, device=dev)
        pb = ops.decode_ord(yb)
        t = timeit(lambda: ops.decode_ord(yb))
        print(f"{name} decode_ord          {t*1e6:8.1f} us  {B*960*HW/t/1e9:7.1f} GB/s")
        for cls, mode in (("OR", "soft"), ("OR", "hard"), ("CE", "soft"), ("CE", "hard")):
            t = timeit(lambda: ops.head_inference(pb, cls, mode, dmin, dmax, K))
            print(f"{name} head {cls}-{mode} prob   {t*1e6:8.1f} us  {B*324*HW/t/1e9:7.1f} GB/s")
        for mode in ("soft", "hard"):
            t = timeit(lambda: ops.head_inference(yb, "OR", mode, dmin, dmax, K, from_logits=True))
            print(f"{name} head OR-{mode} logits {t*1e6:8.1f} us  {B*644*HW/t/1e9:7.1f} GB/s")
        # full-size property: fused path == two-step path
        d1 = ops.head_inference(pb, "OR", "soft", dmin, dmax, K)
        d2 = ops.head_inference(yb, "OR", "soft", dmin, dmax, K, from_logits=True)
        print(f"{name} fused vs two-step relerr {relerr(d2, d1):.2e}")
        del yb, pb


def stage_pool():
    import torch
    from acan_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    for B, N in ((2, 24), (16, 1408)):
        x = torch.relu(torch.randn(B, 2048, N, device=dev))
        w1 = torch.randn(512, 2048, device=dev) * 0.03
        b1 = torch.randn(512, device=dev) * 0.1
        w2 = torch.randn(512, 512, device=dev) * 0.06
        b2 = torch.randn(512, device=dev) * 0.1
        keep = (torch.rand(B, 2048, device=dev) > 0.5).float()
        for km in (None, keep):
            mean, hid, g = ops.pool_branch(x, w1, b1, w2, b2, km)
            m = x.double().mean(-1)
            if km is not None:
                m = m * km.double() * 2
            h = torch.relu(m @ w1.double().T + b1.double())
            gr = torch.relu(h @ w2.double().T + b2.double())
            print(f"pool B={B} N={N} mask={km is not None}: mean {relerr(mean, m):.2e} hid {relerr(hid, h):.2e} g {relerr(g, gr):.2e}")
        t = timeit(lambda: ops.pool_branch(x, w1, b1, w2, b2, None))
        print(f"   pool branch {t*1e6:.1f} us  {B*2048*N*4/t/1e9:.1f} GB/s (x read)")


def stage_kl():
    import torch
    from acan_b200 import ops
    from oracle import acan_oracle as O
    import synth
    dev = "cuda"
    for B, H, W in ((2, 32, 48), (3, 40, 56), (8, 256, 352)):
        n = (H // 8) * (W // 8)
        torch.manual_seed(1)
        sim = torch.softmax(2.5 * torch.randn(B, n, n), -1)
        label = synth.depth_label("probe.kl", B, H, W, 0.5, 10.5)
        dis = O.continuous2discrete(label, 0.72, 10.0, 80)
        loss, grad = ops.attention_loss(sim.to(dev), dis.to(dev), want_grad=True)
        bins = O.nearest_downsample8(dis).reshape(B, -1)
        gt = O.gt_affinity(bins.double())
        ref = O.kl_attention(sim.double(), gt)
        gref = O.attention_loss_grad_sim(sim.double(), gt)
        gtk = ops.gt_sim_map(dis.to(dev))
        print(f"kl B={B} N={n}: loss {float(loss):.6f} ref {float(ref):.6f} rel {abs(float(loss)-float(ref))/abs(float(ref)):.2e} "
              f"grad {relerr(grad.cpu(), gref):.2e} gt {relerr(gtk.cpu(), gt):.2e}")
        if B == 8:
            simd, disd = sim.to(dev), dis.to(dev)
            t = timeit(lambda: ops.attention_loss(simd, disd))
            print(f"   materialised KL {t*1e6:.1f} us  {B*n*n*4/t/1e9:.1f} GB/s")


def _attn_reference(f, v, fp16_operands):
    import torch
    b, dk = f.shape[:2]
    F = f.reshape(b, dk, -1).double()
    V = v.reshape(b, v.shape[1], -1).double() if v is not None else None
    if fp16_operands:
        F = f.reshape(b, dk, -1).half().double()
        V = v.reshape(b, v.shape[1], -1).half().double() if v is not None else None
    S = torch.matmul(F.transpose(1, 2), F) * dk ** -0.5
    P = torch.softmax(S, -1)
    if V is None:
        return S, P, None
    Pq = P.half().double() if fp16_operands else P
    return S, P, torch.matmul(V, Pq.transpose(1, 2))


def _diag(name, got, ref):
    import torch
    err = (got.double() - ref.double()).abs()
    print(f"   {name}: relerr {relerr(got, ref):.3e}  rel-L2 {float(err.norm() / ref.double().norm()):.3e}  nan {int(torch.isnan(got).sum())}")
    if relerr(got, ref) > 1e-2 or torch.isnan(got).any():
        e2 = err[0]
        r, c = e2.shape[-2:]
        print("      err by 8-row group :", [f"{float(e2[i:i+8].max()):.2e}" for i in range(0, min(r, 64), 8)])
        print("      err by 16-col group:", [f"{float(e2[:, i:i+16].max()):.2e}" for i in range(0, min(c, 128), 16)])
        print("      got[0,:4,:8]\n", got[0, :4, :8].cpu())
        print("      ref[0,:4,:8]\n", ref[0, :4, :8].cpu())


def _attn_case(B, h, w, dk, dv, label=False, time_it=False):
    import torch
    from acan_b200 import ops
    from oracle import acan_oracle as O
    import synth
    dev = "cuda"
    torch.manual_seed(B * 1000 + h * w)
    n = h * w
    # realistic post-BN/ReLU features: |F_i|^2 ~ dk/2 -> logits up to ~ sqrt(dk)/2
    f = torch.relu(torch.randn(B, dk, h, w, device=dev))
    v = torch.relu(torch.randn(B, dv, h, w, device=dev)) if dv else None
    dis = None
    if label:
        lab = synth.depth_label("probe.attn", B, h * 8, w * 8, 0.5, 10.5)
        dis = O.continuous2discrete(lab, 0.72, 10.0, 80).to(dev)
    out = ops.attention_forward(f, v, want_sim_map=True, dis_label=dis)
    torch.cuda.synchronize()
    S, P, Oref = _attn_reference(f, v, fp16_operands=True)
    Sx, Px, Ox = _attn_reference(f, v, fp16_operands=False)
    print(f"attn B={B} N={n} dk={dk} dv={dv}: logit max {float(Sx.max()):.2f}")
    _diag("sim_map vs fp16-operand fp64", out["sim_map"], P)
    _diag("sim_map vs exact fp64       ", out["sim_map"], Px)
    st = out["row_stats"]
    lse = (st[..., 0].double() + torch.log2(st[..., 1].double())) * 0.6931471805599453
    print(f"   row lse relerr {relerr(lse, torch.logsumexp(S, -1)):.3e}")
    if v is not None:
        _diag("ctx vs fp16-operand fp64    ", out["ctx"].reshape(B, dv, n), Oref)
        _diag("ctx vs exact fp64           ", out["ctx"].reshape(B, dv, n), Ox)
    if label:
        bins = O.nearest_downsample8(dis.cpu()).reshape(B, -1).double()
        gt = O.gt_affinity(bins)
        ref = float(O.kl_attention(Px.cpu(), gt))
        fused = float(out["loss"])
        again = float(ops.attention_loss_fused(out["ws"], dis))
        mat = float(ops.attention_loss(out["sim_map"], dis))
        print(f"   KL fused {fused:.6f}  fused(standalone) {again:.6f}  materialised {mat:.6f}  fp64 ref {ref:.6f}  rel {abs(fused-ref)/abs(ref):.2e}")
    rows = torch.tensor([n // 4, n // 2, 3 * n // 4])
    rr = ops.attention_rows(out["ws"], rows)
    print(f"   selected rows relerr {relerr(rr, Px[:, rows]):.3e}")
    if time_it:
        ws = out["ws"]
        t_all = timeit(lambda: ops.attention_forward(f, v, ws=ws), iters=10)
        t_aff = timeit(lambda: ops.attention_forward(f, None), iters=10)
        fl = B * 5120.0 * n * n if dv == 2048 and dk == 512 else B * 2.0 * n * n * (dk + dv)
        print(f"   attn fwd (pack+stats+probs+pv) {t_all*1e6:.1f} us -> {fl/t_all/1e12:.1f} TFLOP/s algorithmic; affinity-only call {t_aff*1e6:.1f} us")
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                ops.attention_forward(f, v, ws=ws)
            torch.cuda.synchronize()
        for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:10]:
            print(f"      {ev.key[:70]:70s} n={ev.count:3d} avg {ev.device_time_total/max(ev.count,1):9.1f} us")


def _attn_cl_case(B, h, w, dk=512, dv=2048, time_it=False):
    import torch
    from acan_b200 import ops
    dev = "cuda"
    torch.manual_seed(B * 1000 + h * w)
    n = h * w
    f = torch.relu(torch.randn(B, dk, h, w, device=dev)).half().contiguous(memory_format=torch.channels_last)
    v = torch.relu(torch.randn(B, dv, h, w, device=dev)).half().contiguous(memory_format=torch.channels_last)
    out = ops.attention_forward_cl(f, v, want_sim_map=True)
    torch.cuda.synchronize()
    S, P, Oref = _attn_reference(f.float(), v.float(), fp16_operands=False)
    print(f"attn channels-last B={B} N={n}: ctx is channels_last {out['ctx'].is_contiguous(memory_format=torch.channels_last)} dtype {out['ctx'].dtype}")
    _diag("sim_map vs fp64", out["sim_map"], P)
    _diag("ctx vs fp64    ", out["ctx"].float().reshape(B, dv, n), Oref)
    ref = ops.attention_forward(f.float(), v.float())
    _diag("ctx vs NCHW fp32-ABI kernel", out["ctx"].float().reshape(B, dv, n), ref["ctx"].reshape(B, dv, n).double())
    if time_it:
        import ctypes
        from acan_b200 import _lib
        ws = out["ws"]
        for _ in range(3):
            ops.attention_forward_cl(f, v, ws=ws)
        lib = _lib.load()
        lib.acan_profile_begin(200)
        for _ in range(20):
            ops.attention_forward_cl(f, v, ws=ws)
        ms = (ctypes.c_float * 7)(); cnt = (ctypes.c_int * 7)()
        lib.acan_profile_end(ms, cnt)
        us = [ms[k] / max(cnt[k], 1) * 1e3 for k in range(4)]
        print(f"   stats {us[1]:.1f} us probs {us[2]:.1f} us pv {us[3]:.1f} us total {sum(us):.1f} us -> {B*5120.0*n*n/sum(us)/1e6:.0f} TFLOP/s algorithmic")


def stage_attn_cl():
    _attn_cl_case(1, 8, 16, 64, 128)
    _attn_cl_case(2, 4, 6)
    _attn_cl_case(1, 28, 38)
    _attn_cl_case(2, 20, 80)
    _attn_cl_case(16, 32, 44, time_it=True)
    _attn_cl_case(32, 20, 64, time_it=True)


def stage_attn_small():
    _attn_case(1, 8, 16, 64, 128)        # N=128: one tile, one k-step
    _attn_case(1, 8, 16, 512, 128)       # full dk: 8 k-steps through the ring
    _attn_case(2, 4, 6, 512, 2048, label=True)   # N=24: masked tile, golden-sized


def stage_attn_shapes():
    _attn_case(1, 28, 38, 512, 2048, label=True)   # N=1064 reference NYU default (not a multiple of 128)
    _attn_case(2, 20, 80, 512, 2048, label=True)   # N=1600 reference KITTI default
    _attn_case(2, 32, 44, 512, 2048, label=True)   # N=1408


def stage_attn_full():
    _attn_case(16, 32, 44, 512, 2048, time_it=True)   # C2
    _attn_case(32, 20, 64, 512, 2048, time_it=True)   # C3


STAGES = {k[len("stage_"):]: v for k, v in list(globals().items()) if k.startswith("stage_")}

if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--run":
        STAGES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(STAGES)
    results = {}
    for s in names:
        print(f"===== {s}", flush=True)
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", s], timeout=600)
            results[s] = p.returncode
        except subprocess.TimeoutExpired:
            results[s] = "timeout"
        print(f"===== {s} -> {results[s]} ({time.time()-t0:.1f}s)", flush=True)
    print(json.dumps(results))
