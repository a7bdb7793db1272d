#!/usr/bin/env python
"""Per-shape timing of the fused BN kernels vs the framework's BN (+ReLU) on the C4 training shapes:
   python scripts/bn_time.py   -> us per call (CUDA events), effective GB/s on the algorithmic bytes"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acan_b200 import functions as fn, ops
dev = "cuda"
shapes = [(8, 64, 128, 176), (8, 64, 64, 88), (8, 256, 64, 88), (8, 128, 32, 44), (8, 512, 32, 44), (8, 256, 32, 44), (8, 1024, 32, 44), (8, 2048, 32, 44)]
dt = torch.bfloat16

def ev(f, n=30):
    for _ in range(3): f()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

def graph_time(calls, reps=10):
    """per-call us of a list of callables replayed from one CUDA graph (no host launch overhead in the number)"""
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in calls[:2]: f()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        keep = [f() for f in calls]
    for _ in range(2): g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * len(calls)) * 1e3

for (b, c, h, w) in shapes:
    nrot = max(4, int(300e6 // (b * c * h * w * 2)))            # rotate over > 2x L2 worth of tensors: HBM-resident operands
    xs = [torch.randn(b, c, h, w, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last) for _ in range(nrot)]
    gs = [torch.randn(b, c, h, w, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last) for _ in range(nrot)]
    bn = torch.nn.BatchNorm2d(c).to(dev).train()
    wgt, bia = bn.weight.detach(), bn.bias.detach()
    outs = [ops.bn_train_forward(x, wgt, bia, bn.running_mean, bn.running_var, 0.1, 1e-5, True, None) for x in xs]
    ys, mean, invstd = [o[0] for o in outs], outs[0][1], outs[0][2]
    t_f = graph_time([lambda x=x: ops.bn_train_forward(x, wgt, bia, bn.running_mean, bn.running_var, 0.1, 1e-5, True, None) for x in xs])
    t_b = graph_time([lambda k=k: ops.bn_train_backward(xs[k], ys[k], gs[k], wgt, mean, invstd, False) for k in range(nrot)])
    t_tf = graph_time([lambda x=x: torch.relu(torch.nn.functional.batch_norm(x, bn.running_mean, bn.running_var, wgt, bia, True, 0.1, 1e-5)) for x in xs])
    mb = b * c * h * w * 2 / 1e6
    print(f"[{b},{c},{h},{w}] {mb:6.1f} MB  fused fwd {t_f:6.1f} us ({3*mb/t_f:5.2f} TB/s on 3 passes)  fused bwd {t_b:6.1f} us ({7*mb/t_b:5.2f} TB/s on 7 passes)"
          f"  torch bn+relu fwd {t_tf:6.1f} us", flush=True)
from torch.profiler import profile, ProfilerActivity
b, c, h, w = 8, 256, 32, 44
x = torch.randn(b, c, h, w, device=dev, dtype=dt).contiguous(memory_format=torch.channels_last)
g = torch.randn_like(x)
bn = torch.nn.BatchNorm2d(c).to(dev).train()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        y, mean, invstd = ops.bn_train_forward(x, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, 0.1, 1e-5, True, None)
        ops.bn_train_backward(x, y, g, bn.weight.detach(), mean, invstd, False)
    torch.cuda.synchronize()
for e_ in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    print(f"   {e_.key[:90]:90s} n={e_.count} avg {e_.device_time_total/max(e_.count,1):7.1f} us")
