#!/usr/bin/env python
"""C4 timing: attention fwd+bwd kernels and the full ACAN training step (fwd + loss(beta=1) + bwd), B=8, 256x352."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from acan_b200 import ops, functions as fn
from acan_b200.modules import _Holder, AttentionLoss2d
from acan_b200.network import ResNet, OrdinalRegression2d
torch.backends.cudnn.benchmark = True
dev = "cuda"
B, H, W = 8, 256, 352
h, w = H // 8, W // 8
n = h * w

def ev_time(fn_, iters=10, warm=3):
    for _ in range(warm): fn_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn_()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3

torch.manual_seed(0)
f = torch.relu(torch.randn(B, 512, h, w, device=dev)); v = torch.relu(torch.randn(B, 2048, h, w, device=dev))
lab = torch.randint(0, 80, (B, 1, H, W), device=dev).float()
go = torch.randn(B, 2048, h, w, device=dev) * 1e-3
out = ops.attention_forward(f, v, dis_label=lab)
t_f = ev_time(lambda: ops.attention_forward(f, v, dis_label=lab, ws=out["ws"]))
t_b = ev_time(lambda: ops.attention_backward(f, out["row_stats"], v, out["ctx"], go, lab, 1.0))
print(f"attention fwd+loss (fp32 ABI, incl. pack) {t_f*1e6:.0f} us = {B*5120.0*n*n/t_f/1e12:.0f} TFLOP/s alg; bwd {t_b*1e6:.0f} us = {B*10240.0*n*n/t_b/1e12:.0f} TFLOP/s alg (10240 N^2)")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): ops.attention_backward(f, out["row_stats"], v, out["ctx"], go, lab, 1.0)
    torch.cuda.synchronize()
for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:14]:
    print(f"   {ev.key[:100]:100s} n={ev.count//3:2d} avg {ev.device_time_total/max(ev.count,1):8.1f} us")
# full training step
net = ResNet(0.72, 10.0, 80, "OR", "soft", "attention", H, W, alpha=0, beta=1.0, layers=[3, 4, 23, 3]).to(dev).train()
crit = net.LossFunc(0.72, 10.0, 80, AppearanceLoss=OrdinalRegression2d(ignore_index=0), AttentionLoss=AttentionLoss2d(scale=1), alpha=0.0, beta=1.0)
opt = torch.optim.SGD(net.parameters(), lr=1e-4, momentum=0.9)
x = torch.rand(B, 3, H, W, device=dev); y = torch.rand(B, 1, H, W, device=dev) * 10 + 0.5
def step():
    opt.zero_grad(set_to_none=True)
    o = net(x)
    l1, _, l3, tot = crit(o, y, 1)
    tot.backward()
    opt.step()
t = ev_time(step, iters=5, warm=3)
print(f"full training step (fp32/TF32 convs, beta=1), B={B}: {t*1e3:.1f} ms -> {B/t:.1f} images/s")

# the same step with 16-bit autocast convolutions + channels_last (the attention / loss / head kernels are unchanged)
net = net.to(memory_format=torch.channels_last)
xc = x.contiguous(memory_format=torch.channels_last)
def step_amp():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        o = net(xc)
        l1, _, l3, tot = crit(o, y, 1)
    tot.backward()
    opt.step()
t = ev_time(step_amp, iters=5, warm=3)
print(f"full training step (bf16 autocast + channels_last convs, beta=1), B={B}: {t*1e3:.1f} ms -> {B/t:.1f} images/s")
