#!/usr/bin/env python
"""torch-profiler kernel table of the C4 training step (fwd + loss(beta=1) + bwd + SGD), B=8, 256x352:
    python scripts/train_prof.py [fp32|amp] [topN]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torch.profiler import profile, ProfilerActivity
from acan_b200.modules import AttentionLoss2d
from acan_b200.network import ResNet, OrdinalRegression2d
mode = sys.argv[1] if len(sys.argv) > 1 else "amp"
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
torch.backends.cudnn.benchmark = True
dev = "cuda"
B, H, W = 8, 256, 352
net = ResNet(0.72, 10.0, 80, "OR", "soft", "attention", H, W, alpha=0, beta=1.0, layers=[3, 4, 23, 3]).to(dev).train()
crit = net.LossFunc(0.72, 10.0, 80, AppearanceLoss=OrdinalRegression2d(ignore_index=0), AttentionLoss=AttentionLoss2d(scale=1), alpha=0.0, beta=1.0)
opt = torch.optim.SGD(net.parameters(), lr=1e-4, momentum=0.9)
x = torch.rand(B, 3, H, W, device=dev); y = torch.rand(B, 1, H, W, device=dev) * 10 + 0.5
if mode == "amp":
    net = net.to(memory_format=torch.channels_last)
    x = x.contiguous(memory_format=torch.channels_last)
def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "amp")):
        o = net(x)
        l1, _, l3, tot = crit(o, y, 1)
    tot.backward()
    opt.step()
for _ in range(4):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record(); torch.cuda.synchronize()
print(f"mode={mode} step {e0.elapsed_time(e1)/5:.2f} ms -> {B/(e0.elapsed_time(e1)/5e3):.1f} images/s")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
evs = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in evs)
print(f"total kernel time per step {tot/3/1e3:.2f} ms")
for ev in evs[:top]:
    print(f"{100*ev.device_time_total/tot:6.2f}% {ev.device_time_total/3:9.1f} us/step n={ev.count//3:4d}  {ev.key[:130]}")
