#!/usr/bin/env python
"""aten-level view of the configs[3] training step as the bench runs it (GraphedTrainStep, eager launches so the profiler sees the ops): which
non-convolution, non-acan ops remain, with shapes, and the kernel table.   python scripts/train_ops.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torch.profiler import profile, ProfilerActivity
from acan_b200.modules import AttentionLoss2d
from acan_b200.network import ResNet, OrdinalRegression2d
from acan_b200.training import GraphedTrainStep
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
B, H, W, K = 8, 256, 352, 80
torch.manual_seed(2019)
net = ResNet(0.72, 10.0, K, "OR", "soft", "attention", H, W, alpha=0, beta=1.0, layers=[3, 4, 23, 3]).to(dev).to(memory_format=torch.channels_last).train()
crit = net.LossFunc(0.72, 10.0, K, AppearanceLoss=OrdinalRegression2d(ignore_index=0), AttentionLoss=AttentionLoss2d(scale=1), alpha=0.0, beta=1.0)
opt = torch.optim.SGD(net.parameters(), lr=1e-4, momentum=0.9, weight_decay=5e-4)
x = torch.rand(B, 3, H, W, device=dev).contiguous(memory_format=torch.channels_last)
y = (torch.rand(B, 1, H, W, device=dev) * 10 + 0.5)
step = GraphedTrainStep(net, crit, opt, x, y, amp_dtype=torch.bfloat16, graph=False)
for _ in range(4):
    step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    step(x, y)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.device_time_total > 0]
tot = sum(e.self_device_time_total for e in rows)
print(f"kernel time of one eager step: {tot / 1e3:.2f} ms")
by_op = {}
for e in rows:
    by_op.setdefault(e.key, [0.0, 0])
    by_op[e.key][0] += e.self_device_time_total
    by_op[e.key][1] += e.count
print("--- by op (self device time)")
for k, (t, n) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"{100 * t / tot:6.2f}% {t:9.1f} us n={n:4d}  {k[:90]}")
print("--- non-convolution ops by shape")
rows2 = [e for e in rows if "conv" not in e.key and "cudnn" not in e.key and e.self_device_time_total > 15]
rows2.sort(key=lambda e: -e.self_device_time_total)
for e in rows2[:40]:
    print(f"{e.self_device_time_total:9.1f} us n={e.count:3d} {e.key:38s} {str(e.input_shapes)[:140]}")
