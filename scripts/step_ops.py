#!/usr/bin/env python
"""aten-level view of the bench step (which non-convolution ops remain, with shapes): python scripts/step_ops.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from torch.profiler import profile, ProfilerActivity
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
net = bench.build_net(dev)
x = torch.rand(bench.BATCH, 3, bench.H, bench.W, device=dev).contiguous(memory_format=torch.channels_last)
for _ in range(5):
    bench.step_device(net, x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    bench.step_device(net, x)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.device_time_total > 0 and not e.key.startswith("cudnn") and "conv" not in e.key]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:25]:
    print(f"{e.device_time_total:9.1f} us n={e.count:3d} {e.key:40s} {str(e.input_shapes)[:150]}")
