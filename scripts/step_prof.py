#!/usr/bin/env python
"""torch-profiler kernel table of the bench step:  python scripts/step_prof.py [fp16|bf16] [topN]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from torch.profiler import profile, ProfilerActivity
bench.AMP = torch.float16 if (len(sys.argv) < 2 or sys.argv[1] == "fp16") else torch.bfloat16
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
net = bench.build_net(dev)
x = torch.rand(bench.BATCH, 3, bench.H, bench.W, device=dev).contiguous(memory_format=torch.channels_last)
for _ in range(5):
    bench.step_device(net, x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        bench.step_device(net, x)
    torch.cuda.synchronize()
evs = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in evs)
print(f"AMP={bench.AMP} total kernel time per step {tot/3/1e3:.2f} ms")
for ev in evs[:top]:
    print(f"{100*ev.device_time_total/tot:6.2f}% {ev.device_time_total/3:9.1f} us/step n={ev.count//3:4d}  {ev.key[:120]}")
