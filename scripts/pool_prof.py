import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acan_b200 import ops
from torch.profiler import profile, ProfilerActivity
x = torch.relu(torch.randn(16, 2048, 1408, device="cuda"))
w1 = torch.randn(512, 2048, device="cuda") * 0.03; b1 = torch.randn(512, device="cuda")
w2 = torch.randn(512, 512, device="cuda") * 0.06; b2 = torch.randn(512, device="cuda")
sim = torch.softmax(torch.randn(8, 1408, 1408, device="cuda"), -1); dis = torch.randint(0, 80, (8, 1, 256, 352), device="cuda").float()
for _ in range(3):
    ops.pool_branch(x, w1, b1, w2, b2, None); ops.attention_loss(sim, dis)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        ops.pool_branch(x, w1, b1, w2, b2, None); ops.attention_loss(sim, dis)
    torch.cuda.synchronize()
for ev in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
    print(f"{ev.key[:90]:90s} n={ev.count:3d} avg {ev.device_time_total/max(ev.count,1):9.1f} us")
