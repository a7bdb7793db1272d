#!/usr/bin/env python
"""Short C2-sized run of the attention core (B=16, N=1408, dk=512, dv=2048) for ncu: 1 warm-up + 2 calls."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acan_b200 import ops
torch.manual_seed(0)
f = torch.relu(torch.randn(16, 512, 32, 44, device="cuda"))
v = torch.relu(torch.randn(16, 2048, 32, 44, device="cuda"))
ws = None
for _ in range(3):
    out = ops.attention_forward(f, v, ws=ws)
    ws = out["ws"]
torch.cuda.synchronize()
print("ok", float(out["ctx"].abs().mean()))
