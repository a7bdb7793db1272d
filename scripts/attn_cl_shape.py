#!/usr/bin/env python
"""Short run of the channels-last fp16 attention path (the one the inference pipeline takes) for ncu: python scripts/attn_cl_shape.py B h w"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acan_b200 import ops
B, h, w = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (16, 32, 44)
torch.manual_seed(0)
f = torch.relu(torch.randn(B, 512, h, w, device="cuda")).half().contiguous(memory_format=torch.channels_last)
v = torch.relu(torch.randn(B, 2048, h, w, device="cuda")).half().contiguous(memory_format=torch.channels_last)
ws = None
for _ in range(3):
    out = ops.attention_forward_cl(f, v, ws=ws)
    ws = out["ws"]
torch.cuda.synchronize()
print("ok", float(out["ctx"].float().abs().mean()))
