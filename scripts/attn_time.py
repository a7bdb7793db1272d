#!/usr/bin/env python
"""Per-pass timing of the attention core via the library's event hooks:  python scripts/attn_time.py B h w [iters]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acan_b200 import ops, _lib
B, h, w = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (16, 32, 44)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
torch.manual_seed(0)
f = torch.relu(torch.randn(B, 512, h, w, device="cuda"))
v = torch.relu(torch.randn(B, 2048, h, w, device="cuda"))
ws = ops.attention_forward(f, v)["ws"]
for _ in range(3):
    ops.attention_forward(f, v, ws=ws)
lib = _lib.load()
lib.acan_profile_begin(iters * 8)
for _ in range(iters):
    ops.attention_forward(f, v, ws=ws)
ms = (ctypes.c_float * 7)(); cnt = (ctypes.c_int * 7)()
lib.acan_profile_end(ms, cnt)
n = h * w
names = ["pack", "stats", "probs", "pv"]
us = [ms[k] / max(cnt[k], 1) * 1e3 for k in range(4)]
print(" ".join(f"{a}={b:.1f}us" for a, b in zip(names, us)), f"total={sum(us):.1f}us core={B*5120.0*n*n/sum(us)/1e6:.0f}TF/s pv={B*4096.0*n*n/us[3]/1e6:.0f}TF/s",
      {k: os.environ[k] for k in os.environ if k.startswith("ACAN_")})
