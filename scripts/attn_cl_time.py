#!/usr/bin/env python
"""Per-pass and per-kernel timing of the channels-last fp16 attention path (the one the inference pipeline takes):
    python scripts/attn_cl_time.py B h w [iters] [--kernels]
Per-pass numbers come from the library's CUDA-event hooks, the per-kernel table from the torch profiler (device times)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acan_b200 import ops, _lib

args = [a for a in sys.argv[1:] if not a.startswith("--")]
B, h, w = (int(a) for a in args[:3]) if len(args) >= 3 else (16, 32, 44)
iters = int(args[3]) if len(args) > 3 else 20
torch.manual_seed(0)
f = torch.relu(torch.randn(B, 512, h, w, device="cuda")).half().contiguous(memory_format=torch.channels_last)
v = torch.relu(torch.randn(B, 2048, h, w, device="cuda")).half().contiguous(memory_format=torch.channels_last)
ws = ops.attention_forward_cl(f, v)["ws"]
for _ in range(3):
    ops.attention_forward_cl(f, v, ws=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    ops.attention_forward_cl(f, v, ws=ws)
e1.record(); torch.cuda.synchronize()
whole = e0.elapsed_time(e1) / iters * 1e3
lib = _lib.load()
lib.acan_profile_begin(iters * 8)
for _ in range(iters):
    ops.attention_forward_cl(f, v, ws=ws)
ms = (ctypes.c_float * 7)(); cnt = (ctypes.c_int * 7)()
lib.acan_profile_end(ms, cnt)
n = h * w
names = ["pack", "stats", "probs", "pv"]
us = [ms[k] / max(cnt[k], 1) * 1e3 for k in range(4)]
print(f"B={B} N={n}", " ".join(f"{a}={b:.1f}us" for a, b in zip(names, us)), f"sum={sum(us):.1f}us call={whole:.1f}us "
      f"core(call)={B*5120.0*n*n/whole/1e6:.0f}TF/s pv={B*4096.0*n*n/us[3]/1e6:.0f}TF/s probs={B*1024.0*n*n/us[2]/1e6:.0f}TF/s",
      {k: os.environ[k] for k in os.environ if k.startswith("ACAN_")}, flush=True)
if "--kernels" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            ops.attention_forward_cl(f, v, ws=ws)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    per = len(evs) // 5
    t0 = evs[-per].time_range.start
    for e in evs[-per:]:
        print(f"   +{e.time_range.start - t0:8.1f} us  dur {e.time_range.end - e.time_range.start:7.1f} us  {e.name[:110]}")
