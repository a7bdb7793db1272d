#!/usr/bin/env python
"""Stand-alone timings of the hand-written kernels at the BASELINE shapes (CUDA events, cool GPU: run this first in a gpurun call).
Prints one line per measurement with the SM clock read right after it; used for A/B between kernel versions (profiles/r02_notes.md)."""
import ctypes, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from acan_b200 import _lib, ops

dev = torch.device("cuda", 0)
lib = _lib.load()
try:
    import pynvml
    pynvml.nvmlInit()
    _h = pynvml.nvmlDeviceGetHandleByIndex(0)
    clk = lambda: pynvml.nvmlDeviceGetClockInfo(_h, pynvml.NVML_CLOCK_SM)
except Exception:
    clk = lambda: -1
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}

def ev(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters

def line(name, us, flops=None, extra=""):
    s = f"{name:52s} {us:9.1f} us  sm {clk()} MHz"
    if flops: s += f"  {flops / us / 1e6:7.0f} TFLOP/s = {flops / us / 1e6 / PEAK['bf16_tflops']:.3f} of burst peak"
    print(s + ("  " + extra if extra else ""), flush=True)

cl = torch.channels_last
def feats(b, h, w, dtype_v=torch.float16):
    f = torch.relu(torch.randn(b, 512, h, w, device=dev)).half().contiguous(memory_format=cl)
    v = torch.relu(torch.randn(b, 2048, h, w, device=dev)).to(dtype_v).contiguous(memory_format=cl)
    return f, v

for tag, (b, h, w) in (("C2 B16 N1408", (16, 32, 44)), ("C3 B32 N1280", (32, 20, 64)), ("B16 N1064", (16, 28, 38))):
    f, v = feats(b, h, w)
    n = h * w
    ws = ops.attention_forward_cl(f, v)["ws"]
    us = ev(lambda: ops.attention_forward_cl(f, v, ws=ws))
    line(f"inference core {tag} (default: PDL-chained launches)", us, b * 5120.0 * n * n)
    us2 = ev(lambda: ops.attention_forward_cl(f, v, ws=ws, fused=True))
    line(f"inference core {tag} (single fused launch, opt-in)", us2, b * 5120.0 * n * n)
    _lib.check(lib.acan_profile_begin(256), "pb")
    for _ in range(10): ops.attention_forward_cl(f, v, ws=ws, fused=False)
    torch.cuda.synchronize()
    ms = (ctypes.c_float * 8)(); cnt = (ctypes.c_int * 8)()
    _lib.check(lib.acan_profile_end(ms, cnt), "pe")
    print("    per pass (event brackets, no PDL overlap): " + ", ".join(f"{nm} {ms[i] / max(cnt[i], 1) * 1e3:.1f}" for i, nm in ((1, "diag+stats"), (2, "probs"), (3, "pv"))), flush=True)
    del f, v, ws

# training path, configs[3] shape
b, h, w = 8, 32, 44
n = h * w
f, v = feats(b, h, w, torch.bfloat16)
lab = torch.randint(0, 80, (b, 1, 8 * h, 8 * w), device=dev).float()
go16 = (torch.randn(b, 2048, h, w, device=dev) * 1e-3).bfloat16().contiguous(memory_format=cl)
goh = go16.half()
gl = torch.ones((), device=dev)
o = ops.attention_forward_train(f, v)
ops.attention_loss_fused(o["ws"], lab)
line("train forward C4 (keep=1, fp32 ctx)", ev(lambda: ops.attention_forward_train(f, v)), b * 5120.0 * n * n)
_lib.check(lib.acan_profile_begin(256), "pb")
for _ in range(10): ops.attention_forward_train(f, v)
torch.cuda.synchronize()
ms = (ctypes.c_float * 8)(); cnt = (ctypes.c_int * 8)()
_lib.check(lib.acan_profile_end(ms, cnt), "pe")
print("    per pass: " + ", ".join(f"{nm} {ms[i] / max(cnt[i], 1) * 1e3:.1f}" for i, nm in ((1, "stats"), (2, "probs"), (3, "pv32"))), flush=True)
line("attention loss C4 (fused KL over the workspace)", ev(lambda: ops.attention_loss_fused(o["ws"], lab)))
line("train backward C4, bf16 dO (packed), with loss", ev(lambda: ops.attention_backward_cl(o["ws"], o["value16"], o["ctx32"], go16, gl, torch.float16, torch.bfloat16)), b * 10240.0 * n * n)
line("train backward C4, fp16 dO (in place), with loss", ev(lambda: ops.attention_backward_cl(o["ws"], o["value16"], o["ctx32"], goh, gl, torch.float16, torch.float16)), b * 10240.0 * n * n)
line("train backward C4, fp16 dO, no loss", ev(lambda: ops.attention_backward_cl(o["ws"], o["value16"], o["ctx32"], goh, None, torch.float16, torch.float16)), b * 10240.0 * n * n)
del f, v, o, go16, goh

# tails
z = torch.randn(16, 160, 32, 44, device=dev)
line("fused tail OR soft, C2 (B16 256x352)", ev(lambda: ops.tail_inference(z, "OR", "soft", 0.72, 10.0, 80)))
z3 = torch.randn(32, 160, 20, 64, device=dev)
line("fused tail OR hard, C3 (B32 160x512)", ev(lambda: ops.tail_inference(z3, "OR", "hard", 1.98, 80.0, 80)))
x = torch.relu(torch.randn(16, 2048, 32, 44, device=dev)).half().contiguous(memory_format=cl)
w1, b1, w2, b2 = torch.randn(512, 2048, device=dev), torch.randn(512, device=dev), torch.randn(512, 512, device=dev), torch.randn(512, device=dev)
us = ev(lambda: ops.pool_branch_cl(x, w1, b1, w2, b2))
line("pooling branch C2 (channels-last fp16)", us, extra=f"{x.numel() * 2 / us / 1e3:.0f} GB/s = {x.numel() * 2 / us / 1e3 / PEAK['hbm_gbs']:.2f} of HBM")
