#!/usr/bin/env python
"""stem convolution (7x7 / 2, 3 -> 64 channels, fp16 channels_last, fused bias + ReLU) at configs[1]: 3 input channels vs zero-padded to 4 / 8."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
torch.backends.cudnn.benchmark = True
dev = "cuda"
def ev(fn, iters=20, warm=5):
    for _ in range(warm): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / iters
x32 = torch.rand(16, 3, 256, 352, device=dev).contiguous(memory_format=torch.channels_last)
w = torch.randn(64, 3, 7, 7, device=dev) * 0.05
bias = torch.randn(64, device=dev).half()
for cin in (3, 4, 8):
    wp = torch.nn.functional.pad(w, (0, 0, 0, 0, 0, cin - 3)).half().contiguous(memory_format=torch.channels_last)
    def prep():
        if cin == 3:
            return x32.to(torch.float16).contiguous(memory_format=torch.channels_last)
        xp = torch.zeros((16, cin, 256, 352), device=dev, dtype=torch.float16).contiguous(memory_format=torch.channels_last)
        xp[:, :3].copy_(x32)
        return xp
    xin = prep()
    t_prep = ev(prep)
    t_conv = ev(lambda: torch.cudnn_convolution_relu(xin, wp, bias, (2, 2), (3, 3), (1, 1), 1))
    print(f"C_in={cin}: input prep {t_prep:6.1f} us, conv+bias+relu {t_conv:6.1f} us, total {t_prep + t_conv:6.1f} us", flush=True)
