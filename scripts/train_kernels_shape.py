#!/usr/bin/env python
"""Short run of the training-path bandwidth kernels at the C4 shapes, for ncu:  python scripts/train_kernels_shape.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acan_b200 import ops, functions as fn
dev = "cuda"
torch.manual_seed(0)
b, c, h, w = 8, 1024, 32, 44
x = torch.randn(b, c, h, w, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
g = torch.randn_like(x)
bn = torch.nn.BatchNorm2d(c).to(dev).train()
z = (2 * torch.randn(8, 160, 32, 44, device=dev)).requires_grad_(True)
lab = torch.randint(0, 80, (8, 256, 352), device=dev)
for _ in range(3):
    y, mean, invstd = ops.bn_train_forward(x, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, 0.1, 1e-5, True, None)
    ops.bn_train_backward(x, y, g, bn.weight.detach(), mean, invstd, False)
    z.grad = None
    fn.FusedTailLoss.apply(z, lab, 0).backward()
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()), float(z.grad.abs().mean()))
