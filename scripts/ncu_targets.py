#!/usr/bin/env python
"""One invocation of each hand-written kernel group at the BASELINE shapes inside an NVTX range ("ncu_capture"), after warm-up calls outside it:
    ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "ncu_capture/" -o gpurun_out/prof python scripts/ncu_targets.py
Groups: inference attention core (configs[1]), training attention forward / loss / backward (configs[3]), soft tail, pooling branch, BN fwd / bwd,
flat SGD."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from acan_b200 import _lib, ops

which = set(sys.argv[1:]) or {"infer", "train", "tail", "pool", "bn"}
dev = torch.device("cuda", 0)
cl = torch.channels_last
calls = []
if "infer" in which:
    f = torch.relu(torch.randn(16, 512, 32, 44, device=dev)).half().contiguous(memory_format=cl)
    v = torch.relu(torch.randn(16, 2048, 32, 44, device=dev)).half().contiguous(memory_format=cl)
    ws = ops.attention_forward_cl(f, v)["ws"]
    calls.append(lambda: ops.attention_forward_cl(f, v, ws=ws))
if "train" in which:
    ft = torch.relu(torch.randn(8, 512, 32, 44, device=dev)).half().contiguous(memory_format=cl)
    vt = torch.relu(torch.randn(8, 2048, 32, 44, device=dev)).bfloat16().contiguous(memory_format=cl)
    lab = torch.randint(0, 80, (8, 1, 256, 352), device=dev).float()
    go = (torch.randn(8, 2048, 32, 44, device=dev) * 1e-3).bfloat16().contiguous(memory_format=cl)
    gl = torch.ones((), device=dev)
    o = ops.attention_forward_train(ft, vt)
    ops.attention_loss_fused(o["ws"], lab)
    def train_calls():
        oo = ops.attention_forward_train(ft, vt)
        ops.attention_loss_fused(oo["ws"], lab)
        ops.attention_backward_cl(oo["ws"], oo["value16"], oo["ctx32"], go, gl, torch.float16, torch.bfloat16)
    calls.append(train_calls)
if "tail" in which:
    z = torch.randn(16, 160, 32, 44, device=dev).permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    calls.append(lambda: ops.tail_inference(z, "OR", "soft", 0.72, 10.0, 80))
if "pool" in which:
    x = torch.relu(torch.randn(16, 2048, 32, 44, device=dev)).half().contiguous(memory_format=cl)
    w1, b1, w2, b2 = torch.randn(512, 2048, device=dev), torch.randn(512, device=dev), torch.randn(512, 512, device=dev), torch.randn(512, device=dev)
    calls.append(lambda: ops.pool_branch_cl(x, w1, b1, w2, b2))
if "bn" in which:
    xb = torch.randn(8, 1024, 32, 44, device=dev, dtype=torch.bfloat16).contiguous(memory_format=cl)
    gb = torch.randn_like(xb)
    wgt, bia = torch.ones(1024, device=dev), torch.zeros(1024, device=dev)
    def bn_calls():
        y, mean, invstd = ops.bn_train_forward(xb, wgt, bia, None, None, 0.1, 1e-5, True, None)
        ops.bn_train_backward(xb, y, gb, wgt, mean, invstd, False)
    calls.append(bn_calls)
for c in calls:
    for _ in range(3):
        c()
torch.cuda.synchronize()
with torch.cuda.nvtx.range("ncu_capture"):
    for c in calls:
        c()
torch.cuda.synchronize()
print("ok")
