#!/usr/bin/env python
"""2-GPU check of the replayed training step: after a few GraphedTrainStep calls on different shards the replicas' parameters
must be identical (averaged gradients), and equal to a single-process run that averages the two shards' gradients itself.
    torchrun --nproc-per-node 2 scripts/ddp_graph_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
from acan_b200.parallel import init_distributed, GradAllReducer
from acan_b200.training import GraphedTrainStep
from acan_b200.network import ResNet, OrdinalRegression2d
from acan_b200.modules import AttentionLoss2d

rank, world, local = init_distributed("nccl")
dev = torch.device("cuda", local)
H, W, B = 64, 96, 4
torch.manual_seed(11)
net = ResNet(0.72, 10.0, 80, "OR", "soft", "attention", H, W, alpha=0, beta=1.0, layers=[1, 1, 1, 1]).to(dev).to(memory_format=torch.channels_last).train()
for m in net.modules():
    if isinstance(m, torch.nn.Dropout2d): m.eval()
crit = net.LossFunc(0.72, 10.0, 80, AppearanceLoss=OrdinalRegression2d(ignore_index=0), AttentionLoss=AttentionLoss2d(scale=1), alpha=0.0, beta=1.0)
opt = torch.optim.SGD(net.parameters(), lr=1e-3, momentum=0.9)
torch.manual_seed(100 + rank)
xs = [torch.rand(B, 3, H, W, device=dev).contiguous(memory_format=torch.channels_last) for _ in range(4)]
ys = [torch.rand(B, 1, H, W, device=dev) * 10 + 0.5 for _ in range(4)]
red = GradAllReducer(net.parameters())
mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
step = GraphedTrainStep(net, crit, opt, xs[0], ys[0], amp_dtype=torch.bfloat16, reducer=red, warmup=2, graph=(mode != "eager"),
                        shadow_weights=(mode != "noshadow"))
def probe(tag):
    torch.cuda.synchronize()
    msg = []
    for name in ("decoder.saconv.f_key.conv.weight", "layer4.0.conv1.weight", "decoder.saconv.f_key.bn.weight"):
        p = dict(net.named_parameters())[name]
        for what, t in (("w", p.detach()), ("g", p.grad.detach() if p.grad is not None else None)):
            if t is None:
                msg.append(f"{name.split('.')[-3]}.{what}=None"); continue
            ref = t.clone(); dist.broadcast(ref, 0)
            msg.append(f"{name.split('.')[-3]}.{what}:{float((t - ref).abs().max()):.1e}")
    if rank == 1: print(tag, " ".join(msg), flush=True)
probe("after construction")
for i, (x, y) in enumerate(zip(xs, ys)):
    loss = step(x, y)
    probe(f"after step {i}")
torch.cuda.synchronize()
worst, name_w, diffs = 0.0, "", []
for name, p in net.named_parameters():
    ref = p.detach().clone()
    dist.broadcast(ref, 0)
    d = float((p.detach() - ref).abs().max())
    diffs.append((d / (float(ref.abs().max()) + 1e-30), name))
    if d > worst: worst, name_w = d, name
if rank == 1:
    diffs.sort(reverse=True)
    print(mode, "params that differ from rank 0:", sum(1 for d, _ in diffs if d > 0), "of", len(diffs), "; top:", [(f"{d:.1e}", n) for d, n in diffs[:6]], flush=True)
flat_is_view = step._flat_views is not None and len(step._flat_views) == len([p for p in net.parameters() if p.grad is not None]) if mode != 'eager' else None
print(f"rank {rank}: loss {float(loss):.4f} worst |p - p_rank0| = {worst:.3e} ({name_w}); every gradient has a slot in the flat buffer: {flat_is_view}", flush=True)
dist.barrier(); dist.destroy_process_group()
