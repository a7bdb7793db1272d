#!/usr/bin/env python
"""Timeline of the split training step on N GPUs (torchrun): where does the all-reduce of the late layers' gradients run relative to the
second backward graph?  CUDA events on the compute stream and on a communication stream the collective is issued from.
    torchrun --nnodes=1 --nproc-per-node 2 scripts/ddp_overlap_probe.py [--channels N]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
from acan_b200.parallel import init_distributed, GradAllReducer, capped_group
from acan_b200.network import ResNet, OrdinalRegression2d
from acan_b200.modules import AttentionLoss2d
from acan_b200.training import GraphedTrainStep

NCCL_SMS = int(os.environ.get("PROBE_NCCL_SMS", "16"))
rank, world, local = init_distributed("nccl")
late_group = capped_group(NCCL_SMS)
dev = torch.device("cuda", local)
torch.backends.cudnn.benchmark = True
H, W, B, K = 256, 352, 8, 80
torch.manual_seed(2019)
net = ResNet(0.72, 10.0, K, "OR", "soft", "attention", H, W, alpha=0, beta=1.0, layers=[3, 4, 23, 3]).to(dev).to(memory_format=torch.channels_last).train()
crit = net.LossFunc(0.72, 10.0, K, AppearanceLoss=OrdinalRegression2d(ignore_index=0), AttentionLoss=AttentionLoss2d(scale=1), alpha=0.0, beta=1.0)
opt = torch.optim.SGD(net.parameters(), lr=1e-4, momentum=0.9, weight_decay=5e-4)
red = GradAllReducer(net.parameters())
x = torch.rand(B, 3, H, W, device=dev).contiguous(memory_format=torch.channels_last)
y = (torch.rand(B, 1, H, W, device=dev) * 10 + 0.5)
step = GraphedTrainStep(net, crit, opt, x, y, amp_dtype=torch.bfloat16, reducer=red, overlap=True, nccl_sms=NCCL_SMS, overlap_group=late_group)
assert step._split
for _ in range(5):
    step(x, y)
torch.cuda.synchronize(); dist.barrier()
main = torch.cuda.current_stream(dev)
comm = torch.cuda.Stream(dev, priority=-1)
ev = lambda: torch.cuda.Event(enable_timing=True)
# the graphs back to back with no collective in between: what the split itself costs
for _ in range(3):
    evs = [ev() for _ in range(len(step.graphs) + 1)]
    evs[0].record(main)
    for k, g in enumerate(step.graphs):
        g.replay(); evs[k + 1].record(main)
    torch.cuda.synchronize()
    if rank == 0:
        print("no collective: " + "  ".join(f"graph {k + 1} {evs[k].elapsed_time(evs[k + 1]):.3f} ms" for k in range(len(step.graphs)))
              + f"  all {evs[0].elapsed_time(evs[-1]):.3f} ms", flush=True)
dist.barrier()
stages = step._piece_ranges()
rows = []
for it in range(8):
    t0 = ev(); t0.record(main)
    marks = []
    for k, g in enumerate(step.graphs):
        g.replay()
        eg = ev(); eg.record(main)
        comm.wait_stream(main)
        grp = late_group if k + 1 < len(step.graphs) else None
        with torch.cuda.stream(comm):
            es = ev(); es.record(comm)
            for (lo, hi), _ in stages[k]:
                dist.all_reduce(step._flat[lo:hi], op=dist.ReduceOp.AVG, group=grp)
            ee = ev(); ee.record(comm)
        marks.append((eg, es, ee))
    main.wait_stream(comm)
    step._sgd_range(0, step._n_sgd_chunks)
    eo = ev(); eo.record(main)
    torch.cuda.synchronize()
    rows.append([(t0.elapsed_time(a), t0.elapsed_time(b), t0.elapsed_time(c)) for a, b, c in marks] + [t0.elapsed_time(eo)])
if rank == 0:
    for r in rows[2:]:
        print("ms from step start: " + "  ".join(f"graph {k + 1} ends {a:7.3f} its all-reduce {b:7.3f} -> {c:7.3f}" for k, (a, b, c) in enumerate(r[:-1])) + f"  sgd ends {r[-1]:7.3f}", flush=True)
dist.barrier(); dist.destroy_process_group()
