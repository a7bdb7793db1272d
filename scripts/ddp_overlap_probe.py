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
# the two graphs back to back with no collective in between: what the split itself costs
for _ in range(3):
    a, b_ = ev(), ev()
    a.record(main); step.graphs[0].replay(); m_ = ev(); m_.record(main); step.graphs[1].replay(); b_.record(main)
    torch.cuda.synchronize()
    if rank == 0:
        print(f"no collective: first graph {a.elapsed_time(m_):.3f} ms, second graph {m_.elapsed_time(b_):.3f} ms, both {a.elapsed_time(b_):.3f} ms", flush=True)
dist.barrier()
rows = []
for it in range(8):
    e = {k: ev() for k in ("t0", "g1", "ar1s", "ar1e", "g2", "ar2e", "opt")}
    e["t0"].record(main)
    step.graphs[0].replay()
    e["g1"].record(main)
    comm.wait_stream(main)
    with torch.cuda.stream(comm):
        e["ar1s"].record(comm)
        dist.all_reduce(step._flat[step._n_early:], op=dist.ReduceOp.AVG, group=late_group)
        e["ar1e"].record(comm)
    step.graphs[1].replay()
    e["g2"].record(main)
    comm.wait_stream(main)
    with torch.cuda.stream(comm):
        dist.all_reduce(step._flat[:step._n_early], op=dist.ReduceOp.AVG)
        e["ar2e"].record(comm)
    main.wait_stream(comm)
    step._sgd_range(0, step._n_sgd_chunks)
    e["opt"].record(main)
    torch.cuda.synchronize()
    rows.append({k: e["t0"].elapsed_time(v) for k, v in e.items() if k != "t0"})
if rank == 0:
    for r in rows[2:]:
        print("ms from step start: " + "  ".join(f"{k} {v:7.3f}" for k, v in r.items()), flush=True)
    r = rows[-1]
    print(f"late all-reduce ran {r['ar1s']:.3f} -> {r['ar1e']:.3f} ms; second backward graph ended at {r['g2']:.3f} ms: "
          + ("OVERLAPPED" if r["ar1e"] < r["g2"] - 0.05 else "NOT overlapped (the collective finished after the graph)"))
dist.barrier(); dist.destroy_process_group()
