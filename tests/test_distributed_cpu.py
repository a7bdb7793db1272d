"""world_size-2 gloo test of the multi-GPU host logic (batch sharding + bucketed gradient all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from acan_b200.parallel import GradAllReducer, init_distributed, shard_batch
    r, w, _ = init_distributed("gloo")
    assert (r, w) == (rank, world)
    torch.manual_seed(7)                                     # same weights and same global batch on every rank
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
    x, y = torch.randn(8, 6), torch.randn(8, 1)
    reducer = GradAllReducer(net.parameters(), bucket_mb=0.0005)   # tiny buckets -> several in flight
    assert len(reducer.buckets) >= 2
    xs, ys = shard_batch(x, rank, world), shard_batch(y, rank, world)
    loss = torch.nn.functional.mse_loss(net(xs), ys)
    loss.backward()
    reducer.finish()
    got = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    # expectation: mean over ranks of the per-shard gradients == gradient of the mean of per-shard losses
    ref = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.Linear(16, 16), torch.nn.ReLU(), torch.nn.Linear(16, 1))
    ref.load_state_dict(net.state_dict())
    tot = sum(torch.nn.functional.mse_loss(ref(shard_batch(x, q, world)), shard_batch(y, q, world)) for q in range(world)) / world
    tot.backward()
    want = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
    out[rank] = float((got - want).abs().max())
    # second step reuses the reducer
    net.zero_grad()
    torch.nn.functional.mse_loss(net(xs), ys).backward()
    reducer.finish()
    got2 = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    out[rank + world] = float((got2 - want).abs().max())
    # third step: hooks quiet (the mode used when backward runs inside a CUDA graph), gradients averaged afterwards by reduce_now()
    net.zero_grad()
    reducer.enabled = False
    torch.nn.functional.mse_loss(net(xs), ys).backward()
    local = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    reducer.reduce_now()
    got3 = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    out[rank + 2 * world] = float((got3 - want).abs().max())
    out[rank + 3 * world] = 0.0 if float((local - want).abs().max()) > 1e-4 else 1.0      # the local gradient alone must differ from the mean
    dist.barrier()
    dist.destroy_process_group()


def test_grad_allreduce_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == 4 * world
    assert all(v < 1e-6 for v in out.values()), dict(out)


def _stage_worker(rank, world, port, out):
    """the stage schedule of the overlapped training step on a flat gradient buffer: one asynchronous all-reduce per run of
    training.piece_ranges, issued in stage order, waited for in that order, each followed by the (here: CPU) update of its run"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from acan_b200.parallel import init_distributed
    from acan_b200.training import piece_ranges
    init_distributed("gloo")
    n, ch = 2048 * 9 + 77, 2048
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(n, generator=g)
    want = sum(torch.randn(n, generator=torch.Generator().manual_seed(100 + q)) for q in range(world)) / world
    param = torch.zeros(n)
    for stage_start, pieces in (({2: 0, 1: 3000, 0: 11000}, 4), (None, 4), (None, 1)):
        buf, param = flat.clone(), torch.zeros(n)
        works = []
        for runs in piece_ranges(n, ch, stage_start, pieces):          # (a graph replay would sit between the stages)
            works += [(dist.all_reduce(buf[lo:hi], op=dist.ReduceOp.SUM, async_op=True), (lo, hi)) for (lo, hi), _ in runs]
        for w, (lo, hi) in works:
            w.wait()
            param[lo:hi] -= 0.1 * buf[lo:hi] / world                   # the per-run update
        out[(rank, str(stage_start), pieces)] = float((param + 0.1 * want).abs().max())
    dist.barrier()
    dist.destroy_process_group()


def test_stage_allreduce_schedule_world2_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_stage_worker, args=(world, port, out), nprocs=world, join=True)
    assert len(out) == 3 * world
    assert all(v < 1e-6 for v in out.values()), dict(out)
