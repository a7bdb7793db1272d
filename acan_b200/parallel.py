"""Batch sharding + gradient all-reduce: the only multi-GPU machinery the path needs.

One process per GPU, weights replicated, every image an independent unit: inference needs no collective;
training averages gradients with NCCL all-reduce over NVLink, bucketed in reverse parameter order and
launched from post-accumulate-grad hooks so the transfers overlap the rest of the backward pass.
The reference has no distributed code on its live path; nothing here replaces reference code.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None, high_priority: bool = True, max_ctas: Optional[int] = None, min_ctas: Optional[int] = None) -> tuple:
    """(rank, world_size, local_rank) from the torchrun environment; no-op for a single process.  ``high_priority``: the NCCL
    communicator's stream gets the high CUDA priority, so the CTAs of a gradient all-reduce are placed as soon as SMs free up under the
    backward pass that is still replaying (at equal priority they queue behind every pending CTA of the compute kernels).
    ``max_ctas``: upper bound on the CTAs (= SMs) a collective of this communicator occupies (ncclConfig_t.maxCTAs); the overlapped
    training step (``GraphedTrainStep(overlap=True, nccl_sms=...)``) sizes its cooperative BatchNorm grids to the SMs left over.
    ``min_ctas``: lower bound (ncclConfig_t.minCTAs) -- an exposed all-reduce has the GPU to itself and may use more channels."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
            if high_priority or max_ctas or min_ctas:
                opts = dist.ProcessGroupNCCL.Options()
                opts.is_high_priority_stream = bool(high_priority)
                if max_ctas:
                    opts.config.max_ctas = int(max_ctas)
                if min_ctas:
                    opts.config.min_ctas = int(min_ctas)
                    if not max_ctas:
                        opts.config.max_ctas = max(int(min_ctas), 32)      # NCCL rejects a lower bound without an upper one
                kw["pg_options"] = opts
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def capped_group(max_ctas: int, high_priority: bool = True):
    """A second NCCL communicator over all ranks whose collectives occupy at most ``max_ctas`` SMs: the one an overlapped step issues
    the collective that runs BESIDE compute kernels on (the default communicator keeps its full width for the exposed collectives)."""
    if not (dist.is_initialized() and dist.get_backend() == "nccl"):
        return None
    opts = dist.ProcessGroupNCCL.Options()
    opts.is_high_priority_stream = bool(high_priority)
    opts.config.max_ctas = int(max_ctas)
    return dist.new_group(ranks=list(range(dist.get_world_size())), backend="nccl", pg_options=opts)


def shard_bounds(n_items: int, rank: int, world: int) -> tuple:
    """contiguous shard [lo, hi) of n_items for `rank`; the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


class GradAllReducer:
    """Mean of the gradients over ranks, in buckets of ~bucket_mb MB, overlapped with backward.

    usage:  reducer = GradAllReducer(net.parameters());  loss.backward();  reducer.finish();  optimizer.step()
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], bucket_mb: float = 32.0, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, cur_bytes, limit = [], 0, bucket_mb * (1 << 20)
        for p in reversed(self.params):                      # gradients become ready roughly in reverse order
            cur.append(p)
            cur_bytes += p.numel() * p.element_size()
            if cur_bytes >= limit:
                self.buckets.append(cur)
                cur, cur_bytes = [], 0
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._pending = [len(b) for b in self.buckets]
        self._work = []
        self._hooks = []
        self.enabled = True          # False: the hooks stay quiet (backward being captured into a CUDA graph); use reduce_now() after it
        if self.world > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _launch(self, bi: int) -> None:
        ps = [p for p in self.buckets[bi] if p.grad is not None]
        if not ps:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in ps])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        self._work.append((work, flat, ps))

    def _on_grad(self, p: torch.nn.Parameter) -> None:
        if not self.enabled:
            return
        bi = self._bucket_of[id(p)]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def finish(self) -> None:
        """wait for the buckets in flight, launch any bucket whose hooks never all fired (unused parameters),
        write the averaged gradients back."""
        if self.world > 1:
            for bi, left in enumerate(self._pending):
                if left > 0:
                    self._launch(bi)
            for work, flat, ps in self._work:
                work.wait()
                flat.div_(self.world)
                off = 0
                for p in ps:
                    n = p.numel()
                    p.grad.copy_(flat[off:off + n].view_as(p.grad))
                    off += n
        self._work = []
        self._pending = [len(b) for b in self.buckets]

    def reduce_now(self) -> None:
        """average the gradients that are in ``.grad`` right now, bucket by bucket, without the hooks: for a backward pass that
        ran inside a CUDA graph (a replay is asynchronous, so these launches are queued behind it while it still runs; the
        all-reduce itself then follows the backward instead of overlapping it: ~1 ms of NVLink time on a ~20 ms step)."""
        if self.world == 1:
            return
        inflight = []
        for bucket in self.buckets:
            ps = [p for p in bucket if p.grad is not None]
            if not ps:
                continue
            flat = torch.cat([p.grad.reshape(-1) for p in ps])
            inflight.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True), flat, ps))
        for work, flat, ps in inflight:
            work.wait()
            flat.div_(self.world)
            # (.grad of a channels_last weight is not contiguous in logical order: copy through shaped views, never reshape(-1))
            torch._foreach_copy_([p.grad for p in ps], [c.view(p.grad.shape) for c, p in zip(flat.split([p.numel() for p in ps]), ps)])

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []
