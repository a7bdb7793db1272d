"""Training-step driver for the drop-in network: the whole step -- forward, the two fused losses, backward, the gradient
all-reduce and the optimizer -- captured ONCE into a CUDA graph and replayed.

Why: the step is ~2500 kernel launches (cuDNN convolutions, the fused BN / attention / loss kernels of this package, the
optimizer); launched eagerly from Python the host cannot stay ahead of a B200 and the GPU idles for a quarter of the step
(profiles/r01_notes.md).  Nothing on the path synchronises with the host (the attention backward takes its upstream loss
gradient as a device scalar), so the step is capturable as is.  The reference's trainer (depthest_trainer.py:118-131) calls
``net(images)``, ``criterion(...)``, ``loss.backward()``, ``optimizer.step()`` every iteration; ``GraphedTrainStep(...)(images,
labels)`` is that sequence with static input buffers.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


class GraphedTrainStep:
    """step = GraphedTrainStep(net, criterion, optimizer, images_example, labels_example);  loss = step(images, labels)

    ``criterion(preds, labels, epoch) -> (loss1, loss2, loss3, total)`` as ``ResNet.LossFunc``.  ``reducer`` (optional): a
    ``parallel.GradAllReducer``.  With a reducer of world size > 1 the step is two graphs -- forward + losses + backward, then
    the optimizer -- with ONE NCCL all-reduce launched eagerly between the replays: the backward graph ends by gathering every
    gradient into a single fp32 buffer, the buffer is averaged in place, and the optimizer graph starts by scattering it back
    (the replays are asynchronous, so the host queues the collective while the backward graph still runs).
    ``amp_dtype=None`` runs the step in fp32.  ``graph=False`` keeps the same interface on the eager path (A/B measurements).
    ``shadow_weights`` (default: on with a 16-bit ``amp_dtype``): the convolution weights are fed to the forward pass from 16-bit
    shadow copies that are refreshed from the fp32 parameters by ONE multi-tensor copy after the optimizer step, and their 16-bit
    gradients are moved into the fp32 ``.grad`` tensors by one more -- instead of autocast's per-layer weight cast in every
    forward and per-layer gradient cast in every backward (~250 small kernels, 1.2 ms of a 17 ms step).  Same arithmetic: under
    autocast the weight gradients are produced in 16 bit as well.  Parameters, ``state_dict`` and optimizer state stay fp32.
    Shapes are fixed at construction (the network is shape-specialised anyway)."""

    def __init__(self, net, criterion, optimizer, images: torch.Tensor, labels: torch.Tensor, amp_dtype: Optional[torch.dtype] = torch.bfloat16,
                 reducer=None, epoch: int = 1, warmup: int = 3, graph: bool = True, after_backward: Optional[Callable[[], None]] = None,
                 shadow_weights: Optional[bool] = None):
        if not images.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors: acan_b200 has no CPU path")
        self.net, self.criterion, self.optimizer, self.reducer = net, criterion, optimizer, reducer
        self.amp_dtype, self.epoch, self.after_backward = amp_dtype, epoch, after_backward
        self.static_x = images.clone()
        self.static_y = labels.clone()
        if shadow_weights is None:
            shadow_weights = amp_dtype in (torch.float16, torch.bfloat16)
        self._shadow_names, self._shadow_params, self._shadows = [], [], []
        if shadow_weights and amp_dtype is not None:
            seen = set()
            for mod_name, mod in net.named_modules():
                w = getattr(mod, "weight", None)
                if isinstance(mod, torch.nn.Conv2d) and w is not None and w.requires_grad and w.dtype == torch.float32 and id(w) not in seen:
                    seen.add(id(w))
                    self._shadow_names.append(f"{mod_name}.weight" if mod_name else "weight")
                    self._shadow_params.append(w)
                    self._shadows.append(w.detach().to(amp_dtype, memory_format=torch.preserve_format).requires_grad_(True))
        self.graph = None
        self.opt_graph = None
        self.losses = None
        self.library_launches_per_step = None
        self._split = reducer is not None and getattr(reducer, "world", 1) > 1       # NCCL stays outside the graphs
        self._flat_grad = None
        if not graph:
            return
        if self._split:
            reducer.enabled = False
        side = torch.cuda.Stream(device=images.device)
        side.wait_stream(torch.cuda.current_stream(images.device))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):                  # cuDNN autotuning, workspace allocations, optimizer state, NCCL communicator
                optimizer.zero_grad(set_to_none=True)
                self._body()
        torch.cuda.current_stream(images.device).wait_stream(side)
        torch.cuda.synchronize(images.device)
        optimizer.zero_grad(set_to_none=True)                # captured backward re-creates the .grad tensors in the graph's pool
        if self._split:
            self._bind_flat_gradients()
        self.graph = torch.cuda.CUDAGraph()
        from . import _lib
        n0 = _lib.load().acan_launch_count()
        with torch.cuda.graph(self.graph):
            self.losses = self._body(capture_part=1 if self._split else 0)
        self.library_launches_per_step = int(_lib.load().acan_launch_count() - n0)    # kernels of libacan_b200 in one replay
        if self._split:
            self.opt_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.opt_graph, pool=self.graph.pool()):
                self._body(capture_part=2)

    def _bind_flat_gradients(self) -> None:
        """multi-GPU replay mode: ONE fp32 buffer holds every gradient for the collective.  The backward graph ends with a
        multi-tensor copy of the ``.grad`` tensors into views of it (each view has its gradient's own memory layout, so the copy is
        the fused fast path), the gradient average is a single in-place NCCL all-reduce, and the optimizer graph starts by copying
        the views back."""
        self._flat_params, seen = [], set()
        for p in self.net.parameters():
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                self._flat_params.append(p)
        self._flat_grad = torch.zeros(sum(p.numel() for p in self._flat_params), device=self.static_x.device, dtype=torch.float32)
        self._flat_views, self._flat_grads = None, None

    def _flat_pairs(self):
        if self._flat_views is None:                          # built once, inside the capture, from the gradients the backward produced
            views, grads, off = [], [], 0
            for p in self._flat_params:
                seg = self._flat_grad[off:off + p.numel()]
                off += p.numel()
                g = p.grad
                if g is None:
                    continue
                if g.dim() == 4 and not g.is_contiguous() and g.is_contiguous(memory_format=torch.channels_last):
                    v = seg.view(g.shape[0], g.shape[2], g.shape[3], g.shape[1]).permute(0, 3, 1, 2)
                else:
                    v = seg.view(g.shape)
                views.append(v)
                grads.append(g)
            self._flat_views, self._flat_grads = views, grads
        return self._flat_views, self._flat_grads

    def _body(self, capture_part: int = 0):
        """capture_part 0: the whole step; 1: forward + losses + backward only; 2: what follows the gradient all-reduce."""
        out = None
        if capture_part in (0, 1):
            with torch.autocast("cuda", dtype=self.amp_dtype if self.amp_dtype is not None else torch.bfloat16, enabled=self.amp_dtype is not None):
                if self._shadows:
                    # tie_weights=False: f_query IS f_key (one module object, so one name reaches both), and with tie_weights=True the
                    # ORIGINAL parameter is handed a gradient as well -- under capture that left its .grad out of the optimizer graph
                    preds = torch.func.functional_call(self.net, dict(zip(self._shadow_names, self._shadows)), (self.static_x,), tie_weights=False, strict=False)
                else:
                    preds = self.net(self.static_x)
                out = self.criterion(preds, self.static_y, self.epoch)
            out[3].backward()
            if self._shadows:                                # 16-bit weight gradients -> the fp32 .grad tensors the optimizer / all-reduce see
                for p, s_ in zip(self._shadow_params, self._shadows):
                    if p.grad is None:
                        p.grad = torch.empty_like(p)
                torch._foreach_copy_([p.grad for p in self._shadow_params], [s_.grad for s_ in self._shadows])
                for s_ in self._shadows:
                    s_.grad = None
            if capture_part == 1 and self._flat_grad is not None:
                views, grads = self._flat_pairs()
                torch._foreach_copy_(views, grads)
        if capture_part == 0 and self.reducer is not None:
            if self._split and not self.reducer.enabled:
                self.reducer.reduce_now()                    # warm-up passes of the split mode: hooks are off
            else:
                self.reducer.finish()
        if capture_part in (0, 2):
            if capture_part == 2 and self._flat_grad is not None:
                views, grads = self._flat_pairs()
                torch._foreach_copy_(grads, views)
            if self.after_backward is not None:
                self.after_backward()
            self.optimizer.step()
            if self._shadows:
                with torch.no_grad():
                    torch._foreach_copy_(self._shadows, self._shadow_params)      # refresh the 16-bit copies: one multi-tensor kernel
        return out

    def sync_weights(self) -> None:
        """re-read the fp32 parameters into the 16-bit shadow copies: call after changing the weights from outside the step
        (``load_state_dict``, manual edits); the step itself keeps them in sync."""
        if self._shadows:
            with torch.no_grad():
                torch._foreach_copy_(self._shadows, self._shadow_params)

    def __call__(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """one training step; returns the total loss (a device scalar; static across steps when graphed)."""
        self.static_x.copy_(images, non_blocking=True)
        self.static_y.copy_(labels, non_blocking=True)
        if self.graph is None:
            self.optimizer.zero_grad(set_to_none=True)
            self.losses = self._body()
        else:
            self.graph.replay()
            if self.opt_graph is not None:
                import torch.distributed as dist
                dist.all_reduce(self._flat_grad, op=dist.ReduceOp.AVG, group=self.reducer.group)     # one in-place collective, queued behind the replay
                self.opt_graph.replay()
        return self.losses[3]
