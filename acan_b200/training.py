"""Training-step driver for the drop-in network: the step -- forward, the two fused losses, backward, the gradient all-reduce and the
optimizer -- captured ONCE into CUDA graphs and replayed.

Why: the step is ~2500 kernel launches (cuDNN convolutions, the fused BN / attention / loss kernels of this package, the optimizer);
launched eagerly from Python the host cannot stay ahead of a B200 and the GPU idles for a quarter of the step (profiles/r01_notes.md).
Nothing on the path synchronises with the host (the attention backward takes its upstream loss gradient as a device scalar), so the
step is capturable as is.  The reference's trainer (depthest_trainer.py:118-131) calls ``net(images)``, ``criterion(...)``,
``loss.backward()``, ``optimizer.step()`` every iteration; ``GraphedTrainStep(...)(images, labels)`` is that sequence with static buffers.

Graphs of one step:
  world = 1   [forward + losses + backward]                                   [optimizer]
  world > 1   [forward + losses + backward of the LATE layers]  -> all-reduce(late gradients) on the NCCL stream
              [backward of the EARLY layers]                     -> all-reduce(early gradients)           -> [optimizer]
The late segment (layer4, decoder, classifier: ~70 % of the gradient bytes) is averaged while the second backward graph replays; only
the early segment's collective (~30 %) sits between the backward and the optimizer.  All gradients live in ONE flat fp32 buffer whose
slices ARE the ``.grad`` tensors the optimizer reads (no gather / scatter copies).  The optimizer has its own graph so that a learning
rate changed by the reference's schedulers (creator/scheduler.py: step / poly / plateau / cosine rewrite ``param_groups``) is honoured:
``__call__`` compares the hyper-parameters with the captured ones and re-captures the optimizer graph when they differ; a schedule that
changes them every step switches the optimizer to eager launches (a handful of multi-tensor kernels, hidden behind the replay).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


def _hyper_signature(optimizer) -> tuple:
    sig = []
    for g in optimizer.param_groups:
        sig.append(tuple((k, float(v) if isinstance(v, (int, float)) else repr(v)) for k, v in sorted(g.items()) if k != "params" and not torch.is_tensor(v)))
    return tuple(sig)


class GraphedTrainStep:
    """step = GraphedTrainStep(net, criterion, optimizer, images_example, labels_example);  loss = step(images, labels)

    ``criterion(preds, labels, epoch) -> (loss1, loss2, loss3, total)`` as ``ResNet.LossFunc``.  ``reducer`` (optional): a
    ``parallel.GradAllReducer`` (its process group and world size are used; with world size > 1 the step runs the split schedule above).
    ``amp_dtype=None`` runs the step in fp32.  ``graph=False`` keeps the same interface on the eager path (A/B measurements).
    ``shadow_weights`` (default: on with a 16-bit ``amp_dtype``): the convolution weights are fed to the forward pass from 16-bit
    shadow copies that are refreshed from the fp32 parameters by ONE multi-tensor copy after the optimizer step, and their 16-bit
    gradients are moved into the fp32 ``.grad`` tensors by one more -- instead of autocast's per-layer weight cast in every
    forward and per-layer gradient cast in every backward (~250 small kernels, 1.2 ms of a 17 ms step).  Same arithmetic: under
    autocast the weight gradients are produced in 16 bit as well.  Parameters, ``state_dict`` and optimizer state stay fp32.
    ``overlap`` (world > 1): split the backward at the output of ``net.layer3`` as described above (False: one backward graph, one
    all-reduce of the whole buffer).  ``after_backward``: called (and captured) in front of the optimizer step, e.g. gradient clipping.
    Shapes are fixed at construction (the network is shape-specialised anyway)."""

    def __init__(self, net, criterion, optimizer, images: torch.Tensor, labels: torch.Tensor, amp_dtype: Optional[torch.dtype] = torch.bfloat16,
                 reducer=None, epoch: int = 1, warmup: int = 3, graph: bool = True, after_backward: Optional[Callable[[], None]] = None,
                 shadow_weights: Optional[bool] = None, overlap: bool = True):
        if not images.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors: acan_b200 has no CPU path")
        self.net, self.criterion, self.optimizer, self.reducer = net, criterion, optimizer, reducer
        self.amp_dtype, self.epoch, self.after_backward = amp_dtype, epoch, after_backward
        self.static_x = images.clone()
        self.static_y = labels.clone()
        self.world = getattr(reducer, "world", 1) if reducer is not None else 1
        self.group = getattr(reducer, "group", None)
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
        self.graphs, self.opt_graph, self.losses = [], None, None
        self.library_launches_per_step = None
        self.opt_eager = False
        self._opt_sig, self._sig_changes = None, 0
        self._split = graph and self.world > 1 and overlap and hasattr(net, "layer3")
        self._boundary = None
        self._hook = None
        if not graph:
            return
        if reducer is not None:
            reducer.enabled = False                            # the collectives are issued by this driver, not by grad hooks
        dev = images.device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):                    # cuDNN autotuning, workspace allocations, optimizer state, NCCL communicator
                optimizer.zero_grad(set_to_none=True)
                self._forward_backward(phase=0)
                self._reduce_all_eager()
                self._optimizer_part()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self._bind_flat_gradients()
        optimizer.zero_grad(set_to_none=True)                  # the captured backward must CREATE its gradients (no in-place accumulation)
        for s_ in self._shadows:
            s_.grad = None
        from . import _lib
        n0 = _lib.load().acan_launch_count()
        if self._split:
            self._hook = net.layer3.register_forward_hook(self._grab_boundary)
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1):
            self.losses = self._forward_backward(phase=1 if self._split else 0)
        self.graphs = [g1]
        if self._split:
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, pool=g1.pool()):
                self._forward_backward(phase=2)
            self.graphs.append(g2)
            self._hook.remove()
        self.library_launches_per_step = int(_lib.load().acan_launch_count() - n0)        # kernels of libacan_b200 in one replay
        self._capture_optimizer()

    # ------------------------------------------------------------------ gradients: one flat buffer, .grad = its slices
    def _bind_flat_gradients(self) -> None:
        """ONE fp32 buffer holds every gradient, in ``net.parameters()`` order (conv1 ... layer3 | layer4, decoder, classifier), and the
        ``.grad`` of every parameter is a view of its slice with the parameter's own memory layout: the multi-tensor copy that moves the
        16-bit shadow gradients to fp32 writes straight into it, the optimizer reads it in place, and an all-reduce of a slice of the
        buffer averages a whole segment of the network with no gather or scatter."""
        params, seen = [], set()
        for p in self.net.parameters():
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        self._flat_params = params
        dev = self.static_x.device
        self._flat = torch.zeros(sum(p.numel() for p in params), device=dev, dtype=torch.float32)
        late = set()
        for name in ("layer4", "decoder", "interout", "classifier"):
            m = getattr(self.net, name, None)
            if m is not None:
                late.update(id(p) for p in m.parameters())
        self._views, off, self._n_early = {}, 0, None
        for p in params:
            if self._n_early is None and id(p) in late:
                self._n_early = off
            seg = self._flat[off:off + p.numel()]
            off += p.numel()
            if p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last):
                v = seg.view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2)
            else:
                v = seg.view(p.shape)
            self._views[id(p)] = v
        if self._n_early is None:
            self._n_early = off
        self._late_ids = late
        shadowed = {id(p) for p in self._shadow_params}
        self._plain_params = [p for p in params if id(p) not in shadowed]      # gradients produced by autograd directly (fp32, small)

    def _grab_boundary(self, module, inputs, output):
        self._boundary = output

    # ------------------------------------------------------------------ pieces of one step
    def _forward_backward(self, phase: int = 0):
        """phase 0: forward + losses + the whole backward.  phase 1: forward + losses + backward down to the output of layer3 (late
        parameters and the boundary gradient).  phase 2: backward from the boundary gradient through the early layers."""
        out = None
        late_only = lambda ps: [p for p in ps if id(p) in self._late_ids]            # noqa: E731
        early_only = lambda ps: [p for p in ps if id(p) not in self._late_ids]       # noqa: E731
        if phase in (0, 1):
            with torch.autocast("cuda", dtype=self.amp_dtype if self.amp_dtype is not None else torch.bfloat16, enabled=self.amp_dtype is not None):
                if self._shadows:
                    # tie_weights=False: f_query IS f_key (one module object, so one name reaches both), and with tie_weights=True the
                    # ORIGINAL parameter is handed a gradient as well -- under capture that left its .grad out of the optimizer graph
                    preds = torch.func.functional_call(self.net, dict(zip(self._shadow_names, self._shadows)), (self.static_x,), tie_weights=False, strict=False)
                else:
                    preds = self.net(self.static_x)
                out = self.criterion(preds, self.static_y, self.epoch)
            if phase == 0:
                out[3].backward()
            else:
                b = self._boundary
                leaves = late_only(self._plain_params) + [s for p, s in zip(self._shadow_params, self._shadows) if id(p) in self._late_ids]
                self._boundary_grad = None
                grads = torch.autograd.grad(out[3], [b] + leaves, allow_unused=True)
                self._boundary_grad = grads[0]
                for t, g in zip(leaves, grads[1:]):
                    t.grad = g
        else:
            torch.autograd.backward(self._boundary, grad_tensors=self._boundary_grad)
            self._boundary, self._boundary_grad = None, None
        if hasattr(self, "_views"):
            pick = (lambda ps: ps) if phase == 0 else (late_only if phase == 1 else early_only)
            self._collect(pick)
        return out

    def _collect(self, pick) -> None:
        """gradients of this phase -> their slices of the flat buffer (two multi-tensor copies: the 16-bit shadow gradients, converted;
        the few fp32 ones autograd produced directly)."""
        chosen = {id(p) for p in pick(self._flat_params)}
        dst, src = [], []
        for p, s_ in zip(self._shadow_params, self._shadows):
            if id(p) in chosen and s_.grad is not None:
                dst.append(self._views[id(p)])
                src.append(s_.grad)
        if dst:
            torch._foreach_copy_(dst, src)
        for p, s_ in zip(self._shadow_params, self._shadows):
            if id(p) in chosen:
                s_.grad = None
        dst, src, zero = [], [], []
        for p in self._plain_params:
            if id(p) in chosen:
                if p.grad is not None:
                    dst.append(self._views[id(p)])
                    src.append(p.grad)
                else:
                    zero.append(self._views[id(p)])
        if dst:
            torch._foreach_copy_(dst, src)
        if zero:
            torch._foreach_zero_(zero)
        for p in self._flat_params:
            if id(p) in chosen:
                p.grad = self._views[id(p)]

    def _reduce_all_eager(self) -> None:
        """warm-up passes (before the flat buffer exists): average whatever is in .grad, creating the communicator on the way"""
        if self.world > 1:
            import torch.distributed as dist
            if self._shadows:
                for p, s_ in zip(self._shadow_params, self._shadows):
                    if s_.grad is not None:
                        p.grad = s_.grad.float()
                        s_.grad = None
            for p in self.net.parameters():
                if p.grad is not None:
                    dist.all_reduce(p.grad, op=dist.ReduceOp.AVG, group=self.group)
        elif self._shadows:
            for p, s_ in zip(self._shadow_params, self._shadows):
                if s_.grad is not None:
                    p.grad = s_.grad.float()
                    s_.grad = None

    def _optimizer_part(self) -> None:
        if self.after_backward is not None:
            self.after_backward()
        self.optimizer.step()
        if self._shadows:
            with torch.no_grad():
                torch._foreach_copy_(self._shadows, self._shadow_params)          # refresh the 16-bit copies: one multi-tensor kernel

    def _capture_optimizer(self) -> None:
        self._opt_sig = _hyper_signature(self.optimizer)
        if self.opt_eager:
            self.opt_graph = None
            return
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=self.graphs[0].pool()):
            self._optimizer_part()
        self.opt_graph = g

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
        if not self.graphs:
            self.optimizer.zero_grad(set_to_none=True)
            self.losses = self._forward_backward(phase=0)
            if self.reducer is not None:
                self.reducer.finish()
            if self._shadows:
                for p, s_ in zip(self._shadow_params, self._shadows):
                    if s_.grad is not None:
                        p.grad = s_.grad.float() if p.grad is None else p.grad
                        s_.grad = None
            self._optimizer_part()
            return self.losses[3]
        sig = _hyper_signature(self.optimizer)
        if sig != self._opt_sig:                                # a scheduler rewrote lr / momentum / ...: honour it
            self._sig_changes += 1
            if self._sig_changes >= 3 and not self.opt_eager:
                self.opt_eager = True                           # changes every few steps (poly / cosine): stop re-capturing
            torch.cuda.current_stream(self.static_x.device).synchronize()
            self._capture_optimizer()
        if self.world > 1:
            import torch.distributed as dist
            self.graphs[0].replay()
            if self._split:
                wa = dist.all_reduce(self._flat[self._n_early:], op=dist.ReduceOp.AVG, group=self.group, async_op=True)   # overlaps the second graph
                self.graphs[1].replay()
                wb = dist.all_reduce(self._flat[:self._n_early], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
                wa.wait()
                wb.wait()
            else:
                dist.all_reduce(self._flat, op=dist.ReduceOp.AVG, group=self.group)
        else:
            self.graphs[0].replay()
        if self.opt_graph is not None:
            self.opt_graph.replay()
        else:
            self._optimizer_part()
        return self.losses[3]

    # kept for callers of the first version
    @property
    def graph(self):
        return self.graphs[0] if self.graphs else None
