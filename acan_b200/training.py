"""Training-step driver for the drop-in network: the step -- forward, the two fused losses, backward, the gradient all-reduce and the
optimizer -- captured ONCE into CUDA graphs and replayed.

Why: the step is ~2500 kernel launches (cuDNN convolutions, the fused BN / attention / loss kernels of this package, the optimizer);
launched eagerly from Python the host cannot stay ahead of a B200 and the GPU idles for a quarter of the step (profiles/r01_notes.md).
Nothing on the path synchronises with the host (the attention backward takes its upstream loss gradient as a device scalar), so the
step is capturable as is.  The reference's trainer (depthest_trainer.py:118-131) calls ``net(images)``, ``criterion(...)``,
``loss.backward()``, ``optimizer.step()`` every iteration; ``GraphedTrainStep(...)(images, labels)`` is that sequence with static buffers.

One step:
  world = 1   [graph: forward + losses + backward]  ->  optimizer
  world > 1   [graph: forward + losses + backward]  ->  all-reduce(piece 0) all-reduce(piece 1) ... on the NCCL stream
                                                         optimizer(piece 0) as soon as piece 0 has landed, optimizer(piece 1) ...
  world > 1, overlap=True
              [graph 1: forward + losses + backward of layer4 / decoder / classifier] -> all-reduce(stage 0) beside
              [graph 2: backward of layer3]                                           -> all-reduce(stage 1) beside
              [graph 3: backward of layer2, layer1, the stem]                         -> all-reduce(stage 2: 6 MB, the only exposed one)
              optimizer(stage 0), optimizer(stage 1), optimizer(stage 2)
All gradients live in ONE flat fp32 buffer whose slices ARE the ``.grad`` tensors (no gather / scatter copies).  With torch.optim.SGD
(what the reference trains with, creator/optimizer.py:8-10) the parameters, momentum buffers and 16-bit shadow weights are bound to flat
buffers as well and the update is ONE streaming kernel (acan_sgd_flat: 1.1 ms of multi-tensor passes -> 0.3 ms), launched per piece, so
only the first piece's collective and the last piece's update are exposed.  Its hyper-parameters sit in a device table refreshed from
``param_groups`` whenever a scheduler (creator/scheduler.py: step / poly / plateau / cosine) changed them -- no re-capture.  Any other
optimizer keeps ``optimizer.step()`` in its own CUDA graph, re-captured when the hyper-parameters change (eager after three changes).

``overlap=True`` splits the backward at the outputs of layer3 and layer2 into three graphs and averages each stage's gradients while
the next graph replays.  Measured on 2 B200s (scripts/ddp_overlap_probe.py, profiles/r02_notes.md): the collective does run concurrently, but
cooperative BatchNorm grids that need every SM cannot start while NCCL's CTAs are resident, so the second graph stalled for as long as
the collective ran.  The step therefore captures the BN grids of the second graph over 148 - ``nccl_sms`` SMs and issues that collective on
``overlap_group``, a communicator capped to ``nccl_sms`` CTAs (``parallel.capped_group``); the exposed collective of the early layers
keeps the default communicator's full width.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch


def _capture_mode() -> str:
    """cudaStreamCaptureMode of the step's captures.  With a process group alive, NCCL's watchdog thread polls CUDA events every 100 ms;
    under the default "global" mode such a call from ANOTHER thread invalidates an ongoing capture -- at random, on some ranks only, which
    then leaves the ranks on different collective schedules.  "thread_local" restricts the check to the capturing thread (what
    torch._inductor's CUDA-graph trees use); single-process runs keep the default."""
    import torch.distributed as dist
    return "thread_local" if (dist.is_available() and dist.is_initialized()) else "global"


def piece_ranges(n: int, ch: int, stage_start: Optional[dict], pieces: int) -> list:
    """[[(element range, chunk range), ...] per backward graph]: a flat buffer of ``n`` elements cut into runs of whole optimizer chunks of
    ``ch`` elements.  One graph (``stage_start`` None): ``pieces`` runs (the update of run k hides the all-reduce of run k + 1).  Split
    backward: one run per stage, in the order the stages complete -- stage 0 (layer4 / decoder / classifier: the END of the buffer), stage 1
    (layer3), stage 2 (the rest: the START); ``stage_start[k]`` is the element offset of stage k's first parameter.  A run may start a few
    elements inside the stage that completed before it (chunk alignment): those gradients are complete by then.  Every element belongs to
    exactly one run."""
    nch = (n + ch - 1) // ch

    def cut(c0, c1, k):
        k = max(1, min(k, c1 - c0))
        edges = [c0 + (c1 - c0) * i // k for i in range(k + 1)]
        return [((a * ch, min(b * ch, n)), (a, b - a)) for a, b in zip(edges[:-1], edges[1:]) if b > a]
    if stage_start is not None:
        c0 = min((stage_start[0] + ch - 1) // ch, nch)       # first whole chunk of stage 0
        c1 = min((stage_start[1] + ch - 1) // ch, c0)        # ... of stage 1
        return [cut(c0, nch, 1), cut(c1, c0, 1), cut(0, c1, 1)]
    return [cut(0, nch, pieces)]


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
    Shapes are fixed at construction (the network is shape-specialised anyway).
    ``fused_grad_cast``: move the 16-bit weight gradients into the flat buffer with acan_cast_to_f32_multi (one launch) instead of the
    framework's per-tensor casts."""

    def __init__(self, net, criterion, optimizer, images: torch.Tensor, labels: torch.Tensor, amp_dtype: Optional[torch.dtype] = torch.bfloat16,
                 reducer=None, epoch: int = 1, warmup: int = 3, graph: bool = True, after_backward: Optional[Callable[[], None]] = None,
                 shadow_weights: Optional[bool] = None, overlap: bool = False, pieces: int = 4, nccl_sms: int = 16, overlap_group=None, fused_grad_cast: bool = False,
                 _force_split: bool = False):
        if not images.is_cuda:
            raise RuntimeError("GraphedTrainStep needs CUDA tensors: acan_b200 has no CPU path")
        self.net, self.criterion, self.optimizer, self.reducer = net, criterion, optimizer, reducer
        self.amp_dtype, self.epoch, self.after_backward = amp_dtype, epoch, after_backward
        self.static_x = images.clone()
        self.static_y = labels.clone()
        self.world = getattr(reducer, "world", 1) if reducer is not None else 1
        self.group = getattr(reducer, "group", None)
        self.overlap_group = overlap_group if overlap_group is not None else self.group     # communicator of the collective that runs beside graph 2
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
        self._flat_sgd = False
        self._never_grad = set()
        self._pieces = max(int(pieces), 1)
        self._fused_grad_cast = bool(fused_grad_cast)
        self._split = graph and (self.world > 1 or _force_split) and overlap and hasattr(net, "layer3") and hasattr(net, "layer2")
        # stages of the split backward, in the order their gradients complete: 0 = layer4 / decoder / classifier, 1 = layer3, 2 = the rest;
        # the output of _bound_modules[k] is the boundary between stage k and stage k + 1
        self._bound_modules = [net.layer3, net.layer2] if self._split else []
        self._boundaries, self._boundary_grads, self._hooks = {}, {}, []
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
        self._bind_flat_sgd()
        optimizer.zero_grad(set_to_none=True)                  # the captured backward must CREATE its gradients (no in-place accumulation)
        for s_ in self._shadows:
            s_.grad = None
        from . import _lib
        n0 = _lib.load().acan_launch_count()
        if self._split and not self._stages_contiguous():
            self._split, self._bound_modules = False, []       # parameter order does not follow the stages: one backward graph
        for k, m in enumerate(self._bound_modules):
            self._hooks.append(m.register_forward_hook(lambda mod, inp, out, k=k: self._boundaries.__setitem__(k, out)))
        g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g1, capture_error_mode=_capture_mode()):
            self.losses = self._forward_backward(phase=1 if self._split else 0)
        self.graphs = [g1]
        if self._split:
            # the cooperative BN grids of the LATER graphs leave `nccl_sms` SMs to the collective that runs beside them (a cooperative grid
            # starts only when all of its blocks fit at once; that collective's communicator is capped: parallel.capped_group)
            sm_limit = _lib.load().acan_bn_set_sm_limit(148 - int(nccl_sms))
            try:
                for phase in range(2, len(self._bound_modules) + 2):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, pool=g1.pool(), capture_error_mode=_capture_mode()):
                        self._forward_backward(phase=phase)
                    self.graphs.append(g)
            finally:
                _lib.load().acan_bn_set_sm_limit(sm_limit)
                for h in self._hooks:
                    h.remove()
                self._hooks = []
        self.library_launches_per_step = int(_lib.load().acan_launch_count() - n0)        # kernels of libacan_b200 in one replay
        if self._flat_sgd and self._never_grad:
            # parameters autograd never reached: torch.optim.SGD skips them (no weight decay, no momentum); the frozen table row does the same
            self._seg_group_host = [self._frozen_row if id(p) in self._never_grad else g for p, g in zip(self._flat_params, self._seg_group_host)]
            self._build_sgd_tables()
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
        stage_of = {}
        for stage, names in enumerate((("layer4", "decoder", "interout", "classifier"), ("layer3",))):
            for name in names:
                m = getattr(self.net, name, None)
                if m is not None:
                    stage_of.update((id(p), stage) for p in m.parameters())
        self._stage_of = {id(p): stage_of.get(id(p), 2) for p in params}
        self._views, off = {}, 0
        self._stage_start = {}                                  # stage -> element offset of its first parameter
        for p in params:
            self._stage_start.setdefault(self._stage_of[id(p)], off)
            seg = self._flat[off:off + p.numel()]
            off += p.numel()
            if p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last):
                v = seg.view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2)
            else:
                v = seg.view(p.shape)
            self._views[id(p)] = v
        self._n_early = self._stage_start.get(0, off)            # kept for callers of the two-segment version: start of the late segment
        shadowed = {id(p) for p in self._shadow_params}
        self._plain_params = [p for p in params if id(p) not in shadowed]      # gradients produced by autograd directly (fp32, small)

    def _stages_contiguous(self) -> bool:
        """the split schedule all-reduces one slice of the flat buffer per stage: parameters must come as [stage 2 | stage 1 | stage 0]"""
        order = [self._stage_of[id(p)] for p in self._flat_params]
        return order == sorted(order, reverse=True) and set(order) == {0, 1, 2}

    def _layout_view(self, flat: torch.Tensor, off: int, p: torch.Tensor) -> torch.Tensor:
        seg = flat[off:off + p.numel()]
        if p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last):
            return seg.view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2)
        return seg.view(p.shape)

    def _bind_flat_sgd(self) -> None:
        """torch.optim.SGD without nesterov / dampening: parameters, momentum buffers and 16-bit shadows become slices of flat buffers (same
        order and memory layout as the gradient buffer) and the update becomes acan_sgd_flat.  The optimizer object keeps working as the
        owner of the state (``state_dict`` / ``load_state_dict`` see the same tensors) and of the hyper-parameters."""
        opt = self.optimizer
        if type(opt) is not torch.optim.SGD:
            return
        group_of = {}
        for gi, g in enumerate(opt.param_groups):
            if g.get("nesterov") or g.get("dampening", 0) != 0 or g.get("maximize") or g.get("differentiable"):
                return
            for p in g["params"]:
                group_of[id(p)] = gi
        if any(id(p) not in group_of or p.dtype != torch.float32 for p in self._flat_params):
            return
        dev = self._flat.device
        n = self._flat.numel()
        self._flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self._flat_m = torch.zeros(n, device=dev, dtype=torch.float32)
        self._flat_s = torch.zeros(n, device=dev, dtype=self.amp_dtype) if (self._shadows and self.amp_dtype == torch.bfloat16) else None
        shadow_of = {id(p): i for i, p in enumerate(self._shadow_params)}
        seg_end, seg_group, off = [], [], 0
        frozen = len(opt.param_groups)                        # extra table row (lr 0, momentum 1, wd 0): parameters autograd never reaches
        with torch.no_grad():
            for p in self._flat_params:
                v = self._layout_view(self._flat_p, off, p)
                v.copy_(p.data)
                p.data = v
                st = opt.state.get(p)
                mv = self._layout_view(self._flat_m, off, p)
                if st is not None and st.get("momentum_buffer") is not None:
                    mv.copy_(st["momentum_buffer"])
                    st["momentum_buffer"] = mv
                if self._flat_s is not None and id(p) in shadow_of:
                    sv = self._layout_view(self._flat_s, off, p)
                    sv.copy_(p.data)
                    self._shadows[shadow_of[id(p)]] = sv.requires_grad_(True)
                off += p.numel()
                seg_end.append(off)
                seg_group.append(group_of[id(p)])
        self._seg_end_host, self._seg_group_host = seg_end, seg_group
        self._hyper = torch.zeros(frozen + 1, 4, device=dev, dtype=torch.float32)
        self._frozen_row = frozen
        self._build_sgd_tables()
        self._flat_sgd = True
        self._refresh_hyper()

    def _build_sgd_tables(self) -> None:
        """device tables of acan_sgd_flat: per parameter slice its end offset and parameter group, per chunk of elements its group (or -1)."""
        import numpy as np
        from . import _lib
        dev, n = self._flat.device, self._flat.numel()
        ch = int(_lib.load().acan_sgd_flat_chunk())
        self._sgd_chunk = ch
        n_chunks = (n + ch - 1) // ch
        ends = np.asarray(self._seg_end_host, dtype=np.int64)
        grp = np.asarray(self._seg_group_host, dtype=np.int32)
        starts = np.arange(n_chunks, dtype=np.int64) * ch
        first = np.searchsorted(ends, starts, side="right")
        last = np.searchsorted(ends, np.minimum(starts + ch, n) - 1, side="right")
        cg = np.full(n_chunks, -1, dtype=np.int8)
        for c in range(n_chunks):
            gs = grp[first[c]:last[c] + 1]
            if gs.size and (gs == gs[0]).all():
                cg[c] = gs[0]
        self._seg_end = torch.from_numpy(ends).to(dev)
        self._seg_grp = torch.from_numpy(grp).to(dev)
        self._chunk_group = torch.from_numpy(cg).to(dev)
        self._n_sgd_chunks = n_chunks

    def _refresh_hyper(self) -> None:
        rows = [[float(g["lr"]), float(g.get("momentum", 0.0)), float(g.get("weight_decay", 0.0)), 0.0] for g in self.optimizer.param_groups]
        rows.append([0.0, 1.0, 0.0, 0.0])
        self._hyper.copy_(torch.tensor(rows, dtype=torch.float32))
        self._opt_sig = _hyper_signature(self.optimizer)

    def _sgd_range(self, first_chunk: int, n_chunks: int) -> None:
        from . import _lib
        with torch.cuda.device(self._flat.device):
            _lib.check(_lib.load().acan_sgd_flat(self._flat_p.data_ptr(), self._flat.data_ptr(), self._flat_m.data_ptr(), _lib.ptr(self._flat_s),
                                                 self._chunk_group.data_ptr(), self._seg_end.data_ptr(), self._seg_grp.data_ptr(), int(self._seg_end.numel()),
                                                 self._hyper.data_ptr(), int(first_chunk), int(n_chunks), int(self._flat.numel()), _lib.current_stream()),
                       "acan_sgd_flat")

    # ------------------------------------------------------------------ pieces of one step
    def _forward_backward(self, phase: int = 0):
        """phase 0: forward + losses + the whole backward.  phase 1: forward + losses + backward down to the first boundary (stage-0
        parameters and the boundary gradient).  phase k > 1: backward from boundary k - 2 to boundary k - 1 (the last one: to the input)."""
        out = None
        of_stage = lambda st: (lambda ps: [p for p in ps if self._stage_of[id(p)] == st])      # noqa: E731
        def leaves_of(st):
            return of_stage(st)(self._plain_params) + [s_ for p, s_ in zip(self._shadow_params, self._shadows) if self._stage_of[id(p)] == st]
        if phase in (0, 1):
            from .functions import deferred_bn_counters
            with deferred_bn_counters(), torch.autocast("cuda", dtype=self.amp_dtype if self.amp_dtype is not None else torch.bfloat16, enabled=self.amp_dtype is not None):
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
                leaves = leaves_of(0)
                grads = torch.autograd.grad(out[3], [self._boundaries[0]] + leaves, allow_unused=True)
                self._boundary_grads[0] = grads[0]
                for t, g in zip(leaves, grads[1:]):
                    t.grad = g
        else:
            k = phase - 2                                         # boundary this phase starts from
            if k + 1 < len(self._bound_modules):
                leaves = leaves_of(k + 1)
                grads = torch.autograd.grad(self._boundaries[k], [self._boundaries[k + 1]] + leaves, grad_outputs=self._boundary_grads[k], allow_unused=True)
                self._boundary_grads[k + 1] = grads[0]
                for t, g in zip(leaves, grads[1:]):
                    t.grad = g
            else:
                torch.autograd.backward(self._boundaries[k], grad_tensors=self._boundary_grads[k])
                self._boundaries, self._boundary_grads = {}, {}
        if hasattr(self, "_views"):
            self._collect((lambda ps: ps) if phase == 0 else of_stage(phase - 1))
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
            from . import ops
            if not (self._fused_grad_cast and ops.cast_to_f32_multi(dst, src)):     # one launch; the framework's mixed-dtype foreach copy is
                torch._foreach_copy_(dst, src)                                      # one kernel per tensor
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
                    self._never_grad.add(id(p))
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
        if self._flat_sgd:
            self._sgd_range(0, self._n_sgd_chunks)               # parameters, momentum and 16-bit shadows in one pass
            if self._flat_s is None and self._shadows:
                with torch.no_grad():
                    torch._foreach_copy_(self._shadows, self._shadow_params)
            return
        self.optimizer.step()
        if self._shadows:
            with torch.no_grad():
                torch._foreach_copy_(self._shadows, self._shadow_params)          # refresh the 16-bit copies: one multi-tensor kernel

    def _capture_optimizer(self) -> None:
        self._opt_sig = _hyper_signature(self.optimizer)
        if self.opt_eager or (self._flat_sgd and self.world > 1):          # the multi-GPU flat path launches its pieces itself
            self.opt_graph = None
            return
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=self.graphs[0].pool(), capture_error_mode=_capture_mode()):
            self._optimizer_part()
        self.opt_graph = g

    def sync_weights(self) -> None:
        """re-read the fp32 parameters into the 16-bit shadow copies: call after changing the weights from outside the step
        (``load_state_dict``, manual edits); the step itself keeps them in sync."""
        if self._shadows:
            with torch.no_grad():
                torch._foreach_copy_(self._shadows, self._shadow_params)

    def _piece_ranges(self):
        return piece_ranges(self._flat.numel(), self._sgd_chunk if self._flat_sgd else 2048, self._stage_start if self._split else None, self._pieces)

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
            if self._flat_sgd:
                self._refresh_hyper()                           # device table: nothing to re-capture
            else:
                self._sig_changes += 1
                if self._sig_changes >= 3 and not self.opt_eager:
                    self.opt_eager = True                       # changes every few steps (poly / cosine): stop re-capturing
                torch.cuda.current_stream(self.static_x.device).synchronize()
                self._capture_optimizer()
        if self.world > 1:
            import torch.distributed as dist
            ar = lambda lo, hi, grp=None: dist.all_reduce(self._flat[lo:hi], op=dist.ReduceOp.AVG, group=grp if grp is not None else self.group, async_op=True)   # noqa: E731
            stages = self._piece_ranges()
            works = []
            for k, g in enumerate(self.graphs):
                g.replay()
                # collectives of all but the last stage run BESIDE the next graph: capped communicator; the last one is exposed: full width
                grp = self.overlap_group if (self._split and k + 1 < len(self.graphs)) else None
                works += [(ar(*er, grp), cr) for er, cr in stages[k]]        # queued on the communicator's stream, one after the other
            if self._flat_sgd and self.after_backward is None:
                for w, (c0, nc) in works:
                    w.wait()                                     # the compute stream waits for THIS piece only
                    self._sgd_range(c0, nc)
                if self._flat_s is None and self._shadows:
                    with torch.no_grad():
                        torch._foreach_copy_(self._shadows, self._shadow_params)
                return self.losses[3]
            for w, _ in works:
                w.wait()
        else:
            for g in self.graphs:                                # one graph (more only when a test forces the split schedule on one GPU)
                g.replay()
        if self.opt_graph is not None:
            self.opt_graph.replay()
        else:
            self._optimizer_part()
        return self.losses[3]

    # kept for callers of the first version
    @property
    def graph(self):
        return self.graphs[0] if self.graphs else None
