#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: ACAN NYU-shape images/sec on N B200s, CAM attention TFLOP/s vs peak.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

Workload = BASELINE.json configs[1]: full ACAN forward (dilated ResNet-101 backbone + context aggregation module +
ordinal soft inference), NYU-shaped 256x352 synthetic batch of 16 images per GPU (weak scaling, no data-path
collective: every image is an independent unit).  One step = net(images) -> net.inference(y), the
reference's eval call pattern (depthest_trainer.py:184-185), with the drop-in modules of acan_b200/.

  value     images/s with the batch already resident in HBM (CUDA events over EXACTLY K steps, max over ranks); `timed_1s` repeats the
            measurement over a region of at least one second, with the SM clock sampled throughout.
  e2e       the same call with HOST buffers: pinned images -> H2D, forward, inference, depth -> D2H, every step.
  parity    the benched network's depth for the first images of the timed batch against the reference's CPU path on the same weights
            (N = 1): the run FAILS above the tolerance written there.
  roofline  the dominant hand-written kernel (aggregation pass O = P V on tcgen05), its CUDA-event time measured live in this run via the
            library's event hooks, against MEASURED_PEAKS.json; `attention_core`, `head`, `tail` give the same arithmetic for the whole
            fused CAM core, the depth head and the fused classifier tail.
  train     configs[3] / [4] sub-record: the beta = 1 training step (forward + losses + backward + NCCL gradient all-reduce + SGD), B = 8
            per GPU, with the all-reduce time and a replica-consistency check (--no-train skips it; --workload train prints it alone).
  gpu_reference  (N = 1) the UNMODIFIED reference modules (baseline/_ref) on the same B200, fp32 eager and bf16 autocast: configs[1]
            forward, the CAM module alone, the configs[3] training step -- the like-for-like GPU baseline.
  c3_kitti_hard  (N = 1) configs[2]: KITTI 160x512, batch 32, hard ordinal inference, with the count of bin indices that differ from
            ATen's fp32 upsample -> decode -> count path on the same logits.
  cpu_baseline   the reference's CPU path on the host cores, bounded sample (rank 0, N = 1).
--impl reference times that CPU path alone: the reference's own modules (baseline/_ref, kind "reference"; the oracle port if absent).
"""
from __future__ import annotations

import argparse
import ctypes
import faulthandler
import gc
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "baseline")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

H, W, BATCH = 256, 352, 16
DMIN, DMAX, K = 0.72, 10.0, 80
LAYERS101 = [3, 4, 23, 3]
AMP = torch.float16      # 16-bit autocast for the cuDNN convolutions; fp16 lets the CAM consume F and V in place
WORKLOAD = "configs[1]: full ACAN fwd (ResNet-101 + CAM + OR soft inference), NYU 256x352, batch 16 per GPU"
# benched arithmetic = fp16 activations through 101 folded-BN layers: max-norm relative depth error vs the fp32 reference; the fp32
# drop-in network holds 1e-3 (tests/test_gpu_fullsize.py), DESIGN.md section 4 explains the split
PARITY_TOL_MAX, PARITY_TOL_MEAN = 2.5e-1, 4e-2
PROGRESS = {"phase": "not started", "t": 0.0}      # last milestone of the training sub-record (reported if its deadline passes)


def milestone(name):
    PROGRESS["phase"], PROGRESS["t"] = name, time.time()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:          # multi-GPU: a hang must be attributable from stderr
        print(f"[bench rank {os.environ.get('RANK', '0')} {time.strftime('%H:%M:%S')}] train: {name}", file=sys.stderr, flush=True)


TRAIN_DEADLINE_S = 300   # N > 1: wall-clock bound on the training sub-record of the default line (it normally takes 40-90 s)
REF_SAMPLE_BATCH = 4     # images per step of the CPU reference arm (a bounded sample of the 16-image step: ~1-3 s of host time per step)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm": d["hbm_gbs"], "tensor": d["bf16_tflops"], "tensor_sustained": d.get("bf16_tflops_sustained"), "src": "measured"}
    return {"hbm": 6650.0, "tensor": 1590.0, "tensor_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """SM clock / throttle reasons sampled every 20 ms during the timed region through NVML (nvidia_ml_py); the
    nvidia-smi query of the profiling recipe reads the same counters, but its stdout is block-buffered in a pipe."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.sm, self.mask, self.max_mhz, self.err = [], 0, None, None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception as e:  # pragma: no cover
            self.err = f"NVML unavailable: {e}"

    def _pump(self):
        while not self._stop.is_set():
            try:
                self.sm.append(float(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception as e:  # pragma: no cover
                self.err = str(e)
                return
            time.sleep(0.02)

    def stop(self):
        self._stop.set()
        if self.err and not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "samples": 0, "reasons": [self.err]}
        self.thread.join(timeout=1.0)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                "reasons": sorted(name for bit, name in self.REASONS.items() if self.mask & bit)}


def build_net(device, hh=H, ww=W, batch=BATCH, inference="soft", dmin=DMIN, dmax=DMAX):
    from acan_b200.network import ResNet
    torch.manual_seed(2019)
    net = ResNet(dmin, dmax, K, "OR", inference, "attention", hh, ww, alpha=0, beta=0, layers=LAYERS101)
    net = net.to(device).to(memory_format=torch.channels_last)
    # BN calibration (never eval() on raw running statistics): one train()-mode pass with momentum 1
    # puts the batch statistics of a synthetic batch into the running buffers, then eval().
    bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for m in bns:
        m.momentum = 1.0
    net.train()
    for m in net.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=AMP):
        net(torch.rand(batch, 3, hh, ww, device=device).contiguous(memory_format=torch.channels_last))
    for m in bns:
        m.momentum = 0.1
    from acan_b200.inference import optimize_for_inference
    net.eval()
    fast = optimize_for_inference(net)               # BN folded into the convolutions, fused conv epilogues, fp16 channels_last
    fast.source_state = {k: v.detach().float().cpu() for k, v in net.state_dict().items()}   # for the CPU baseline (same weights)
    return fast


def step_device(net, images):
    return net.inference(net(images))                # the reference's eval call pattern (depthest_trainer.py:184-185)


# ------------------------------------------------------------------------------------ the reference's CPU path

def cpu_reference_callable(sd_cpu=None, hh=H, ww=W, calibrate_batch=0):
    """(fn(images) -> depth, kind): the UNMODIFIED reference network (baseline/_ref) in eval mode on the host, loaded with `sd_cpu`
    (or, with sd_cpu None, the reference's own random init under seed 2019 followed by the BN calibration pass of SURVEY section 8d);
    the oracle port of the same forward when no copy of the reference is present."""
    try:
        import refload
        ref = refload.load_reference()
    except Exception:
        ref = None
    if ref is not None:
        torch.manual_seed(2019)
        net = ref.models.create_network(refload.reference_args(height=hh, width=ww))
        if sd_cpu is not None:
            net.load_state_dict(sd_cpu)
        elif calibrate_batch:
            bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
            for m in bns:
                m.momentum = 1.0
            net.train()
            for m in net.modules():
                if isinstance(m, torch.nn.Dropout2d):
                    m.eval()
            with torch.no_grad():
                net(torch.rand(calibrate_batch, 3, hh, ww))
            for m in bns:
                m.momentum = 0.1
        net.eval()

        def fn(x):
            with torch.no_grad():
                return net.inference(net(x))
        return fn, "reference"
    from oracle import acan_oracle as O
    if sd_cpu is None:
        from acan_b200.network import ResNet
        torch.manual_seed(2019)
        sd_cpu = {k: v.detach() for k, v in ResNet(DMIN, DMAX, K, "OR", "soft", "attention", hh, ww, layers=LAYERS101).state_dict().items()}

    def fn(x):
        with torch.no_grad():
            return O.inference(O.acan_forward(x, sd_cpu, LAYERS101, "OR", training=False)["y"], "OR", "soft", DMIN, DMAX, K)
    return fn, "port"


def time_cpu(fn, batch, n_iter, warm):
    x = torch.rand(batch, 3, H, W)
    times = []
    for i in range(warm + n_iter):
        t0 = time.perf_counter()
        fn(x)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return times


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, on the arm's config (B = 16 per step)."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    fn, kind = cpu_reference_callable(None, calibrate_batch=4)
    times = time_cpu(fn, REF_SAMPLE_BATCH, args.steps, args.warmup)
    tot = sum(times)
    val = REF_SAMPLE_BATCH * len(times) / tot
    sample = (f"bounded sample: {REF_SAMPLE_BATCH} of the step's {BATCH} images per step x {len(times)} steps ({tot:.1f} s), fp32, torch CPU ops on {threads} threads, eval-mode BN on calibrated statistics; "
              + ("the unmodified reference modules (baseline/_ref: models.create_network -> net.inference(net(x)))" if kind == "reference"
                 else "oracle port of the reference forward (no copy of the reference on this host)"))
    print(json.dumps({
        "impl": "reference", "metric": "ACAN NYU-shape forward images/sec", "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / len(times) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ------------------------------------------------------------------------------------ helpers

def ev_time(fn_, iters=10, warm=3):
    for _ in range(warm):
        fn_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn_()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e-3 / iters


def free_cuda():
    gc.collect()
    torch.cuda.empty_cache()


TRAIN_BATCH = 8
TRAIN_WORKLOAD = "configs[3]: ACAN training step with attention loss (beta=1, alpha=0), fwd + loss + bwd + SGD, NYU 256x352, batch 8 per GPU"


def run_train(args, rank, world, device, lib):
    """configs[3] (and configs[4] under torchrun): ResNet-101 + CAM + fused ordinal tail loss + fused attention loss, forward + backward +
    SGD(momentum) step; N > 1 averages gradients with NCCL (acan_b200.training.GraphedTrainStep: the late layers' segment of the flat
    gradient buffer is all-reduced while the early layers' backward graph replays).  bf16 autocast, channels_last; the attention core runs
    fp16 probabilities x bf16 values / gradients in place (fp32 accumulate), the losses fp32.  Returns the record (rank 0) or None."""
    import torch.distributed as dist
    from acan_b200 import ops
    from acan_b200.modules import AttentionLoss2d
    from acan_b200.network import OrdinalRegression2d, ResNet
    from acan_b200.parallel import GradAllReducer, capped_group
    from acan_b200.training import GraphedTrainStep
    B = TRAIN_BATCH
    milestone("building the network")
    torch.manual_seed(2019)
    net = ResNet(DMIN, DMAX, K, "OR", "soft", "attention", H, W, alpha=0, beta=1.0, layers=LAYERS101).to(device)
    net = net.to(memory_format=torch.channels_last).train()
    crit = net.LossFunc(DMIN, DMAX, K, AppearanceLoss=OrdinalRegression2d(ignore_index=0), AttentionLoss=AttentionLoss2d(scale=1), alpha=0.0, beta=1.0)
    opt = torch.optim.SGD(net.parameters(), lr=1e-4, momentum=0.9, weight_decay=5e-4)
    reducer = GradAllReducer(net.parameters(), bucket_mb=32.0)
    n_rot = 8
    torch.manual_seed(2019 + rank)
    host_x = [torch.rand(B, 3, H, W).contiguous(memory_format=torch.channels_last).pin_memory() for _ in range(n_rot)]
    host_y = [(torch.rand(B, 1, H, W) * 10 + 0.5).mul_((torch.rand(B, 1, H, W) > 0.1).float()).pin_memory() for _ in range(n_rot)]   # ~10 % invalid (0)
    dev_x = [t.to(device, non_blocking=True) for t in host_x]
    dev_y = [t.to(device, non_blocking=True) for t in host_y]

    graph_note = None
    milestone("GraphedTrainStep: eager warm-up steps (per-parameter all-reduces) and CUDA-graph captures")
    try:
        step = GraphedTrainStep(net, crit, opt, dev_x[0], dev_y[0], amp_dtype=torch.bfloat16, reducer=reducer if world > 1 else None,
                                graph=not args.no_graph, overlap=args.split_backward, pieces=args.pieces, nccl_sms=args.nccl_sms,
                                overlap_group=capped_group(args.nccl_sms) if (args.two_communicators and args.split_backward and world > 1) else None)
    except Exception as e:          # same kernels either way: a failed capture must not cost the measurement
        torch.cuda.synchronize()
        graph_note = f"eager (CUDA graph capture failed: {type(e).__name__}: {str(e)[:160]})"
        print(f"[bench rank {rank}] {graph_note}", file=sys.stderr, flush=True)
        step = None
    milestone("captures done on this rank (ok)" if step is not None else "capture FAILED on this rank")
    if world > 1:                   # every rank must run the SAME collective schedule: one failed capture sends all of them to the eager path
        ok = torch.tensor([0 if step is None else 1], device=device, dtype=torch.int32)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok) == 0 and step is not None:
            graph_note = "eager (CUDA graph capture failed on another rank)"
            step = None
    if step is None:
        reducer.enabled = True
        step = GraphedTrainStep(net, crit, opt, dev_x[0], dev_y[0], amp_dtype=torch.bfloat16, reducer=reducer if world > 1 else None, graph=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    milestone("first replayed steps (stage all-reduces + SGD)")
    for i in range(max(args.warmup, 3)):
        step(dev_x[i % n_rot], dev_y[i % n_rot])
    barrier()
    milestone("timed steps")
    launches0 = lib.acan_launch_count()
    sampler = ClockSampler(device.index)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(dev_x[i % n_rot], dev_y[i % n_rot])
    e1.record()
    barrier()
    elapsed = e0.elapsed_time(e1) * 1e-3
    # >= 1 s region
    long_steps = max(args.steps, int(math.ceil(1.2 / max(elapsed / args.steps, 1e-4))))
    e0.record()
    for i in range(long_steps):
        loss = step(dev_x[i % n_rot], dev_y[i % n_rot])
    e1.record()
    barrier()
    elapsed_long = e0.elapsed_time(e1) * 1e-3
    clocks = sampler.stop()
    launches = lib.acan_launch_count() - launches0
    if step.graphs:                                  # replays do not pass through the library's host-side launch counter
        launches = step.library_launches_per_step * args.steps
    else:
        launches = launches * args.steps // (args.steps + long_steps)
    assert bool(torch.isfinite(loss)), "training loss is not finite"
    # end to end: pinned host images + labels in, the loss value out, every step
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        loss_host = float(step(host_x[i % n_rot], host_y[i % n_rot]))     # pinned host -> the step's static device buffers, loss -> host
    barrier()
    e2e_elapsed = time.perf_counter() - t0
    # the gradient all-reduce on its own (what the split schedule hides / exposes): the whole flat buffer, and the exposed early segment
    ar_ms = ar_early_ms = None
    n_grad = sum(p.numel() for p in net.parameters() if p.requires_grad)
    if world > 1:
        flat = getattr(step, "_flat", None)
        if flat is None:
            flat = torch.zeros(n_grad, device=device)
        keep = flat.clone()
        ar_ms = ev_time(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG), iters=10) * 1e3
        ne = getattr(step, "_stage_start", {}).get(1, getattr(step, "_n_early", n_grad)) if getattr(step, "_split", False) else getattr(step, "_n_early", n_grad)
        ar_early_ms = ev_time(lambda: dist.all_reduce(flat[:ne], op=dist.ReduceOp.AVG), iters=10) * 1e3
        flat.copy_(keep)
    t = torch.tensor([elapsed, e2e_elapsed, elapsed_long], device=device, dtype=torch.float64)
    in_sync = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # averaged gradients keep the replicas' weights identical: compare a checksum of all parameters across ranks
        chk = torch.stack([p.detach().double().sum() for p in net.parameters()]).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool((hi - lo).abs() <= 1e-9 * hi.abs())
    elapsed, e2e_elapsed, elapsed_long = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        return None
    # rooflines of the step's hand-written kernel groups, stand-alone at the step's shapes
    pk = peaks()
    n = (H // 8) * (W // 8)
    cl = torch.channels_last
    f = torch.relu(torch.randn(B, 512, H // 8, W // 8, device=device)).half().contiguous(memory_format=cl)
    v = torch.relu(torch.randn(B, 2048, H // 8, W // 8, device=device)).bfloat16().contiguous(memory_format=cl)
    lab = torch.randint(0, K, (B, 1, H, W), device=device).float()
    go = (torch.randn(B, 2048, H // 8, W // 8, device=device) * 1e-3).bfloat16().contiguous(memory_format=cl)
    gl = torch.ones((), device=device)
    o = ops.attention_forward_train(f, v)
    ops.attention_loss_fused(o["ws"], lab)
    t_f = ev_time(lambda: ops.attention_forward_train(f, v))
    t_l = ev_time(lambda: ops.attention_loss_fused(o["ws"], lab))
    t_b = ev_time(lambda: ops.attention_backward_cl(o["ws"], o["value16"], o["ctx32"], go, gl, torch.float16, torch.bfloat16))
    fwd_flops, bwd_flops = B * 5120.0 * n * n, B * 10240.0 * n * n
    del f, v, lab, go, o
    # the dominant hand-written kernels of the step by time are the BN (+ReLU/+residual) passes (105 layers): the backward kernel at the
    # most common large layer shape (layer3's 1024-channel maps), operands rotated through > 2x L2, replayed from a CUDA graph so that the
    # number is device time, not host launch rate
    bc, bh, bw = 1024, H // 8, W // 8
    n_rot = 14
    xs = [torch.randn(B, bc, bh, bw, device=device, dtype=torch.bfloat16).contiguous(memory_format=cl) for _ in range(n_rot)]
    gs = [torch.randn(B, bc, bh, bw, device=device, dtype=torch.bfloat16).contiguous(memory_format=cl) for _ in range(n_rot)]
    wgt, bia = torch.ones(bc, device=device), torch.zeros(bc, device=device)
    outs = [ops.bn_train_forward(xx, wgt, bia, None, None, 0.1, 1e-5, True, None) for xx in xs]
    ys_, mean_, invstd_ = [t_[0] for t_ in outs], outs[0][1], outs[0][2]

    def graph_time(calls, reps=10):
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for c_ in calls[:2]:
                c_()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize()
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            keep = [c_() for c_ in calls]
        for _ in range(2):
            g_.replay()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            g_.replay()
        b_.record()
        torch.cuda.synchronize()
        del keep
        return a.elapsed_time(b_) * 1e-3 / (reps * len(calls))
    t_bn_f = graph_time([lambda xx=xx: ops.bn_train_forward(xx, wgt, bia, None, None, 0.1, 1e-5, True, None) for xx in xs])
    t_bn_b = graph_time([lambda k=k: ops.bn_train_backward(xs[k], ys_[k], gs[k], wgt, mean_, invstd_, False) for k in range(n_rot)])
    bn_elems = float(B * bc * bh * bw)
    if world > 1:
        if getattr(step, "_split", False):
            sched = ("dp%d: batch-sharded replicas; three graphs (forward + losses + backward of layer4/decoder/classifier | backward of layer3 | backward of "
                     "layer2..stem); the flat fp32 gradient slice of each stage is averaged by NCCL BESIDE the next graph (communicator capped to %d CTAs, the "
                     "BN grids of graphs 2-3 sized for the remaining SMs), only the last stage's %.1f MB all-reduce is exposed; then the one-pass SGD kernel per stage"
                     % (world, args.nccl_sms, step._stage_start.get(1, 0) * 4 / 1e6))
        else:
            sched = ("dp%d: batch-sharded replicas; forward + losses + backward graph, then the flat fp32 gradient buffer is averaged in %d pieces on the (high-priority) "
                     "NCCL stream, the one-pass SGD kernel updating piece k while piece k+1 is in flight" % (world, args.pieces))
    else:
        sched = "dp1"
    return {
        "metric": "ACAN NYU-shape training images/sec", "value": world * B * args.steps / elapsed, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "timed_1s": {"steps": long_steps, "seconds": elapsed_long, "ms_per_step": elapsed_long / long_steps * 1e3, "value": world * B * long_steps / elapsed_long},
        "config": {"workload": TRAIN_WORKLOAD, "parallelism": sched,
                   "precision": "bf16 autocast channels_last cuDNN convolutions; fused BN/ReLU kernels bf16 in / fp32 statistics; attention fwd/bwd on the channels-last path: fp16 probabilities x bf16 values / gradients read in place, fp32 accumulate; losses fp32; SGD momentum on fp32 weights",
                   "launch": graph_note or ("eager" if args.no_graph else "step replayed from CUDA graphs (acan_b200.training.GraphedTrainStep): forward + losses + backward; SGD as one flat streaming kernel (acan_sgd_flat) with its hyper-parameters in a device table"),
                   "l2": f"{8} rotating input batches; activations (> 10 GB per step) exceed the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": world * B * args.steps / e2e_elapsed, "unit": "images/s", "h2d_bytes_per_step": B * 4 * H * W * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_elapsed / args.steps * 1e3},
        "gpu_launches": int(launches),
        "allreduce_ms": ar_ms, "allreduce_exposed_segment_ms": ar_early_ms,
        "allreduce_note": (f"timed on the step's communicator, capped to {args.nccl_sms} CTAs so that it can run beside the backward graphs (uncapped: 0.70 ms for the whole buffer at 2 GPUs)"
                           if (world > 1 and args.split_backward and not args.two_communicators) else None), "gradient_bytes": n_grad * 4, "replicas_in_sync": in_sync,
        "roofline": {"bound": "hbm", "kernel": "bn_bwd_coop_kernel<bf16> (BN + ReLU backward: one cooperative launch, 14 B per element algorithmic: dy, y, x read for the "
                               "statistics, read again -- from L2 -- for dx, dx written), layer3 shape [8,1024,32,44], stand-alone, graph-replayed over 14 rotating operand sets",
                     "achieved": 14.0 * bn_elems / t_bn_b / 1e9, "peak": pk["hbm"], "unit": "GB/s", "frac": 14.0 * bn_elems / t_bn_b / 1e9 / pk["hbm"],
                     "traffic": 76.1e6, "peak_source": pk["src"] + " (copy)", "launch_us": t_bn_b * 1e6,
                     "note": "traffic = dram read 71.5 MB + write 4.6 MB from profiles/r01b_train_kernels_bn_tailloss.ncu-rep (the 23 MB of dx are still in L2 when the kernel ends)"},
        "bn_forward": {"bound": "hbm", "what": "bn_fwd_coop_kernel<bf16>, same shape, 6 B per element algorithmic", "achieved": 6.0 * bn_elems / t_bn_f / 1e9, "peak": pk["hbm"],
                       "unit": "GB/s", "frac": 6.0 * bn_elems / t_bn_f / 1e9 / pk["hbm"], "launch_us": t_bn_f * 1e6},
        "attention_bwd": {"bound": "tensor", "what": "acan_attn_bwd_nhwc (delta, dV, dP -> dS with the attention-loss term, symmetrised dF: four tcgen05 launches, operands in place, "
                                                     "probabilities kept from the forward), stand-alone at the step's shape, algorithmic 10240*N^2 FLOP/img",
                          "achieved": bwd_flops / t_b / 1e12, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": bwd_flops / t_b / 1e12 / pk["tensor"], "launch_us": t_b * 1e6},
        "attention_fwd_train": {"bound": "tensor", "what": "acan_attn_fwd_nhwc keep=1 (statistics + probability + aggregation passes, fp32 context), algorithmic 5120*N^2 FLOP/img",
                                "achieved": fwd_flops / t_f / 1e12, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": fwd_flops / t_f / 1e12 / pk["tensor"], "launch_us": t_f * 1e6},
        "attention_loss_us": t_l * 1e6, "loss": loss_host,
    }


# ------------------------------------------------------------------------------------ N = 1 extras

def gpu_reference_extras(device, sd_cpu, ours_c2_ms, ours_train_ms, ours_cam_us):
    """the UNMODIFIED reference modules on this GPU (depthest_main.py:59-77 settings: cudnn.benchmark on, PyTorch's default TF32 policy):
    configs[1] forward + inference, the CAM module alone, the configs[3] training step; fp32 eager and bf16 autocast."""
    try:
        import refload
        ref = refload.load_reference()
    except Exception as e:
        return {"unavailable": str(e)[:200]}
    out = {"what": "unmodified reference modules (baseline/_ref) on the same B200, eager PyTorch, cudnn.benchmark=True, CUDA events"}
    torch.manual_seed(0)
    net = ref.models.create_network(refload.reference_args(height=H, width=W, beta=1.0))
    net.load_state_dict(sd_cpu)
    net = net.to(device).eval()
    x = torch.rand(BATCH, 3, H, W, device=device)
    feat = torch.relu(torch.randn(BATCH, 2048, H // 8, W // 8, device=device))
    with torch.no_grad():
        t32 = ev_time(lambda: net.inference(net(x)), iters=5, warm=2)
        c32 = ev_time(lambda: net.decoder(feat), iters=10, warm=2)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            t16 = ev_time(lambda: net.inference(net(x)), iters=5, warm=2)
            c16 = ev_time(lambda: net.decoder(feat), iters=10, warm=2)
    out["c2_forward_ms"] = {"fp32": t32 * 1e3, "bf16_autocast": t16 * 1e3, "ours": ours_c2_ms,
                            "speedup_vs_fp32": t32 * 1e3 / ours_c2_ms, "speedup_vs_bf16": t16 * 1e3 / ours_c2_ms}
    out["cam_module_us"] = {"fp32": c32 * 1e6, "bf16_autocast": c16 * 1e6, "ours": ours_cam_us,
                            "speedup_vs_fp32": c32 * 1e6 / ours_cam_us, "speedup_vs_bf16": c16 * 1e6 / ours_cam_us,
                            "what": "SADecoder.forward on [16,2048,32,44] features (projections + attention + pooling branch + merge)"}
    del x, feat
    if ours_train_ms is not None:
        net.train()
        crit = ref.models.create_lossfunc(refload.reference_args(height=H, width=W, beta=1.0), net)
        opt = torch.optim.SGD(net.parameters(), lr=1e-4, momentum=0.9, weight_decay=5e-4)
        xb = torch.rand(TRAIN_BATCH, 3, H, W, device=device)
        yb = (torch.rand(TRAIN_BATCH, 1, H, W, device=device) * 10 + 0.5) * (torch.rand(TRAIN_BATCH, 1, H, W, device=device) > 0.1).float()

        def one(amp):
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                loss = crit(net(xb), yb.clone(), 1)[3]
            loss.backward()
            opt.step()
        s32 = ev_time(lambda: one(False), iters=4, warm=2)
        s16 = ev_time(lambda: one(True), iters=4, warm=2)
        out["c4_train_step_ms"] = {"fp32": s32 * 1e3, "bf16_autocast": s16 * 1e3, "ours": ours_train_ms,
                                   "speedup_vs_fp32": s32 * 1e3 / ours_train_ms, "speedup_vs_bf16": s16 * 1e3 / ours_train_ms}
    return out


def fp32_parity_arm(device, sd_cpu, x_par, want, steps):
    """the same network in strict fp32 (TF32 off): BN folded, fused conv epilogues, NCHW, the CAM through the drop-in module's generic path
    (fp16 tensor-core operands, fp32 in / out).  Its depth holds the north star's 1e-3 against the reference where 101 layers of fp16
    activations cannot; its images/s is the price.  Also: the drop-in network under PyTorch's DEFAULT GPU numerics for the reference
    (TF32 convolutions), for scale."""
    from acan_b200.inference import optimize_for_inference
    from acan_b200.network import ResNet
    keep = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        net = ResNet(DMIN, DMAX, K, "OR", "soft", "attention", H, W, alpha=0, beta=0, layers=LAYERS101)
        net.load_state_dict(sd_cpu)
        net = net.to(device).eval()
        f32 = optimize_for_inference(net, torch.float32)
        rel = lambda got: (float((got.float().cpu() - want).abs().max() / want.abs().max()), float(((got.float().cpu() - want).abs() / want.abs()).mean()))   # noqa: E731
        xp = x_par.to(device)
        x16 = torch.rand(BATCH, 3, H, W, device=device)
        with torch.no_grad():
            e_max, e_mean = rel(f32.predict_depth(xp))
            t = ev_time(lambda: f32.predict_depth(x16), iters=max(3, min(steps, 5)), warm=2)
            torch.backends.cudnn.allow_tf32 = True
            t_max, t_mean = rel(net.inference(net(xp)))
            torch.backends.cudnn.allow_tf32 = False
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = keep
    del net, f32
    free_cuda()
    return {"what": "acan_b200.inference.optimize_for_inference(net, torch.float32), TF32 off, eager launches, batch 16", "value": BATCH / t, "unit": "images/s",
            "ms_per_step": t * 1e3, "depth_rel_err_max_norm": e_max, "depth_rel_err_mean": e_mean, "tolerance_max_norm": 1e-3,
            "tf32_convolutions_for_scale": {"what": "fp32 drop-in network with cudnn.allow_tf32=True (PyTorch's default numerics for the reference on a GPU)",
                                            "depth_rel_err_max_norm": t_max, "depth_rel_err_mean": t_mean}}


def c3_line(device, steps):
    """BASELINE.json configs[2]: KITTI 160x512, batch 32, hard ordinal inference; device-resident images/s from the CUDA-graph replay, and the
    bin indices of the fused tail against ATen's fp32 upsample_bilinear2d -> decode_ord -> count(p > 0.5) on the same stride-8 logits."""
    from acan_b200 import ops
    from acan_b200.inference import GraphedPredictor
    hh, ww, b = 160, 512, 32
    dmin, dmax = 1.98, 80.0
    net = build_net(device, hh, ww, b, "hard", dmin, dmax)
    xs = [torch.rand(b, 3, hh, ww, device=device).contiguous(memory_format=torch.channels_last) for _ in range(4)]
    pred = GraphedPredictor(net, xs[0])
    for i in range(3):
        pred.predict(xs[i % 4])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(steps):
        pred.predict(xs[i % 4])
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    with torch.no_grad():
        z = net(xs[0])["y"]._z.float()
        _, idx = ops.tail_inference(z, "OR", "hard", dmin, dmax, K, want_index=True)
        up = torch.nn.functional.interpolate(z.contiguous(), scale_factor=8, mode="bilinear", align_corners=True).view(b, K, 2, hh, ww)
        e = torch.exp(up)
        want = (e[:, :, 1] / (e[:, :, 0] + e[:, :, 1]) > 0.5).float().sum(1, keepdim=True)
        bad = int((idx.float() != want).sum())
    del pred, net, xs, up, e, want
    free_cuda()
    return {"workload": "configs[2]: KITTI 160x512, batch 32, ResNet-101 + CAM + OR hard inference", "value": b * steps / sec, "unit": "images/s",
            "ms_per_step": sec / steps * 1e3, "steps": steps,
            "hard_index_mismatches_vs_aten_fp32": bad, "pixels": b * hh * ww}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the steps eagerly instead of replaying their CUDA graphs")
    ap.add_argument("--no-split-backward", dest="split_backward", action="store_false",
                    help="train, N > 1: one backward graph and a pipelined all-reduce + SGD instead of the default three backward graphs whose "
                         "stage all-reduces run beside the next graph (A/B)")
    ap.add_argument("--two-communicators", action="store_true",
                    help="split backward: overlapped all-reduces on a second, capped communicator and the exposed one on an uncapped default communicator "
                         "(measured equal at 2 GPUs; the default is one capped communicator: fewer moving parts)")
    ap.add_argument("--nccl-sms", type=int, default=16, help="split backward: CTAs of the communicator of the overlapped all-reduces = SMs the BN grids of graphs 2-3 leave free")
    ap.add_argument("--nccl-min-ctas", type=int, default=0, help="N > 1: lower bound on the CTAs of a collective (ncclConfig_t.minCTAs); 0 = NCCL's default")
    ap.add_argument("--pieces", type=int, default=4, help="train, N > 1: pieces of the flat gradient buffer (all-reduce of piece k+1 runs beside the update of piece k)")
    ap.add_argument("--no-train", action="store_true", help="skip the configs[3]/[4] training sub-record")
    ap.add_argument("--no-long", action="store_true", help="skip the >= 1 s repetition of the timed region (runs under ncu)")
    ap.add_argument("--no-extras", action="store_true", help="skip the N = 1 extras (reference on the GPU, configs[2] line)")
    ap.add_argument("--workload", default="infer", choices=["infer", "train"],
                    help="infer (default): BASELINE.json configs[1], the headline (+ the train sub-record).  train: only configs[3]/[4]")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    knobs = sorted(k for k in os.environ if k.startswith("ACAN_"))
    if knobs:
        sys.exit(f"bench.py refuses to run with experiment variables set ({', '.join(knobs)}): the shipped library ignores them, and a number "
                 "measured under an experiment build is not a bench value")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from acan_b200 import _lib
    from acan_b200.parallel import init_distributed
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    # split backward (default): ONE communicator capped to --nccl-sms CTAs -- the overlapped stage all-reduces must leave the BN grids their SMs,
    # and the only exposed collective is the last stage's 5.8 MB, which does not need more
    init_distributed("nccl" if world > 1 else None, max_ctas=args.nccl_sms if (args.split_backward and not args.two_communicators) else None,
                     min_ctas=args.nccl_min_ctas or None)
    lib = _lib.load()
    _lib.check(lib.acan_device_check(), "acan_device_check")
    torch.backends.cudnn.benchmark = True                 # as depthest_main.py:65
    if args.workload == "train":
        rec = run_train(args, rank, world, device, lib)
        if rank == 0:
            rec["cpu_baseline"] = None
            print(json.dumps(rec))
        if world > 1:
            dist.destroy_process_group()
        return

    net = build_net(device)
    n_rot = 8                                             # 8 distinct input batches = 138 MB > 126 MB L2
    host = [torch.rand(BATCH, 3, H, W).contiguous(memory_format=torch.channels_last).pin_memory() for _ in range(n_rot)]
    dev_in = [h.to(device, non_blocking=True) for h in host]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------------------------------------------
    # default: the step replayed from a CUDA graph (acan_b200.inference.GraphedPredictor); --no-graph launches it eagerly
    from acan_b200.inference import GraphedPredictor, StreamedPredictor
    pred, launch_note = None, "eager"
    if not args.no_graph:
        try:
            pred = GraphedPredictor(net, dev_in[0])
            launch_note = "step replayed from a CUDA graph (acan_b200.inference.GraphedPredictor); roofline brackets from an eager pass over the same steps"
        except Exception as e:      # same kernels either way: a failed capture must not cost the measurement
            torch.cuda.synchronize()
            launch_note = f"eager (CUDA graph capture failed: {type(e).__name__}: {str(e)[:120]})"
    run = (lambda x: step_device(net, x)) if pred is None else pred.predict
    for i in range(args.warmup):
        run(dev_in[i % n_rot])
    barrier()
    launches0 = lib.acan_launch_count()
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        depth = run(dev_in[i % n_rot])
    e1.record()
    barrier()
    elapsed = e0.elapsed_time(e1) * 1e-3
    launches = lib.acan_launch_count() - launches0 if pred is None else pred.library_launches_per_step * args.steps
    clocks_k = sampler.stop()
    depth_first = run(dev_in[0]).clone()                   # for the parity record

    # ---- per-kernel CUDA-event brackets (roofline): the same K steps on the same inputs, launched eagerly with the library's
    #      event hooks on (events cannot be timed from inside a replayed graph); not part of `value`
    for i in range(2):
        step_device(net, dev_in[i % n_rot])
    barrier()
    _lib.check(lib.acan_profile_begin(args.steps * 16 + 64), "acan_profile_begin")
    for i in range(args.steps):
        with torch.cuda.nvtx.range("acan_step"):          # what the committed ncu launch list is filtered on (profiles/)
            step_device(net, dev_in[i % n_rot])
    barrier()
    ms = (ctypes.c_float * 8)()
    cnt = (ctypes.c_int * 8)()
    _lib.check(lib.acan_profile_end(ms, cnt), "acan_profile_end")
    # ... and once more with ONE bracket around the whole attention call of every step (no events between its passes)
    _lib.check(lib.acan_profile_begin_coarse(args.steps * 8 + 64), "acan_profile_begin_coarse")
    for i in range(args.steps):
        step_device(net, dev_in[i % n_rot])
    barrier()
    ms_c = (ctypes.c_float * 8)()
    cnt_c = (ctypes.c_int * 8)()
    _lib.check(lib.acan_profile_end(ms_c, cnt_c), "acan_profile_end")

    # ---- the device-resident measurement again over a region of at least one second (the K-step region is ~0.1 s: too short for the clock
    #      sampler, and short enough to run at boost clocks); after the per-kernel brackets, which are taken on the still-cool device
    long_steps = args.steps if args.no_long else max(args.steps, int(math.ceil(1.2 / max(elapsed / args.steps, 1e-4))))
    sampler = ClockSampler(local)
    e0.record()
    for i in range(long_steps):
        depth = run(dev_in[i % n_rot])
    e1.record()
    barrier()
    elapsed_long = e0.elapsed_time(e1) * 1e-3
    clocks = sampler.stop()
    # ---- end to end: host buffers in, host depth out, every step: the upload of batch i+1 and the download of depth i-1 ride
    #      on copy streams while batch i computes; every step still moves its own 17.3 MB in and 5.8 MB out, and the timed
    #      region ends when the LAST depth map is in host memory
    if pred is None:
        pred = StreamedPredictor(net, device)
    for i in range(3):
        pred.submit(host[i % n_rot])
    pred.flush()
    barrier()
    t0 = time.perf_counter()
    got = 0
    for i in range(args.steps):
        out_host = pred.submit(host[i % n_rot])
        got += out_host is not None
    out_host = pred.flush()
    got += out_host is not None
    barrier()
    e2e_elapsed = time.perf_counter() - t0
    assert got == args.steps and bool(torch.isfinite(out_host).all())

    t = torch.tensor([elapsed, e2e_elapsed, elapsed_long], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed, e2e_elapsed, elapsed_long = float(t[0]), float(t[1]), float(t[2])

    out = None
    if rank == 0:
        pk = peaks()
        n = (H // 8) * (W // 8)
        per_launch = lambda k: (ms[k] / max(cnt[k], 1)) * 1e-3          # noqa: E731  seconds per bracket
        pv_flops = BATCH * 2.0 * n * n * 2048                           # a4: 2*N^2*dv per image
        core_flops = BATCH * 5120.0 * n * n                             # a2+a4: 2*N^2*(dk+dv) per image
        t_pv = per_launch(3)
        t_core = per_launch(0) + per_launch(1) + per_launch(2) + per_launch(3)
        head_bytes = BATCH * 324.0 * H * W                              # prob -> depth: (K+1)*HW*4 per image
        t_tail = per_launch(4)                                          # in-step: fused upsample + decode_ord + head on stride-8 logits
        from acan_b200 import ops
        prob = torch.rand(BATCH, K, H, W, device=device)
        t_head = ev_time(lambda: ops.head_inference(prob, "OR", "soft", DMIN, DMAX, K), iters=10)
        del prob
        cl = torch.channels_last
        fcl = torch.relu(torch.randn(BATCH, 512, H // 8, W // 8, device=device)).half().contiguous(memory_format=cl)
        vcl = torch.relu(torch.randn(BATCH, 2048, H // 8, W // 8, device=device)).half().contiguous(memory_format=cl)
        ws_core = ops.attention_forward_cl(fcl, vcl)["ws"]
        t_core_call = ev_time(lambda: ops.attention_forward_cl(fcl, vcl, ws=ws_core), iters=20)
        del fcl, vcl, ws_core
        xdec = torch.relu(torch.randn(BATCH, 2048, H // 8, W // 8, device=device)).half().contiguous(memory_format=cl)
        with torch.no_grad():
            t_cam = ev_time(lambda: net.decoder(xdec), iters=10)
        del xdec
        # traffic: dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this shape from the ncu --set full capture
        # profiles/r01b_attn_cl_4launch_pdl.ncu-rep (157.2 MB + 65.0 MB; algorithmic V 92 + P 63 MB read, O 92 MB fp16 written)
        roof = {"bound": "tensor", "kernel": "gemm_kernel<MODE_PVT, CG=2> (aggregation O = P V on tcgen05, CTA pairs, V read in place MN-major)",
                "achieved": pv_flops / t_pv / 1e12,
                "peak": pk["tensor"], "unit": "TFLOP/s", "frac": pv_flops / t_pv / 1e12 / pk["tensor"], "traffic": 222.2e6,
                "peak_source": pk["src"] + " (bf16 burst; the kernel is timed inside the step, so the sustained figure is given beside it)",
                "frac_vs_sustained": (pv_flops / t_pv / 1e12 / pk["tensor_sustained"]) if pk.get("tensor_sustained") else None,
                "launch_us": t_pv * 1e6, "launches_timed": cnt[3]}
        extra = {
            "attention_core": {"bound": "tensor", "what": "whole CAM core call (diagonal + fallback statistics + probability + aggregation passes), algorithmic "
                                       "5120*N^2 FLOP/img, CUDA events around 20 stand-alone calls at the step's shape; `us_in_step` = the per-pass event "
                                       "brackets inside the timed steps (events between the passes disable their launch overlap)",
                               "achieved": core_flops / t_core_call / 1e12, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": core_flops / t_core_call / 1e12 / pk["tensor"],
                               "call_us": t_core_call * 1e6,
                               "us_in_step": (ms_c[7] / max(cnt_c[7], 1)) * 1e3, "frac_in_step": core_flops / ((ms_c[7] / max(cnt_c[7], 1)) * 1e-3) / 1e12 / pk["tensor"],
                               "in_step_what": "ONE CUDA-event bracket around the whole acan_attn_fwd_nhwc call inside each of the K eagerly launched steps",
                               "frac_in_step_sum_of_pass_brackets": core_flops / t_core / 1e12 / pk["tensor"],
                               "us_in_step_per_pass": {"pack": per_launch(0) * 1e6, "stats": per_launch(1) * 1e6, "probs": per_launch(2) * 1e6, "pv": t_pv * 1e6}},
            "head": {"bound": "hbm", "what": "OR soft inference prob -> depth kernel, 324*HW B/img, measured stand-alone at the same shape", "achieved": head_bytes / t_head / 1e9,
                     "peak": pk["hbm"], "unit": "GB/s", "frac": head_bytes / t_head / 1e9 / pk["hbm"], "launch_us": t_head * 1e6},
            "tail": {"what": "fused classifier tail in the step: x8 bilinear + decode_ord + OR soft head from stride-8 logits (MUFU / shared-memory bound; "
                             "replaces upsample 1.9 ms + layout copy 1.0 ms + decode_ord 0.22 ms + head 0.07 ms of the unfused step)", "launch_us": t_tail * 1e6},
            "pool_branch_us": per_launch(5) * 1e6,
            "cam_module_us": t_cam * 1e6,
        }
        out = {
            "metric": "ACAN NYU-shape forward images/sec", "value": world * BATCH * args.steps / elapsed, "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
            "timed_1s": {"steps": long_steps, "seconds": elapsed_long, "ms_per_step": elapsed_long / long_steps * 1e3, "value": world * BATCH * long_steps / elapsed_long},
            "config": {"workload": WORKLOAD, "parallelism": f"dp{world} (batch-sharded replicas, no collective in inference)",
                       "precision": "fp16 channels_last cuDNN convs with folded BN and fused bias/ReLU/residual epilogues (acan_b200.inference.optimize_for_inference); attention core fp16 operands / fp32 accumulate, F and V consumed in place; fused upsample+decode+head tail in fp32",
                       "l2": f"{n_rot} rotating input batches (138 MB) + ~190 MB of fp16 weights per step exceed the 126 MB L2; no explicit flush",
                       "launch": launch_note},
            "clocks": clocks,
            "e2e": {"value": world * BATCH * args.steps / e2e_elapsed, "unit": "images/s", "h2d_bytes_per_step": BATCH * 3 * H * W * 4,
                    "d2h_bytes_per_step": BATCH * H * W * 4, "ms_per_step": e2e_elapsed / args.steps * 1e3},
            "gpu_launches": int(launches), "roofline": roof, **extra,
        }
    sd_cpu = net.source_state
    x_par = host[0][:2].contiguous().float()
    del pred, net, dev_in
    free_cuda()

    # ---- configs[3] / [4]: the training step as a sub-record (all ranks take part: it contains the gradient all-reduce)
    train = None
    if not args.no_train:
        # N > 1: the sub-record contains collectives.  A hang there (a rank that died, a communicator that never formed) must not cost the
        # headline, which has none: past the deadline rank 0 prints the line it already has and every rank leaves.
        guard, done = None, threading.Event()
        if world > 1:
            def expire():
                if done.is_set():
                    return
                try:
                    faulthandler.dump_traceback(file=sys.stderr, all_threads=True)      # where this rank is stuck
                except Exception:
                    pass
                if rank == 0:
                    out["train"] = {"error": f"the multi-GPU training sub-record did not finish within {TRAIN_DEADLINE_S} s at N = {world}; the headline (no collective in inference) is unaffected",
                                    "rank0_last_milestone": PROGRESS["phase"], "seconds_in_it": round(time.time() - PROGRESS["t"], 1)}
                    out["cpu_baseline"] = None
                    sys.stdout.write(json.dumps(out) + "\n")
                    sys.stdout.flush()
                os._exit(0)
            guard = threading.Timer(TRAIN_DEADLINE_S, expire)
            guard.daemon = True
            guard.start()
            faulthandler.dump_traceback_later(TRAIN_DEADLINE_S + 45, exit=True)       # C-level backstop: fires even if a blocked call holds the GIL
        try:
            train = run_train(args, rank, world, device, lib)
        except Exception as e:
            if world == 1:
                raise
            train = {"error": f"{type(e).__name__}: {str(e)[:300]}"}
            print(f"[bench rank {rank}] training sub-record failed: {train['error']}", file=sys.stderr, flush=True)
        done.set()
        if guard is not None:
            guard.cancel()
            faulthandler.cancel_dump_traceback_later()
        if train is not None and "error" in train and world > 1:       # the other ranks may be stuck in a collective: no orderly shutdown
            if rank == 0:
                out["train"], out["cpu_baseline"] = train, None
                sys.stdout.write(json.dumps(out) + "\n")
                sys.stdout.flush()
            os._exit(0)
        free_cuda()
    if rank == 0:
        out["train"] = train
        if world == 1 and not args.no_extras:
            out["gpu_reference"] = gpu_reference_extras(device, sd_cpu, out["ms_per_step"], train["ms_per_step"] if train else None, out["cam_module_us"])
            free_cuda()
            out["c3_kitti_hard"] = c3_line(device, args.steps)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            torch.set_num_threads(threads)
            fn, kind = cpu_reference_callable(sd_cpu)
            want = fn(x_par)                                  # the reference's depth for the first two images of the timed batch
            gotd = depth_first[:2].float().cpu()
            e_max = float((gotd - want).abs().max() / want.abs().max())
            e_mean = float(((gotd - want).abs() / want.abs()).mean())
            out["parity"] = {"depth_rel_err_max_norm": e_max, "depth_rel_err_mean": e_mean, "tolerance": {"max_norm": PARITY_TOL_MAX, "mean": PARITY_TOL_MEAN},
                             "config": f"configs[1] network as benched (ResNet-101, 256x352, fp16 folded-BN), images 0-1 of timed batch 0, vs the reference's CPU path ({kind}) on the same weights",
                             "note": "random-init weights: an untrained 101-layer network amplifies the 2^-11 roundings of 16-bit activations; the fp32 arm below holds 1e-3"}
            out["fp32_parity_arm"] = fp32_parity_arm(device, sd_cpu, x_par, want, args.steps)
            times = time_cpu(fn, REF_SAMPLE_BATCH, 6, 1)
            out["cpu_baseline"] = {"value": REF_SAMPLE_BATCH * len(times) / sum(times), "unit": "images/s", "cores": threads, "kind": kind,
                                   "sample": f"bounded sample: {REF_SAMPLE_BATCH} of the step's {BATCH} images per step x {len(times)} steps of the same network ({'unmodified reference modules' if kind == 'reference' else 'oracle port'}, fp32 torch CPU ops), {sum(times):.1f} s"}
            print(json.dumps(out))
            assert e_max <= PARITY_TOL_MAX and e_mean <= PARITY_TOL_MEAN, f"parity of the benched network vs the reference failed: {out['parity']}"
            assert out["fp32_parity_arm"]["depth_rel_err_max_norm"] <= 1e-3, f"fp32 parity arm failed: {out['fp32_parity_arm']}"
        else:
            out["cpu_baseline"] = None
            print(json.dumps(out))
    if world > 1:
        sys.stdout.flush()
        bye = threading.Timer(30.0, lambda: os._exit(0))      # the line is out: a communicator teardown that hangs must not turn into a timeout
        bye.daemon = True
        bye.start()
        dist.destroy_process_group()
        bye.cancel()


if __name__ == "__main__":
    main()
