#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: ACAN NYU-shape images/sec on N B200s, CAM attention TFLOP/s vs peak.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    (N > 1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...)

Workload = BASELINE.json configs[1]: full ACAN forward (dilated ResNet-101 backbone + context aggregation module +
ordinal soft inference), NYU-shaped 256x352 synthetic batch of 16 images per GPU (weak scaling, no data-path
collective: every image is an independent unit).  One step = net(images) -> net.inference(y), the
reference's eval call pattern (depthest_trainer.py:184-185), with the drop-in modules of acan_b200/.

  value    images/s with the batch already resident in HBM (CUDA events, max over ranks).
  e2e      the same call with HOST buffers: pinned images -> H2D, forward, inference, depth -> D2H, every step.
  roofline the dominant hand-written kernel (aggregation pass O^T = V P^T on tcgen05), its CUDA-event time measured
           live in this run via the library's event hooks, against MEASURED_PEAKS.json; `attention_core` and `head`
           give the same arithmetic for the whole fused CAM core and the depth head.
  cpu_baseline  the oracle port of the reference forward on the host cores, bounded sample (rank 0, N = 1).
--impl reference times that CPU path alone (the reference is PyTorch-on-CPU code; the oracle is its restatement).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

H, W, BATCH = 256, 352, 16
DMIN, DMAX, K = 0.72, 10.0, 80
LAYERS101 = [3, 4, 23, 3]
AMP = torch.float16      # 16-bit autocast for the cuDNN convolutions; fp16 lets the CAM consume F and V in place
WORKLOAD = "configs[1]: full ACAN fwd (ResNet-101 + CAM + OR soft inference), NYU 256x352, batch 16 per GPU"


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


def build_net(device):
    from acan_b200.network import ResNet
    torch.manual_seed(2019)
    net = ResNet(DMIN, DMAX, K, "OR", "soft", "attention", H, W, alpha=0, beta=0, layers=LAYERS101)
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
        net(torch.rand(BATCH, 3, H, W, device=device).contiguous(memory_format=torch.channels_last))
    for m in bns:
        m.momentum = 0.1
    from acan_b200.inference import optimize_for_inference
    net.eval()
    fast = optimize_for_inference(net)               # BN folded into the convolutions, fused conv epilogues, fp16 channels_last
    fast.source_state = {k: v.detach().float().cpu() for k, v in net.state_dict().items()}   # for the CPU baseline (same weights)
    return fast


def step_device(net, images):
    return net.inference(net(images))                # the reference's eval call pattern (depthest_trainer.py:184-185)


def cpu_reference_forward(sd_cpu, n_iter, warm, threads):
    """oracle port of the reference forward (ResNet-101 + CAM + decode_ord + soft OR) on the host cores, B = 1."""
    from oracle import acan_oracle as O
    torch.set_num_threads(threads)
    x = torch.rand(1, 3, H, W)
    times = []
    with torch.no_grad():
        for i in range(warm + n_iter):
            t0 = time.perf_counter()
            out = O.acan_forward(x, sd_cpu, LAYERS101, "OR", training=False)
            depth = O.inference(out["y"], "OR", "soft", DMIN, DMAX, K)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    return times, depth


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (oracle port), all host threads."""
    if rank != 0:
        return
    from acan_b200.network import ResNet
    torch.manual_seed(2019)
    net = ResNet(DMIN, DMAX, K, "OR", "soft", "attention", H, W, alpha=0, beta=0, layers=LAYERS101)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    threads = os.cpu_count() or 1
    times, _ = cpu_reference_forward(sd, args.steps, args.warmup, threads)
    tot = sum(times)
    val = len(times) / tot
    sample = f"B=1 image per step x {len(times)} steps, fp32, torch CPU ops, eval-mode BN"
    print(json.dumps({
        "impl": "reference", "metric": "ACAN NYU-shape forward images/sec", "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / len(times) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


TRAIN_BATCH = 8
TRAIN_WORKLOAD = "configs[3]: ACAN training step with attention loss (beta=1, alpha=0), fwd + loss + bwd + SGD, NYU 256x352, batch 8 per GPU"


def run_train(args, rank, world, device, lib):
    """--workload train: BASELINE.json configs[3] (and configs[4] under torchrun): ResNet-101 + CAM + fused ordinal tail loss +
    fused attention loss, forward + backward + SGD(momentum) step; N > 1 averages gradients with the bucketed NCCL all-reduce
    of acan_b200.parallel (overlapped with backward).  bf16 autocast, channels_last; the hand-written kernels run in fp16
    operands / fp32 accumulate (attention) and fp32 (losses)."""
    import torch.distributed as dist
    from acan_b200 import ops
    from acan_b200.modules import AttentionLoss2d
    from acan_b200.network import OrdinalRegression2d, ResNet
    from acan_b200.parallel import GradAllReducer
    B = TRAIN_BATCH
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

    from acan_b200.training import GraphedTrainStep
    graph_note = None
    try:
        step = GraphedTrainStep(net, crit, opt, dev_x[0], dev_y[0], amp_dtype=torch.bfloat16, reducer=reducer if world > 1 else None,
                                graph=not args.no_graph)
    except Exception as e:          # same kernels either way: a failed capture must not cost the measurement
        torch.cuda.synchronize()
        graph_note = f"eager (CUDA graph capture failed: {type(e).__name__}: {str(e)[:120]})"
        reducer.enabled = True
        step = GraphedTrainStep(net, crit, opt, dev_x[0], dev_y[0], amp_dtype=torch.bfloat16, reducer=reducer if world > 1 else None, graph=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(dev_x[i % n_rot], dev_y[i % n_rot])
    barrier()
    launches0 = lib.acan_launch_count()
    sampler = ClockSampler(device.index)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(dev_x[i % n_rot], dev_y[i % n_rot])
    e1.record()
    barrier()
    clocks = sampler.stop()
    elapsed = e0.elapsed_time(e1) * 1e-3
    launches = lib.acan_launch_count() - launches0
    if step.graph is not None:                       # replays do not pass through the library's host-side launch counter
        launches = step.library_launches_per_step * args.steps
    assert bool(torch.isfinite(loss)), "training loss is not finite"
    # end to end: pinned host images + labels in, the loss value out, every step
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        loss_host = float(step(host_x[i % n_rot], host_y[i % n_rot]))     # pinned host -> the step's static device buffers, loss -> host
    barrier()
    e2e_elapsed = time.perf_counter() - t0
    t = torch.tensor([elapsed, e2e_elapsed], device=device, dtype=torch.float64)
    in_sync = None
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # averaged gradients keep the replicas' weights identical: compare a checksum of all parameters across ranks
        chk = torch.stack([p.detach().double().sum() for p in net.parameters()]).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        in_sync = bool((hi - lo).abs() <= 1e-9 * hi.abs())
    elapsed, e2e_elapsed = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # roofline of the dominant hand-written group of the step: attention forward (+ fused loss) and backward, stand-alone, same shape
    pk = peaks()
    n = (H // 8) * (W // 8)
    f = torch.relu(torch.randn(B, 512, H // 8, W // 8, device=device))
    v = torch.relu(torch.randn(B, 2048, H // 8, W // 8, device=device))
    lab = torch.randint(0, K, (B, 1, H, W), device=device).float()
    go = torch.randn(B, 2048, H // 8, W // 8, device=device) * 1e-3
    o = ops.attention_forward(f, v, dis_label=lab)

    def ev(fn_, iters=10):
        for _ in range(3):
            fn_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn_()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e-3 / iters
    t_f = ev(lambda: ops.attention_forward(f, v, dis_label=lab, ws=o["ws"]))
    t_b = ev(lambda: ops.attention_backward(f, o["row_stats"], v, o["ctx"], go, lab, 1.0))
    bwd_flops = B * 10240.0 * n * n
    del f, v, lab, go, o
    # the dominant hand-written kernels of the step by time are the BN (+ReLU/+residual) passes (105 layers, ~3.3 ms of the step): the
    # backward kernel at the most common large layer shape (layer3's 1024-channel maps), operands rotated through > 2x L2, replayed from
    # a CUDA graph so that the number is device time, not host launch rate
    bc, bh, bw = 1024, H // 8, W // 8
    n_rot = 14
    xs = [torch.randn(B, bc, bh, bw, device=device, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last) for _ in range(n_rot)]
    gs = [torch.randn(B, bc, bh, bw, device=device, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last) for _ in range(n_rot)]
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
    out = {
        "metric": "ACAN NYU-shape training images/sec", "value": world * B * args.steps / elapsed, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": TRAIN_WORKLOAD, "parallelism": (f"dp{world} (batch-sharded replicas; NCCL gradient all-reduce " +
                                   ("bucketed, overlapped with backward from grad hooks)" if args.no_graph else "one in-place all-reduce of a flat fp32 gradient buffer between the backward graph and the optimizer graph)")) if world > 1 else "dp1",
                   "precision": "bf16 autocast channels_last cuDNN convolutions; fused BN/ReLU kernels bf16 in / fp32 statistics; attention fwd/bwd fp16 operands / fp32 accumulate; losses fp32; SGD momentum on fp32 weights",
                   "launch": graph_note or ("eager" if args.no_graph else "whole step (fwd + losses + bwd + SGD) replayed from a CUDA graph; N > 1: backward graph, one NCCL all-reduce, optimizer graph (acan_b200.training.GraphedTrainStep)"),
                   "l2": f"{n_rot} rotating input batches; activations (> 10 GB per step) exceed the 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": world * B * args.steps / e2e_elapsed, "unit": "images/s", "h2d_bytes_per_step": B * 4 * H * W * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_elapsed / args.steps * 1e3},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "bn_bwd_coop_kernel<bf16> (BN + ReLU backward: one cooperative launch, 14 B per element algorithmic: dy, y, x read for the "
                               "statistics, read again -- from L2 -- for dx, dx written), layer3 shape [8,1024,32,44], stand-alone, graph-replayed over 14 rotating operand sets",
                     "achieved": 14.0 * bn_elems / t_bn_b / 1e9, "peak": pk["hbm"], "unit": "GB/s", "frac": 14.0 * bn_elems / t_bn_b / 1e9 / pk["hbm"],
                     "traffic": 76.1e6, "peak_source": pk["src"] + " (copy)", "launch_us": t_bn_b * 1e6,
                     "note": "traffic = dram read 71.5 MB + write 4.6 MB from profiles/r01b_train_kernels_bn_tailloss.ncu-rep (the 23 MB of dx are still in L2 when the kernel ends)"},
        "bn_forward": {"bound": "hbm", "what": "bn_fwd_coop_kernel<bf16>, same shape, 6 B per element algorithmic", "achieved": 6.0 * bn_elems / t_bn_f / 1e9, "peak": pk["hbm"],
                       "unit": "GB/s", "frac": 6.0 * bn_elems / t_bn_f / 1e9 / pk["hbm"], "launch_us": t_bn_f * 1e6},
        "attention_bwd": {"bound": "tensor", "what": "acan_attn_bwd (P recompute, dV, dP -> dS with fused attention-loss term, dF: five tcgen05 GEMMs + packs), fp32-NCHW ABI, "
                                                     "stand-alone at the step's shape, algorithmic 10240*N^2 FLOP/img",
                          "achieved": bwd_flops / t_b / 1e12, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": bwd_flops / t_b / 1e12 / pk["tensor"], "launch_us": t_b * 1e6},
        "attention_fwd_loss_us": t_f * 1e6, "loss": loss_host, "replicas_in_sync": in_sync, "cpu_baseline": None,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="train workload: launch the step eagerly instead of replaying its CUDA graph")
    ap.add_argument("--workload", default="infer", choices=["infer", "train"],
                    help="infer (default): BASELINE.json configs[1], the headline.  train: configs[3]/[4], the beta=1 training step, B=8 per GPU")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

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
    init_distributed("nccl" if world > 1 else None)
    lib = _lib.load()
    _lib.check(lib.acan_device_check(), "acan_device_check")
    torch.backends.cudnn.benchmark = True                 # as depthest_main.py:65
    if args.workload == "train":
        run_train(args, rank, world, device, lib)
        return

    net = build_net(device)
    n_rot = 8                                             # 8 distinct input batches = 138 MB > 126 MB L2
    host = [torch.rand(BATCH, 3, H, W).contiguous(memory_format=torch.channels_last).pin_memory() for _ in range(n_rot)]
    dev_in = [h.to(device, non_blocking=True) for h in host]
    depth_host = torch.empty(BATCH, 1, H, W).pin_memory()
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
    clocks = sampler.stop()
    elapsed = e0.elapsed_time(e1) * 1e-3
    launches = lib.acan_launch_count() - launches0 if pred is None else pred.library_launches_per_step * args.steps

    # ---- per-kernel CUDA-event brackets (roofline): the same K steps on the same inputs, launched eagerly with the library's
    #      event hooks on (events cannot be timed from inside a replayed graph); not part of `value`
    for i in range(2):
        step_device(net, dev_in[i % n_rot])
    barrier()
    _lib.check(lib.acan_profile_begin(args.steps * 16 + 64), "acan_profile_begin")
    for i in range(args.steps):
        step_device(net, dev_in[i % n_rot])
    barrier()
    ms = (ctypes.c_float * 7)()
    cnt = (ctypes.c_int * 7)()
    _lib.check(lib.acan_profile_end(ms, cnt), "acan_profile_end")

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

    t = torch.tensor([elapsed, e2e_elapsed], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed, e2e_elapsed = float(t[0]), float(t[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    n = (H // 8) * (W // 8)
    per_launch = lambda k: (ms[k] / max(cnt[k], 1)) * 1e-3          # noqa: E731  seconds per bracket
    pv_flops = BATCH * 2.0 * n * n * 2048                           # a4: 2*N^2*dv per image
    core_flops = BATCH * 5120.0 * n * n                             # a2+a4: 2*N^2*(dk+dv) per image
    t_pv = per_launch(3)
    t_core = per_launch(0) + per_launch(1) + per_launch(2) + per_launch(3)
    head_bytes = BATCH * 324.0 * H * W                              # prob -> depth: (K+1)*HW*4 per image
    t_tail = per_launch(4)                                          # in-step: fused upsample + decode_ord + head on stride-8 logits
    # the HBM-bound head kernel (prob [B,K,H,W] -> depth) is no longer on the eval path (the fused tail reads 0.9 MB / image);
    # its roofline is measured here on its own, same shape, CUDA events, 3 warm-up + 10 timed launches
    from acan_b200 import ops
    prob = torch.rand(BATCH, K, H, W, device=device)
    for _ in range(3):
        ops.head_inference(prob, "OR", "soft", DMIN, DMAX, K)
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for _ in range(10):
        ops.head_inference(prob, "OR", "soft", DMIN, DMAX, K)
    h1.record()
    torch.cuda.synchronize()
    t_head = h0.elapsed_time(h1) * 1e-4
    del prob
    # the whole attention core as the step runs it (one acan_attn_fwd_nhwc_f16 call: diagonal, fallback statistics, probability
    # and aggregation passes chained by programmatic dependent launch), stand-alone at the same shape, no event hooks in between
    # (the in-step per-pass brackets below record events between the passes, which serialises them)
    fcl = torch.relu(torch.randn(BATCH, 512, H // 8, W // 8, device=device)).half().contiguous(memory_format=torch.channels_last)
    vcl = torch.relu(torch.randn(BATCH, 2048, H // 8, W // 8, device=device)).half().contiguous(memory_format=torch.channels_last)
    ws_core = ops.attention_forward_cl(fcl, vcl)["ws"]
    for _ in range(3):
        ops.attention_forward_cl(fcl, vcl, ws=ws_core)
    h0.record()
    for _ in range(20):
        ops.attention_forward_cl(fcl, vcl, ws=ws_core)
    h1.record()
    torch.cuda.synchronize()
    t_core_call = h0.elapsed_time(h1) * 1e-3 / 20
    del fcl, vcl, ws_core
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
                           "call_us": t_core_call * 1e6, "frac_in_step": core_flops / t_core / 1e12 / pk["tensor"],
                           "us_in_step": {"pack": per_launch(0) * 1e6, "stats": per_launch(1) * 1e6, "probs": per_launch(2) * 1e6, "pv": t_pv * 1e6}},
        "head": {"bound": "hbm", "what": "OR soft inference prob -> depth kernel, 324*HW B/img, measured stand-alone at the same shape", "achieved": head_bytes / t_head / 1e9 if t_head > 0 else None,
                 "peak": pk["hbm"], "unit": "GB/s", "frac": head_bytes / t_head / 1e9 / pk["hbm"] if t_head > 0 else None, "launch_us": t_head * 1e6},
        "tail": {"what": "fused classifier tail in the step: x8 bilinear + decode_ord + OR soft head from stride-8 logits (expf/div-bound; "
                         "replaces upsample 1.9 ms + layout copy 1.0 ms + decode_ord 0.22 ms + head 0.07 ms of the unfused step)", "launch_us": t_tail * 1e6},
        "pool_branch_us": per_launch(5) * 1e6,
    }
    out = {
        "metric": "ACAN NYU-shape forward images/sec", "value": world * BATCH * args.steps / elapsed, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "parallelism": f"dp{world} (batch-sharded replicas, no collective in inference)",
                   "precision": "fp16 channels_last cuDNN convs with folded BN and fused bias/ReLU/residual epilogues (acan_b200.inference.optimize_for_inference); attention core fp16 operands / fp32 accumulate, F and V consumed in place; fused upsample+decode+head tail in fp32",
                   "l2": f"{n_rot} rotating input batches (138 MB) + ~190 MB of fp16 weights per step exceed the 126 MB L2; no explicit flush",
                   "launch": launch_note},
        "clocks": clocks,
        "e2e": {"value": world * BATCH * args.steps / e2e_elapsed, "unit": "images/s", "h2d_bytes_per_step": BATCH * 3 * H * W * 4,
                "d2h_bytes_per_step": BATCH * H * W * 4, "ms_per_step": e2e_elapsed / args.steps * 1e3},
        "gpu_launches": int(launches), "roofline": roof, **extra,
    }

    if world == 1 and not args.no_cpu_baseline:
        sd = net.source_state
        threads = os.cpu_count() or 1
        times, d_cpu = cpu_reference_forward(sd, 60, 2, threads)
        out["cpu_baseline"] = {"value": len(times) / sum(times), "unit": "images/s", "cores": threads, "kind": "port",
                               "sample": f"B=1 image x {len(times)} forwards of the same network (oracle port, fp32 torch CPU ops), {sum(times):.1f} s"}
    else:
        out["cpu_baseline"] = None
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
