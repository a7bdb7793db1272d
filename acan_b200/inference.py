"""Inference-time graph optimisation of the host network around the hot path (library calls only).

``optimize_for_inference(net)`` returns a frozen fp16 / channels_last copy of an ``acan_b200.network.ResNet`` in which
  * every eval-mode BatchNorm is folded into the convolution in front of it (an exact algebraic rewrite:
    W' = W * g / sqrt(var + eps),  b' = beta + (b - mean) * g / sqrt(var + eps)),
  * conv + bias + ReLU and conv + bias + residual-add + ReLU are single cuDNN fused-epilogue calls
    (``torch.cudnn_convolution_relu`` / ``torch.cudnn_convolution_add_relu``),
so the only kernels left between the convolutions are the hand-written ones of the context aggregation module and
the fused classifier tail.  In the profile of the unfused network the BN / ReLU / add kernels were 40-60 % of the step
(``profiles/``); none of this touches the reference's semantics beyond fp16 rounding, and the result is checked against
the unfused network in ``tests/test_gpu_parity.py``.  The returned module has the reference's ``forward`` /
``inference`` / ``predict_depth`` surface but NOT its ``state_dict`` layout -- it is an inference artefact, not a checkpoint.
"""
from __future__ import annotations

import copy
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .network import Bottleneck, ResNet
from .training import _capture_mode


def fold_conv_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d):
    """(weight, bias) of the convolution equal to bn(conv(x)) in eval mode, in fp32."""
    w = conv.weight.detach().float()
    b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return w * scale.view(-1, 1, 1, 1), bn.bias.detach().float() + (b - bn.running_mean.detach().float()) * scale


class FusedConv(nn.Module):
    """conv + bias (+ residual) (+ ReLU) as one cuDNN call on channels_last tensors (fp16 by default; fp32 for the parity arm)."""

    def __init__(self, weight: torch.Tensor, bias: torch.Tensor, stride, padding, dilation, relu: bool, dtype: torch.dtype = torch.float16):
        super().__init__()
        # fp32 stays NCHW: on this stack (torch 2.11 / cuDNN 9) an fp32 3x3 convolution with channels_last weights and TF32 disabled
        # returns garbage (deviation 3); channels_last is used in half precision only
        fmt = torch.channels_last if dtype == torch.float16 else torch.contiguous_format
        self.register_buffer("weight", weight.to(dtype).contiguous(memory_format=fmt))
        self.register_buffer("bias", bias.to(dtype).contiguous())
        self.stride, self.padding, self.dilation, self.relu = tuple(stride), tuple(padding), tuple(dilation), relu
        self._fused_ok = True

    def forward(self, x: torch.Tensor, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self._fused_ok and self.relu:
            try:
                if residual is None:
                    return torch.cudnn_convolution_relu(x, self.weight, self.bias, self.stride, self.padding, self.dilation, 1)
                return torch.cudnn_convolution_add_relu(x, self.weight, residual, 1.0, self.bias, self.stride, self.padding, self.dilation, 1)
            except RuntimeError:
                self._fused_ok = False          # configuration the fused-epilogue path does not support: plain calls from now on
        y = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation)       # bias None: moved into the consumer (FusedBottleneck)
        if residual is not None:
            y = y + residual
        return torch.relu_(y) if self.relu else y


class FusedBottleneck(nn.Module):
    def __init__(self, blk: Bottleneck, dtype: torch.dtype = torch.float16):
        super().__init__()
        def fc(conv, bn, relu):
            w, b = fold_conv_bn(conv, bn)
            return FusedConv(w, b, conv.stride, conv.padding, conv.dilation, relu, dtype)
        self.c1, self.c2 = fc(blk.conv1, blk.bn1, True), fc(blk.conv2, blk.bn2, True)
        w3, b3 = fold_conv_bn(blk.conv3, blk.bn3)
        self.project = None
        if blk.downsample is not None:
            # out = relu(c3(.) + b3 + project(x) + bp): the projection's folded bias moves into c3's (added in fp32), the projection
            # itself runs bias-free -- a convolution without a fused activation would otherwise pay a separate broadcast-add pass
            wp, bp = fold_conv_bn(blk.downsample[0], blk.downsample[1])
            b3 = b3 + bp
            self.project = FusedConv(wp, bp, blk.downsample[0].stride, blk.downsample[0].padding, blk.downsample[0].dilation, False, dtype)
            self.project.bias = None
        self.c3 = FusedConv(w3, b3, blk.conv3.stride, blk.conv3.padding, blk.conv3.dilation, True, dtype)

    def forward(self, x):
        skip = x if self.project is None else self.project(x)
        return self.c3(self.c2(self.c1(x)), skip)


class InferenceNet(nn.Module):
    """frozen channels_last network: fused backbone -> SADecoder -> fused tail.  dtype fp16 (default): the benched configuration, the CAM on
    its half-precision in-place kernel path.  dtype fp32: the PARITY arm -- the same folded convolutions in strict fp32 (run it with TF32
    off), the CAM through the drop-in module's generic path (fp16 tensor-core operands, fp32 in / out): <= 1e-3 on depth against the
    reference at the benched configuration, where 101 layers of fp16 activations cannot be."""

    def __init__(self, net: ResNet, dtype: torch.dtype = torch.float16):
        super().__init__()
        if net.use_inter:
            raise NotImplementedError("optimize_for_inference: the auxiliary (alpha != 0) branch is a training-time output")
        self.dtype = dtype
        w, b = fold_conv_bn(net.conv1, net.bn1)
        # (the 3-channel stem runs on an sm_80 implicit-GEMM kernel, 88 us at configs[1]; zero-padding the input to 4 channels makes the
        #  convolution 64 us but costs 35 us of padding, to 8 channels 323 us: measured with scripts/stem_bench.py, left as it is)
        self.stem = FusedConv(w, b, net.conv1.stride, net.conv1.padding, net.conv1.dilation, True, dtype)
        self.maxpool = net.maxpool
        self.blocks = nn.Sequential(*[FusedBottleneck(blk, dtype) for layer in (net.layer1, net.layer2, net.layer3, net.layer4) for blk in layer])
        self.decoder = copy.deepcopy(net.decoder)
        if dtype != torch.float16:
            self.classifier_conv = copy.deepcopy(net.classifier.conv)
            self.min_depth, self.max_depth, self.num_classes = net.min_depth, net.max_depth, net.num_classes
            self.classifierType, self.inferenceType = net.classifierType, net.inferenceType
            self.eval()
            for p in self.parameters():
                p.requires_grad_(False)
            return
        key = self.decoder.saconv.f_key                 # F = ReLU(BN(conv(x))) -> one fused call as well
        wk, bk = fold_conv_bn(key.conv, key.bn)
        self.decoder.saconv.fused_key = FusedConv(wk, bk, key.conv.stride, key.conv.padding, key.conv.dilation, True)
        val = self.decoder.saconv.f_value
        self.decoder.saconv.fused_value = nn.Sequential(
            FusedConv(val.conv1.weight.detach(), val.conv1.bias.detach(), val.conv1.stride, val.conv1.padding, val.conv1.dilation, True),
            FusedConv(val.conv2.weight.detach(), val.conv2.bias.detach(), val.conv2.stride, val.conv2.padding, val.conv2.dilation, True))
        self.classifier_conv = copy.deepcopy(net.classifier.conv).half().to(memory_format=torch.channels_last)
        self.min_depth, self.max_depth, self.num_classes = net.min_depth, net.max_depth, net.num_classes
        self.classifierType, self.inferenceType = net.classifierType, net.inferenceType
        self.eval()
        for p in self.parameters():
            p.requires_grad_(False)

    def _features(self, x):
        x = x.to(self.dtype).contiguous(memory_format=torch.channels_last if self.dtype == torch.float16 else torch.contiguous_format)
        x = self.stem(x)
        mp = self.maxpool
        if self.dtype == torch.float16 and (mp.kernel_size, mp.stride, mp.padding, mp.dilation, mp.ceil_mode) == (3, 2, 1, 1, False) and x.shape[1] % 8 == 0:
            from . import ops
            x = ops.maxpool3x3s2_cl(x)                      # 3x3 / 2 max pooling of the channels_last stem output in one HBM pass
        else:
            x = mp(x)
        x = self.blocks(x)
        x, sim_map = self.decoder(x)
        return self._classify(x), sim_map

    def _classify(self, x):
        """classifier.conv (1x1, 2048 -> 2K or K channels) as a plain GEMM over the channels_last rows: for 160 output channels cuDNN
        falls back to an sm_80 implicit-GEMM kernel (248 us at configs[1], 5 % of the step); the cuBLASLt GEMM takes ~25 us.  The
        result is viewed back as [B,C,h,w] in channels_last memory, which is the layout the fused tail kernel reads."""
        conv = self.classifier_conv
        b, c, h, w = x.shape
        rows = x.permute(0, 2, 3, 1)
        if not rows.is_contiguous() or conv.kernel_size != (1, 1) or conv.stride != (1, 1) or conv.padding != (0, 0):
            return conv(x)
        z = F.linear(rows.reshape(b * h * w, c), conv.weight.view(conv.out_channels, c), conv.bias)
        return z.view(b, h, w, conv.out_channels).permute(0, 3, 1, 2)

    @torch.no_grad()
    def forward(self, x):
        from .modules import LazyProb
        z, sim_map = self._features(x)
        y = LazyProb(z, "OR" if self.classifierType == "OR" else "CE", self.min_depth, self.max_depth, self.num_classes)
        return {"inter_y": None, "sim_map": sim_map, "y": y}

    @torch.no_grad()
    def inference(self, y):
        from . import ops
        from .modules import LazyProb
        if isinstance(y, list):
            y = y[-1]
        if isinstance(y, dict):
            y = y["y"]
        mode = "soft" if self.inferenceType == "soft" else "hard"
        if isinstance(y, LazyProb):
            return y.depth(mode)
        return ops.head_inference(y, "OR" if self.classifierType == "OR" else "CE", mode, self.min_depth, self.max_depth, self.num_classes)

    @torch.no_grad()
    def predict_depth(self, x):
        return self.inference(self.forward(x))


def optimize_for_inference(net: ResNet, dtype: torch.dtype = torch.float16) -> InferenceNet:
    """frozen, BN-folded, fused-epilogue copy of ``net`` (which must hold calibrated / trained BN statistics): fp16 (the benched
    configuration) or fp32 (the parity arm, see InferenceNet)."""
    if not isinstance(net, ResNet):
        raise TypeError("optimize_for_inference expects an acan_b200.network.ResNet")
    if dtype not in (torch.float16, torch.float32):
        raise ValueError("optimize_for_inference: dtype must be torch.float16 or torch.float32")
    return InferenceNet(net.eval(), dtype)


class StreamedPredictor:
    """Host-to-host inference with the copies off the critical path (the role ``DataPrefetcher`` plays in the reference's
    trainer, trainers/trainer.py:23-49): while batch i computes, batch i+1 is uploaded on a copy stream and the depth of
    batch i-1 is downloaded on another.  ``submit(host_images) -> host depth of the PREVIOUS submission (or None)``;
    ``flush()`` returns the last one.  Host tensors should be pinned; every step still moves its own input and output."""

    def __init__(self, net, device=None):
        self.net = net
        self.device = device or next(net.buffers()).device
        self.h2d, self.d2h = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self._pending = None          # (event, host_out) of the download in flight
        self._slots = [None, None]    # device input double buffer
        self._i = 0

    @torch.no_grad()
    def submit(self, host_images: torch.Tensor):
        cur = torch.cuda.current_stream(self.device)
        slot = self._i & 1
        self._i += 1
        with torch.cuda.stream(self.h2d):
            self.h2d.wait_stream(cur)                         # the slot's previous consumer (two steps ago) has been enqueued
            x = host_images.to(self.device, non_blocking=True)
            self._slots[slot] = x
            up = torch.cuda.Event()
            up.record(self.h2d)
        cur.wait_event(up)
        x.record_stream(cur)
        depth = self.net.inference(self.net(x))
        done = torch.cuda.Event()
        done.record(cur)
        prev = self._collect()
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(done)
            host_out = torch.empty(depth.shape, dtype=depth.dtype, pin_memory=True)
            host_out.copy_(depth, non_blocking=True)
            depth.record_stream(self.d2h)
            ev = torch.cuda.Event()
            ev.record(self.d2h)
        self._pending = (ev, host_out)
        return prev

    def _collect(self):
        if self._pending is None:
            return None
        ev, out = self._pending
        ev.synchronize()
        self._pending = None
        return out

    def flush(self):
        return self._collect()


class GraphedPredictor:
    """The eval call pattern ``net.inference(net(images))`` (depthest_trainer.py:184-185) captured once into CUDA graphs and
    replayed: ~250 launches per step (cuDNN convolutions + the kernels of this package) stop costing host time, and the host
    copies ride on their own streams.  Two slots (static device input + static depth output + one graph each) let the upload
    of batch i+1 and the download of depth i-1 overlap the replay of batch i.

        pred = GraphedPredictor(fast_net, example_images_on_device)
        depth = pred.predict(images_on_device)                       # device in, device out (valid until the next-but-one call)
        prev = pred.submit(pinned_host_images) ... pred.flush()      # host in, host out, as StreamedPredictor
    """

    def __init__(self, net, example: torch.Tensor, warmup: int = 3):
        if not example.is_cuda:
            raise RuntimeError("GraphedPredictor needs a CUDA example batch: acan_b200 has no CPU path")
        self.net, self.device = net, example.device
        self.h2d, self.d2h = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
        self.x = [example.clone(), example.clone()]
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):                  # cuDNN autotuning + workspace allocations happen outside the capture
                net.inference(net(self.x[0]))
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        from . import _lib
        self.graphs, self.depth = [], []
        n0 = _lib.load().acan_launch_count()
        for slot in range(2):
            g = torch.cuda.CUDAGraph()
            # one memory pool PER graph: with a shared pool the second capture may place its output in memory the first graph uses
            # for temporaries, and replaying slot 0 would then corrupt depth[1] while its download is still in flight
            with torch.cuda.graph(g, capture_error_mode=_capture_mode()), torch.no_grad():
                d = net.inference(net(self.x[slot]))
            self.graphs.append(g)
            self.depth.append(d)
        self.library_launches_per_step = int(_lib.load().acan_launch_count() - n0) // 2
        self.host_out = [torch.empty(self.depth[0].shape, dtype=self.depth[0].dtype, pin_memory=True) for _ in range(2)]
        self._done = [None, None]         # compute of the slot's last replay finished
        self._down = [None, None]         # download of the slot's last depth finished
        self._pending = None              # slot whose download has not been handed to the caller yet
        self._i = 0

    def predict(self, images: torch.Tensor) -> torch.Tensor:
        """device-resident images -> depth on the device (the static output of the slot: consume before the next-but-one call)."""
        slot = self._i & 1
        self._i += 1
        cur = torch.cuda.current_stream(self.device)
        if self._down[slot] is not None:
            cur.wait_event(self._down[slot])
        self.x[slot].copy_(images, non_blocking=True)
        self.graphs[slot].replay()
        return self.depth[slot]

    def submit(self, host_images: torch.Tensor):
        """pinned host images in; returns the host depth of the PREVIOUS submission (None the first time)."""
        slot = self._i & 1
        self._i += 1
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.h2d):
            if self._done[slot] is not None:
                self.h2d.wait_event(self._done[slot])         # the replay that last read this slot's input (two steps ago)
            self.x[slot].copy_(host_images, non_blocking=True)
            up = torch.cuda.Event()
            up.record(self.h2d)
        cur.wait_event(up)
        if self._down[slot] is not None:
            cur.wait_event(self._down[slot])                  # this slot's previous depth has left the device
        self.graphs[slot].replay()
        done = torch.cuda.Event()
        done.record(cur)
        self._done[slot] = done
        prev = self._collect()
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(done)
            self.host_out[slot].copy_(self.depth[slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.d2h)
        self._down[slot] = ev
        self._pending = slot
        return prev

    def _collect(self):
        if self._pending is None:
            return None
        slot, self._pending = self._pending, None
        self._down[slot].synchronize()
        return self.host_out[slot]

    def flush(self):
        return self._collect()
