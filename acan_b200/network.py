"""Host-side mirror of ``code/models/model.py`` around the B200 hot path.

``ResNet`` keeps the reference's constructor, child order (conv1, bn1, relu1, maxpool, layer1-4, decoder,
[interout,] classifier -- creator/params.py:17-19 slices ``children()[:8]`` / ``[8:]``), ``state_dict`` keys and the
``forward -> {'inter_y','sim_map','y'}`` / ``inference(y)`` / ``LossFunc.forward(preds,label,epoch)`` surface
(model.py:80-99, 216-272).  The dilated-ResNet encoder and the 1x1/3x3 convolutions are cuDNN library calls
(out of scope to hand-write); the context aggregation module, ``decode_ord``, the four
inference heads, ``continuous2discrete`` and the attention loss run the sm_100a kernels.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn as nn

from . import functions as fn
from . import ops
from .modules import AttentionLoss2d, LazyLogitsProb, LazyProb, LazySimMap, SADecoder, _init_like_reference


def continuous2discrete(depth, d_min, d_max, n_c):
    """model.py:19-23 (torch-0.4 semantics of the mask)."""
    return ops.continuous2discrete(depth, d_min, d_max, n_c)


def discrete2continuous(depth, d_min, d_max, n_c):
    """model.py:25-27 (elementwise on a handful of values; plain torch)."""
    return torch.exp(depth / (n_c - 1) * math.log(d_max / d_min) + math.log(d_min))


class BaseClassificationModel_(nn.Module):
    """model.py:29-102: decode_ord and the OR/CE soft/hard inference heads as streaming kernels."""

    def __init__(self, min_depth, max_depth, num_classes, classifierType, inferenceType, decoderType):
        super().__init__()
        self.min_depth = min_depth
        self.max_depth = max_depth
        self.num_classes = num_classes
        self.classifierType = classifierType
        self.inferenceType = inferenceType
        self.decoderType = decoderType

    def decode_ord(self, y):
        return fn.DecodeOrd.apply(y.float())

    def hard_cross_entropy(self, pred_score, d_min, d_max, n_c):
        return ops.head_inference(pred_score, "CE", "hard", d_min, d_max, n_c)

    def soft_cross_entropy(self, pred_score, d_min, d_max, n_c):
        return ops.head_inference(pred_score, "CE", "soft", d_min, d_max, n_c)

    def hard_ordinal_regression(self, pred_prob, d_min, d_max, n_c):
        return ops.head_inference(pred_prob, "OR", "hard", d_min, d_max, n_c)

    def soft_ordinal_regression(self, pred_prob, d_min, d_max, n_c):
        return ops.head_inference(pred_prob, "OR", "soft", d_min, d_max, n_c)

    def inference(self, y):
        if isinstance(y, list):
            y = y[-1]
        if isinstance(y, dict):
            y = y["y"]
        cls = "OR" if self.classifierType == "OR" else "CE"
        mode = "soft" if self.inferenceType == "soft" else "hard"
        if isinstance(y, LazyProb):                      # eval: fused upsample + decode_ord + head on the stride-8 logits
            return y.depth(mode)
        if isinstance(y, LazyLogitsProb):                # training-mode handle: decode_ord (and the upsampling) fused into the head
            if y.z is not None:
                return ops.tail_inference(y.z.detach(), cls, mode, self.min_depth, self.max_depth, self.num_classes)
            return ops.head_inference(y.logits.detach(), cls, mode, self.min_depth, self.max_depth, self.num_classes, from_logits=True)
        return ops.head_inference(y, cls, mode, self.min_depth, self.max_depth, self.num_classes)

    def forward(self):
        raise NotImplementedError


def make_decoder(height, width, command="attention"):
    if command != "attention":
        raise RuntimeError("decoder not found. The decoder must be attention.")
    return SADecoder(height=height, width=width)


def make_classifier(classifierType="OR", num_classes=80, use_inter=False, channel1=1024, channel2=2048):
    classes = 2 * num_classes if classifierType == "OR" else num_classes
    interout = None
    if use_inter:
        interout = nn.Sequential(OrderedDict([
            ("dropout1", nn.Dropout2d(0.2, inplace=True)),
            ("conv1", nn.Conv2d(channel1, channel1 // 2, kernel_size=3, stride=1, padding=1)),
            ("relu", nn.ReLU(inplace=True)),
            ("dropout2", nn.Dropout2d(0.2, inplace=False)),
            ("upsample", nn.Upsample(scale_factor=8, mode="bilinear", align_corners=True))]))
    classifier = nn.Sequential(OrderedDict([
        ("conv", nn.Conv2d(channel2, classes, kernel_size=1, stride=1, padding=0)),
        ("upsample", nn.Upsample(scale_factor=8, mode="bilinear", align_corners=True))]))
    return [interout, classifier]


def _conv_bn(cin, cout, k, stride=1, dilation=1):
    pad = dilation * (k // 2)
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, dilation=dilation, padding=pad, bias=False), nn.BatchNorm2d(cout)


class Bottleneck(nn.Module):
    """1x1 -> 3x3 (strided / dilated) -> 1x1 residual unit, width x4 at the output; cuDNN convolutions.
    Attribute names (conv1..3, bn1..3, downsample) are the checkpoint keys of model.py:128-160."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        wide = planes * self.expansion
        self.conv1, self.bn1 = _conv_bn(inplanes, planes, 1)
        self.conv2, self.bn2 = _conv_bn(planes, planes, 3, stride, dilation)
        self.conv3, self.bn3 = _conv_bn(planes, wide, 1)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        # fn.bn_act: BN (+ residual add) (+ ReLU) -- the fused batch-statistics kernels in train() on channels_last
        # half-precision activations (the autocast training step), the framework's own ops otherwise
        skip = x if self.downsample is None else fn.bn_act(self.downsample[1], self.downsample[0](x), relu=False)
        t = fn.bn_act(self.bn1, self.conv1(x), relu=True)
        t = fn.bn_act(self.bn2, self.conv2(t), relu=True)
        return fn.bn_act(self.bn3, self.conv3(t), relu=True, residual=skip)


# (planes, stride, dilation) of the four stages: output stride 8, layer3 dilated x2, layer4 x4 (model.py:177-180)
_STAGES = ((64, 1, 1), (128, 2, 1), (256, 1, 2), (512, 1, 4))


class ResNet(BaseClassificationModel_):
    def __init__(self, min_depth, max_depth, num_classes, classifierType, inferenceType, decoderType,
                 height, width, alpha=0, beta=0, layers=(3, 4, 6, 3), block=Bottleneck):
        super().__init__(min_depth, max_depth, num_classes, classifierType, inferenceType, decoderType)
        # registration order == reference child order: the optimizer factory slices children()[:8] / [8:]
        self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu1 = nn.ReLU(True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.inplanes = 64
        for i, ((planes, stride, dilation), count) in enumerate(zip(_STAGES, layers), start=1):
            setattr(self, f"layer{i}", self._make_layer(block, planes, count, stride=stride, dilation=dilation))
        self.alpha = alpha
        self.beta = beta
        self.decoder = make_decoder(height, width, decoderType)
        self.use_inter = self.alpha != 0.0
        self.interout, self.classifier = make_classifier(classifierType, num_classes, self.use_inter, channel1=1024, channel2=2048)
        self.lazy_head = True            # eval: 'y' is a LazyProb handle (False: always materialise like the reference)
        self.parameter_initialization()

    def parameter_initialization(self):
        _init_like_reference(self)          # model.py:189-197

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1):
        wide = planes * block.expansion
        project = None
        if stride != 1 or self.inplanes != wide:
            project = nn.Sequential(nn.Conv2d(self.inplanes, wide, kernel_size=1, stride=stride, bias=False), nn.BatchNorm2d(wide))
        units = [block(self.inplanes, planes, stride=stride, dilation=dilation, downsample=project)]
        self.inplanes = wide
        units.extend(block(wide, planes, stride=1, dilation=dilation) for _ in range(blocks - 1))
        return nn.Sequential(*units)

    def encode(self, x):
        """(layer3 output, layer4 output).  With the auxiliary branch in train() the reference runs ``interout`` -- whose first
        module is Dropout2d(0.2, inplace=True) -- BEFORE layer4 (model.py:216-222), so layer4 sees the dropped features.  Same here,
        but the dropout is applied out of place: layer3's output is saved by the ReLU / BatchNorm that produced it, and an in-place
        multiply on it makes autograd refuse the backward pass ("modified by an inplace operation")."""
        x = self.maxpool(fn.bn_act(self.bn1, self.conv1(x), relu=True))
        x = self.layer3(self.layer2(self.layer1(x)))
        if getattr(self, "use_inter", False) and self.interout is not None:
            d = self.interout.dropout1
            if d.training and d.p > 0:
                x = torch.nn.functional.dropout2d(x, d.p, True, False)
        return x, self.layer4(x)

    def _inter_logits(self, mid, with_upsample=False):
        """the auxiliary branch on ``mid`` = encode()[0] (its leading dropout already applied there) up to -- or, with_upsample,
        including -- its x8 upsampling (model.py:113-121)"""
        t = mid
        for name, m in self.interout.named_children():
            if name == "dropout1" or (name == "upsample" and not with_upsample):
                continue
            t = m(t)
        return t

    def forward(self, x):
        mid, x = self.encode(x)
        lazy_train = self.classifierType == "OR" and self.lazy_head
        inter_y = None
        if self.use_inter:
            if lazy_train and (self.training or torch.is_grad_enabled()) and hasattr(self.interout, "upsample"):
                # same lazy handle as 'y': the drop-in auxiliary loss works from the stride-8 logits
                # (the reference's auxiliary branch ends in its 3x3 convolution: channel1 // 2 = 512 channels, i.e. 256 ordinal pairs,
                #  whatever num_classes is -- model.py:113-121)
                z_inter = self._inter_logits(mid)
                inter_y = LazyLogitsProb(None, z_inter.shape[1] // 2, z=z_inter, upsample=self.interout.upsample)
            else:
                inter_y = self._inter_logits(mid, with_upsample=True)
        sim_map = None
        if self.decoderType == "attention":
            x, sim_map = self.decoder(x)
        if self.lazy_head and not self.training and not torch.is_grad_enabled():
            # eval: keep the stride-8 logits; inference() fuses upsample + decode_ord + head
            y = LazyProb(self.classifier.conv(x), "OR" if self.classifierType == "OR" else "CE",
                         self.min_depth, self.max_depth, self.num_classes)
            if self.use_inter and self.classifierType == "OR":
                inter_y = self.decode_ord(inter_y)
            return {"inter_y": inter_y, "sim_map": sim_map, "y": y}
        if lazy_train:
            # 'y' keeps the STRIDE-8 pair logits; the drop-in ordinal loss / inference fuse the x8 upsampling and decode_ord
            # (SURVEY §8f rows f-1, f-2); any other consumer gets decode_ord(upsample(z)) with the usual autograd
            y = LazyLogitsProb(None, self.num_classes, z=self.classifier.conv(x), upsample=self.classifier.upsample)
        else:
            y = self.classifier(x)
            if self.classifierType == "OR":
                y = self.decode_ord(y)
        if self.classifierType == "OR" and self.use_inter and not isinstance(inter_y, LazyLogitsProb):
            inter_y = self.decode_ord(inter_y)
        return {"inter_y": inter_y, "sim_map": sim_map, "y": y}

    @torch.no_grad()
    def predict_depth(self, x):
        """eval shortcut: forward + inference with decode_ord fused into the head kernel
        (the [B,K,H,W] probability tensor is never written).  Same result as inference(forward(x))."""
        _, feat = self.encode(x)
        feat, _ = self.decoder(feat)
        cls = "OR" if self.classifierType == "OR" else "CE"
        mode = "soft" if self.inferenceType == "soft" else "hard"
        return ops.tail_inference(self.classifier.conv(feat), cls, mode, self.min_depth, self.max_depth, self.num_classes)

    class LossFunc(nn.Module):
        """model.py:235-272: loss1 + alpha * loss2 + beta * loss3.  ``preds['sim_map']`` may be a LazySimMap."""

        def __init__(self, min_depth, max_depth, num_classes, AppearanceLoss=None, AuxiliaryLoss=None, AttentionLoss=None,
                     alpha=0., beta=0.):
            super().__init__()
            self.min_depth = min_depth
            self.max_depth = max_depth
            self.num_classes = num_classes
            self.alpha = alpha
            self.beta = beta
            self.AppearanceLoss = AppearanceLoss
            self.AuxiliaryLoss = AuxiliaryLoss
            self.AttentionLoss = AttentionLoss

        def forward(self, preds, label, epoch):
            inter_y = preds["inter_y"]
            sim_map = preds["sim_map"]
            y = preds["y"]
            dis_label = continuous2discrete(label, self.min_depth, self.max_depth, self.num_classes)
            loss1 = self.AppearanceLoss(y, dis_label.squeeze(1).long())
            loss2 = 0
            if self.alpha != 0.:
                loss2 = self.AuxiliaryLoss(inter_y, dis_label.squeeze(1).long())
            loss3 = 0
            if self.beta != 0.:
                loss3 = self.AttentionLoss(sim_map, dis_label)
            total_loss = loss1 + self.alpha * loss2 + self.beta * loss3
            return loss1, loss2, loss3, total_loss


class OrdinalRegression2d(nn.Module):
    """losses.py:64-158 with reduction='sum', no class weights (the reference's live configuration).  Given the lazy
    training-mode 'y' handle it runs the fused decode_ord + loss kernels (acan_ordinal_loss_fwd/bwd);
    given a dense probability tensor it is the plain torch restatement."""

    def __init__(self, ignore_index=None, reduction="sum", use_weights=False, weight=None):      # defaults as losses.py:144-145
        super().__init__()
        if use_weights or reduction != "sum":
            raise NotImplementedError("only the reference's default OrdinalRegression2d configuration is mirrored")
        self.ignore_index = ignore_index

    def forward(self, pred, label):
        if self.ignore_index is None:
            self.ignore_index = pred.shape[1] + 1                  # as the reference (losses.py:131-132)
        if isinstance(pred, LazyLogitsProb) and pred.z is not None:
            return fn.FusedTailLoss.apply(pred.z, label, self.ignore_index)              # fused with the x8 upsampling and decode_ord
        if isinstance(pred, LazyLogitsProb) and pred.shape[2] * pred.shape[3] % 4 == 0:
            return fn.FusedOrdinalLoss.apply(pred.logits, label, self.ignore_index)      # fused with decode_ord, one pass
        b, c, h, w = pred.shape
        k = torch.arange(c, device=pred.device).view(1, c, 1, 1)
        below = (k < label.view(b, 1, h, w)).to(pred.dtype)
        ent = -(torch.log(pred.clamp(1e-8, 1e8)) * below + torch.log((1 - pred).clamp(1e-8, 1e8)) * (1 - below))
        return ent.sum(dim=1)[label != self.ignore_index].mean()


class CrossEntropy2d(nn.Module):
    """losses.py:164-196 with reduction='sum', no class weights (the configuration models/get_lossfunc.py:24-26 builds): a per-class
    binary cross-entropy on softmax(scores) against the one-hot target, optionally smoothed (``eps``) with a uniform or gaussian prior.
    Runs the fused kernels (acan_ce_loss_fwd/bwd) on class scores [B,K,H,W]."""

    def __init__(self, ignore_index=None, reduction="sum", use_weights=False, weight=None, eps=0.0, priorType="uniform"):
        super().__init__()
        if use_weights or reduction != "sum":
            raise NotImplementedError("only the reference's default CrossEntropy2d configuration is mirrored")
        if priorType not in ("uniform", "gaussian"):
            raise ValueError(f"priorType {priorType!r}")
        self.ignore_index, self.eps, self.priorType = ignore_index, float(eps), priorType

    def forward(self, pred, label):
        if self.ignore_index is None:
            self.ignore_index = pred.shape[1] + 1                  # as the reference (losses.py:131-132)
        return fn.FusedCrossEntropyLoss.apply(pred, label, self.ignore_index, self.eps, self.priorType)


RESNET_LAYERS = {"resnet50": [3, 4, 6, 3], "resnet101": [3, 4, 23, 3]}


def create_network(args):
    """models/get_network.py:3-19."""
    if args.encoder not in RESNET_LAYERS:
        raise RuntimeError("network not found.The network must be either of resnet50 or resnet101.")
    return ResNet(min_depth=args.min_depth, max_depth=args.max_depth, num_classes=args.classes,
                  classifierType=args.classifier, inferenceType=args.inference, decoderType=args.decoder,
                  height=args.height, width=args.width, alpha=args.alpha, beta=args.beta, layers=RESNET_LAYERS[args.encoder])


def create_lossfunc(args, net):
    """models/get_lossfunc.py:8-41 for the OR and CE classifiers (the OHEM variant, which sorts on the host, is not mirrored)."""
    if args.classifier == "OR":
        app, aux = OrdinalRegression2d(ignore_index=0), OrdinalRegression2d(ignore_index=0)
    elif args.classifier == "CE":
        kw = dict(ignore_index=0, eps=getattr(args, "eps", 0.0), priorType=getattr(args, "prior", "uniform"))
        app, aux = CrossEntropy2d(**kw), CrossEntropy2d(**kw)
    else:
        raise NotImplementedError("acan_b200.create_lossfunc mirrors the OR and CE configurations; use the reference's loss for OHEM")
    return net.LossFunc(min_depth=args.min_depth, max_depth=args.max_depth, num_classes=args.classes,
                        AppearanceLoss=app, AuxiliaryLoss=aux, AttentionLoss=AttentionLoss2d(scale=1), alpha=args.alpha, beta=args.beta)


def install(net, criterion=None):
    """Swap the hot path of a network built by the REFERENCE's ``create_network`` for the B200 kernels, in place.

    Replaces ``net.decoder`` by a ``SADecoder`` carrying the same parameters/buffers (same names, so optimizers
    built afterwards and checkpoints are unaffected), rebinds ``decode_ord`` / ``inference`` and the four head
    functions, and swaps ``criterion.AttentionLoss``.  See INTEGRATION.md."""
    import types
    ref_dec = net.decoder
    h, w = ref_dec.image_context.avgpool.kernel_size
    new_dec = SADecoder(height=h * 8, width=w * 8)
    new_dec.load_state_dict(ref_dec.state_dict())
    new_dec.to(next(ref_dec.parameters()).device)
    new_dec.train(ref_dec.training)
    net.decoder = new_dec
    for name in ("decode_ord", "hard_cross_entropy", "soft_cross_entropy", "hard_ordinal_regression",
                 "soft_ordinal_regression", "inference"):
        setattr(net, name, types.MethodType(getattr(BaseClassificationModel_, name), net))
    if criterion is not None and getattr(criterion, "AttentionLoss", None) is not None:
        criterion.AttentionLoss = AttentionLoss2d(scale=1)
    app = getattr(criterion, "AppearanceLoss", None) if criterion is not None else None
    if (type(app).__name__ == "OrdinalRegression2d" and not getattr(app, "use_weights", False)
            and getattr(app, "reduction", "sum") == "sum" and getattr(app, "ignore_index", 0) is not None):
        criterion.AppearanceLoss = OrdinalRegression2d(ignore_index=app.ignore_index)     # fused decode_ord + loss kernels
    net.lazy_head = True
    fwd = ResNet.forward
    net.forward = types.MethodType(fwd, net)          # forward with the lazy 'y' handles (same dict, same keys)
    net.encode = types.MethodType(ResNet.encode, net)
    net._inter_logits = types.MethodType(ResNet._inter_logits, net)
    aux = getattr(criterion, "AuxiliaryLoss", None) if criterion is not None else None
    if (type(aux).__name__ == "OrdinalRegression2d" and not getattr(aux, "use_weights", False)
            and getattr(aux, "reduction", "sum") == "sum" and getattr(aux, "ignore_index", 0) is not None):
        criterion.AuxiliaryLoss = OrdinalRegression2d(ignore_index=aux.ignore_index)
    return net
