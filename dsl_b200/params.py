"""Parameter inventory of the FCOS detector with the reference's state_dict names and OIHW fp32 layout
(SURVEY.md §8b "State / ownership"), stored in ONE flat fp32 device buffer so EMA / SGD / grad all-reduce are single
launches. nn.Parameters / buffers of the plugin modules are views into it.

Regions of the flat buffer (in this order):
  A  trainable, base lr / weight decay      conv weights of layer2-4, FPN, head; GroupNorm gamma/beta; Scale
  B  trainable, bias lr x2 / weight decay 0  conv biases (mmcv DefaultOptimizerConstructor, paramwise_cfg
                                             bias_lr_mult=2, bias_decay_mult=0 — configs/fcos_semi/*.py optimizer)
  F  frozen parameters and buffers           stem, layer1 (frozen_stages=1), every BatchNorm (requires_grad=False,
                                             norm_eval=True) incl. running stats
"""
import math
from collections import OrderedDict

import torch

RESNET_BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}  # reference: mmdet/models/backbones/resnet.py:358-366


class P:
    __slots__ = ("name", "shape", "region", "kind")

    def __init__(self, name, shape, region, kind):
        self.name, self.shape, self.region, self.kind = name, tuple(shape), region, kind


def resnet_spec(depth=50, frozen_stages=1, prefix="backbone."):
    """Caffe-style bottleneck ResNet (resnet.py:304-656); names as in the reference's state_dict."""
    out = []

    def conv(name, o, i, k, frozen):
        out.append(P(prefix + name + ".weight", (o, i, k, k), "F" if frozen else "A", "conv"))

    def bn(name, c):
        out.append(P(prefix + name + ".weight", (c,), "F", "bn_w"))
        out.append(P(prefix + name + ".bias", (c,), "F", "bn_b"))
        out.append(P(prefix + name + ".running_mean", (c,), "F", "bn_mean"))
        out.append(P(prefix + name + ".running_var", (c,), "F", "bn_var"))

    conv("conv1", 64, 3, 7, True)
    bn("bn1", 64)
    inpl = 64
    for li, nb in enumerate(RESNET_BLOCKS[depth]):
        planes = 64 * 2 ** li
        frozen = (li + 1) <= frozen_stages
        for bi in range(nb):
            p = f"layer{li + 1}.{bi}"
            conv(p + ".conv1", planes, inpl, 1, frozen)
            bn(p + ".bn1", planes)
            conv(p + ".conv2", planes, planes, 3, frozen)
            bn(p + ".bn2", planes)
            conv(p + ".conv3", planes * 4, planes, 1, frozen)
            bn(p + ".bn3", planes * 4)
            if bi == 0:
                conv(p + ".downsample.0", planes * 4, inpl, 1, frozen)
                bn(p + ".downsample.1", planes * 4)
            inpl = planes * 4
    return out


RLA_CHANNEL = 32  # resnet_rla.py:9,162: rla_channel, the width of the recurrent state h


def rla_resnet_spec(layers=(3, 4, 6, 3), frozen_stages=1, prefix="backbone."):
    """RLA_ResNet (resnet_rla.py:140-400): PyTorch-style bottlenecks (stride on conv2) whose conv1 acts on cat(x, h),
    plus per stage a shared 1x1 `conv_outs` (4*planes -> 32), a shared 3x3 `recurrent_convs` (32 -> 32) and one
    BatchNorm(32) per block (`stage_bns`). Names as in the reference's state_dict. Unlike ResNet(norm_cfg requires_grad=
    False) the BatchNorm affine parameters of the non-frozen stages are TRAINABLE (only the statistics are frozen by
    norm_eval, :389-399), except stage_bns[3][2] (:375-377). Listed stage by stage (conv_outs first, recurrent_convs
    last) so that every stage is one contiguous range of the flat gradient (the gradient buckets of the backward)."""
    out = []

    def conv(name, o, i, k, frozen):
        out.append(P(prefix + name + ".weight", (o, i, k, k), "F" if frozen else "A", "conv"))

    def bn(name, c, frozen):
        out.append(P(prefix + name + ".weight", (c,), "F" if frozen else "A", "bn_w"))
        out.append(P(prefix + name + ".bias", (c,), "F" if frozen else "A", "bn_b"))
        out.append(P(prefix + name + ".running_mean", (c,), "F", "bn_mean"))
        out.append(P(prefix + name + ".running_var", (c,), "F", "bn_var"))

    conv("conv1", 64, 3, 7, True)
    bn("bn1", 64, True)
    inpl = 64
    for li, nb in enumerate(layers):
        planes = 64 * 2 ** li
        frozen = (li + 1) <= frozen_stages
        conv(f"conv_outs.{li}", RLA_CHANNEL, planes * 4, 1, frozen)
        for bi in range(nb):
            p = f"stages.{li}.{bi}"
            conv(p + ".conv1", planes, inpl + RLA_CHANNEL, 1, frozen)
            bn(p + ".bn1", planes, frozen)
            conv(p + ".conv2", planes, planes, 3, frozen)
            bn(p + ".bn2", planes, frozen)
            conv(p + ".conv3", planes * 4, planes, 1, frozen)
            bn(p + ".bn3", planes * 4, frozen)
            if bi == 0:
                conv(p + ".downsample.0", planes * 4, inpl, 1, frozen)
                bn(p + ".downsample.1", planes * 4, frozen)
            inpl = planes * 4
        for bi in range(nb):
            bn(f"stage_bns.{li}.{bi}", RLA_CHANNEL, frozen or (li == 3 and bi == 2))
        conv(f"recurrent_convs.{li}", RLA_CHANNEL, RLA_CHANNEL, 3, frozen)
    return out


def fpn_spec(in_channels=(512, 1024, 2048), out_channels=256, prefix="neck."):
    """FPN(start_level=1, add_extra_convs='on_output', num_outs=5) (necks/fpn.py:61-149)."""
    out = []
    for i, c in enumerate(in_channels):
        out.append(P(f"{prefix}lateral_convs.{i}.conv.weight", (out_channels, c, 1, 1), "A", "conv"))
        out.append(P(f"{prefix}lateral_convs.{i}.conv.bias", (out_channels,), "B", "bias"))
    for i in range(5):
        out.append(P(f"{prefix}fpn_convs.{i}.conv.weight", (out_channels, out_channels, 3, 3), "A", "conv"))
        out.append(P(f"{prefix}fpn_convs.{i}.conv.bias", (out_channels,), "B", "bias"))
    return out


def head_spec(num_classes=80, in_channels=256, feat_channels=256, stacked_convs=4, num_levels=5, prefix="bbox_head."):
    """FCOSHead (dense_heads/fcos_head.py:112-116, anchor_free_head.py:89-139)."""
    out = []
    for br in ("cls_convs", "reg_convs"):
        for i in range(stacked_convs):
            cin = in_channels if i == 0 else feat_channels
            out.append(P(f"{prefix}{br}.{i}.conv.weight", (feat_channels, cin, 3, 3), "A", "conv"))
            out.append(P(f"{prefix}{br}.{i}.conv.bias", (feat_channels,), "B", "bias"))
            out.append(P(f"{prefix}{br}.{i}.gn.weight", (feat_channels,), "A", "gn_w"))
            out.append(P(f"{prefix}{br}.{i}.gn.bias", (feat_channels,), "A", "gn_b"))
    out.append(P(prefix + "conv_cls.weight", (num_classes, feat_channels, 3, 3), "A", "conv"))
    out.append(P(prefix + "conv_cls.bias", (num_classes,), "B", "bias_cls"))
    out.append(P(prefix + "conv_reg.weight", (4, feat_channels, 3, 3), "A", "conv"))
    out.append(P(prefix + "conv_reg.bias", (4,), "B", "bias"))
    out.append(P(prefix + "conv_centerness.weight", (1, feat_channels, 3, 3), "A", "conv"))
    out.append(P(prefix + "conv_centerness.bias", (1,), "B", "bias"))
    for i in range(num_levels):
        out.append(P(f"{prefix}scales.{i}.scale", (), "A", "scale"))
    return out


class ParamStore:
    """Flat fp32 storage + named views. `spec` order = the reference's state_dict order (kept for load/save)."""

    ALIGN = 4  # floats: every tensor starts 16-byte aligned

    def __init__(self, spec, device):
        self.spec = list(spec)
        self.device = torch.device(device)
        off = 0
        self.offsets = {}
        self.region_range = {}
        for region in ("A", "B", "F"):
            start = off
            for p in self.spec:
                if p.region != region:
                    continue
                n = int(math.prod(p.shape)) if p.shape else 1
                self.offsets[p.name] = (off, n)
                off += (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            self.region_range[region] = (start, off)
        self.numel = off
        self.n_train = self.region_range["B"][1]  # regions A + B are contiguous from 0
        self.flat = torch.zeros(self.numel, dtype=torch.float32, device=self.device)
        self.views = OrderedDict()
        for p in self.spec:
            o, n = self.offsets[p.name]
            self.views[p.name] = self.flat[o:o + n].view(p.shape)

    def __getitem__(self, name):
        return self.views[name]

    def num_params(self):
        return sum(int(math.prod(p.shape)) if p.shape else 1 for p in self.spec)

    def state_dict(self):
        """Reference-named state dict (adds the BatchNorm num_batches_tracked entries the reference carries)."""
        sd = OrderedDict()
        for p in self.spec:
            sd[p.name] = self.views[p.name]
            if p.kind == "bn_var":
                sd[p.name.replace("running_var", "num_batches_tracked")] = torch.zeros((), dtype=torch.long,
                                                                                       device=self.device)
        return sd

    @torch.no_grad()
    def load_state_dict(self, sd, strict=True):
        missing = []
        for p in self.spec:
            if p.name in sd:
                self.views[p.name].copy_(sd[p.name].to(self.device, torch.float32).reshape(p.shape))
            else:
                missing.append(p.name)
        if strict and missing:
            raise KeyError(f"missing keys in state_dict: {missing[:5]} ... ({len(missing)} total)")
        return missing

    @torch.no_grad()
    def init_reference(self, seed=0):
        """Synthetic initialisation following the reference's init_cfg where it has one: head Normal(0, .01) with
        conv_cls bias = -log(99) (fcos_head.py:83-91), FPN Xavier-uniform (fpn.py:74-75), backbone Kaiming-normal
        (resnet.py:402-409); BatchNorm statistics are synthetic (no checkpoint is available offline)."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        for p in self.spec:
            v = self.views[p.name]
            if p.kind == "conv":
                o, i, kh, kw = p.shape
                if "bbox_head" in p.name:
                    t = torch.randn(p.shape, generator=g) * 0.01
                elif "neck" in p.name or "lateral_convs" in p.name or "fpn_convs" in p.name:
                    bound = math.sqrt(6.0 / (i * kh * kw + o * kh * kw))
                    t = (torch.rand(p.shape, generator=g) * 2 - 1) * bound
                else:
                    t = torch.randn(p.shape, generator=g) * math.sqrt(2.0 / (o * kh * kw))
                v.copy_(t)
            elif p.kind == "bias_cls":
                v.fill_(-math.log((1 - 0.01) / 0.01))
            elif p.kind in ("bias", "gn_b", "bn_b", "bn_mean"):
                v.zero_()
            elif p.kind in ("gn_w", "bn_var", "scale"):
                v.fill_(1.0)
            elif p.kind == "bn_w":
                # the last norm of each residual branch gets a small gain so activations stay bounded through 16+
                # residual additions with random weights (a pretrained checkpoint would do the same job)
                v.fill_(0.25 if p.name.endswith("bn3.weight") else 1.0)
        return self
