"""Drop-in boundary: the reference's own operator / plugin interface for the hot path, answered by libdslb.so.

The reference resolves every model class by `type` string through one mmdet registry (mmdet/models/builder.py:6-14)
and lets a config pull extra modules in with `custom_imports` (tools/train.py:93-95). Importing this module
re-registers, with force=True, the keys the fcos_semi configs name:

    DETECTORS['FCOS']  -> FCOS          (mmdet/models/detectors/{fcos,single_stage,base}.py)
    BACKBONES['ResNet'], NECKS['FPN'] -> ResNet, FPN   (backbones/resnet.py, necks/fpn.py; standalone modules)
    LOSSES[...] / RUNNERS['SemiEpochBasedRunner'] -> dsl_b200.losses / dsl_b200.runner
    HEADS['FCOSHead']  -> FCOSHead      (mmdet/models/dense_heads/{fcos_head,anchor_free_head}.py)
    HOOKS['EMAOWNHook']-> EMAOWNHook    (mmdet/runner/hooks/ema.py + SemiEpochBasedRunner.EMA,
                                         mmdet/runner/hooks/semi_epoch_based_runner.py:368-409)

so `tools/train.py <cfg> --cfg-options custom_imports.imports=[dsl_b200.plugin]` runs unchanged. Same constructor
kwargs, same state_dict names and OIHW fp32 layout (checkpoints load both ways), same call signatures and the same
exceptions for bad input. Parameters are nn.Parameters that VIEW one flat fp32 buffer (dsl_b200.params.ParamStore), so
torch optimizers, DDP, EMA and checkpointing address them as usual while the kernels see contiguous memory.

There is NO eager / CPU fallback: the modules construct on any device (so configs can be built and inspected on a
CPU box), but forward / loss / get_bboxes need CUDA and raise otherwise. Unsupported architectural options raise
NotImplementedError at construction — never a silent different path.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from .params import RESNET_BLOCKS, ParamStore, fpn_spec, head_spec, resnet_spec, rla_resnet_spec

INF = 1e8
_DEFAULT_RANGES = ((-1, 64), (64, 128), (128, 256), (256, 512), (512, INF))


def _need_cuda(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"dsl_b200.{what} runs only on CUDA tensors (sm_100a kernels in libdslb.so; there is no "
                           "CPU / eager fallback)")


class _Box(nn.Module):
    """Name-only container, so parameters keep the reference's dotted state_dict names."""


class _StoreModule(nn.Module):
    """nn.Module whose parameters / buffers are views into a ParamStore's flat fp32 buffer."""

    def _bind_store(self, store, prefix_strip=""):
        self.store = store
        self._views = []
        for p in store.spec:
            name = p.name[len(prefix_strip):] if prefix_strip and p.name.startswith(prefix_strip) else p.name
            parts = name.split(".")
            mod = self
            for q in parts[:-1]:
                if not hasattr(mod, q):
                    mod.add_module(q, _Box())
                mod = getattr(mod, q)
            v = store.views[p.name]
            if p.kind in ("bn_mean", "bn_var"):
                mod.register_buffer(parts[-1], v)
                if p.kind == "bn_var":
                    mod.register_buffer("num_batches_tracked", torch.zeros((), dtype=torch.long, device=v.device))
            else:
                mod.register_parameter(parts[-1], nn.Parameter(v, requires_grad=p.region in ("A", "B")))
            self._views.append((mod, parts[-1], p))

    def _apply(self, fn, *a, **k):
        # .to() / .cuda() / .float(): move the FLAT buffer once, then re-point every parameter at its new view
        # (nn.Module._apply would give each parameter its own storage and break the flat layout the kernels rely on).
        st = self.store
        new_flat = fn(st.flat)
        if new_flat.dtype != torch.float32:
            raise TypeError("dsl_b200 keeps fp32 master parameters (bf16 operands are derived caches)")
        if new_flat is not st.flat:
            st.flat = new_flat
            st.device = new_flat.device
            for p in st.spec:
                o, n = st.offsets[p.name]
                st.views[p.name] = st.flat[o:o + n].view(p.shape)
            self._on_store_moved()
        for mod, leaf, p in self._views:
            v = st.views[p.name]
            if leaf in mod._parameters:
                mod._parameters[leaf].data = v
                if mod._parameters[leaf].grad is not None:
                    mod._parameters[leaf].grad = None
            else:
                mod._buffers[leaf] = v
                if p.kind == "bn_var":
                    mod._buffers["num_batches_tracked"] = fn(mod._buffers["num_batches_tracked"])
        return self

    def _on_store_moved(self):
        pass

    def trainable_parameters(self):
        """[(spec entry, nn.Parameter)] in flat-buffer order (regions A then B)."""
        out = []
        for mod, leaf, p in self._views:
            if leaf in mod._parameters and p.region in ("A", "B"):
                out.append((p, mod._parameters[leaf]))
        out.sort(key=lambda t: self.store.offsets[t[0].name][0])
        return out


def _check(cond, msg):
    if not cond:
        raise NotImplementedError("dsl_b200 plugin: " + msg)


def _loss_cfg(cfg, typ, **expect):
    cfg = dict(cfg or {})
    _check(cfg.get("type", typ) == typ, f"loss type {cfg.get('type')} (only {typ} is implemented in the fused kernel)")
    for k, v in expect.items():
        _check(cfg.get(k, v) == v, f"{typ}.{k}={cfg.get(k)} (kernel implements {v})")
    _check(float(cfg.get("loss_weight", 1.0)) == 1.0, f"{typ}.loss_weight != 1.0")
    return cfg


# ====================================================================================================== autograd
class _TrainStep(torch.autograd.Function):
    """forward: CUDA forward + targets + loss (+ gradient of the head outputs); backward: the CUDA backward plan.
    Inputs after the two python objects are the trainable parameters, so autograd hands their gradients to the
    optimizer / DDP exactly as for the reference's eager module."""

    @staticmethod
    def forward(ctx, owner, net, *params):
        ctx.owner, ctx.net = owner, net
        losses = net.losses()
        keys = list(losses.keys())
        ctx.nloss = len(keys)
        owner._loss_keys = keys
        return tuple(losses[k].clone() for k in keys)

    @staticmethod
    def backward(ctx, *gouts):
        # The fused loss kernel already produced d(sum of losses)/d(head outputs); the upstream gradient of every loss
        # term is therefore assumed equal (loss = sum of the returned losses, BaseDetector._parse_losses) and applied as
        # ONE scalar scale on the flat gradient buffer. owner.check_grad_outputs=True verifies it (costs a host sync).
        net, owner = ctx.net, ctx.owner
        g0 = next(g for g in gouts if g is not None)
        if owner.check_grad_outputs:
            for g in gouts:
                if g is not None and not torch.equal(g, g0):
                    raise NotImplementedError("dsl_b200: the fused backward needs the same upstream gradient for every "
                                              "loss term")
        net.backward()
        net.grad.mul_(g0.to(torch.float32))
        grads = []
        for p, _ in owner._trainable:
            o, n = net.store.offsets[p.name]
            grads.append(net.grad[o:o + n].view(p.shape))
        return (None, None) + tuple(grads)


# ====================================================================================================== detector
class FCOS(_StoreModule):
    """FCOS detector (mmdet/models/detectors/fcos.py:5-17, single_stage.py:10-203, base.py:17-243) on libdslb.so.

    Same config schema as the reference (`configs/fcos_semi/*.py` model dict). Supported architecture = what those
    configs name: ResNet-50/101 caffe style, frozen BN, frozen_stages=1; FPN start_level=1, extra convs on_output,
    5 outs, ReLU before the extra convs; FCOSHead with 4 stacked GN convs, centerness on the regression branch.
    """

    def __init__(self, backbone, neck, bbox_head, train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None):
        super().__init__()
        bb, nk, hd = dict(backbone), dict(neck), dict(bbox_head)
        _check(bb.get("type", "ResNet") in ("ResNet", "RLA_ResNet"), f"backbone type {bb.get('type')}")
        self.backbone_kind = "rla" if bb.get("type") == "RLA_ResNet" else "resnet"
        if self.backbone_kind == "rla":      # configs/fcos_semi/RLA_*.py:3-13
            bb["depth"] = RLA_ResNet.check_cfg(bb)
        else:
            _check(bb.get("depth") in (50, 101), f"ResNet depth {bb.get('depth')}")
            _check(bb.get("style", "pytorch") == "caffe", "ResNet style must be 'caffe' (stride on conv1)")
            _check(bb.get("norm_eval", True) and not dict(bb.get("norm_cfg", {})).get("requires_grad", True),
                   "BatchNorm must be frozen (norm_eval=True, requires_grad=False)")
            _check(tuple(bb.get("out_indices", (0, 1, 2, 3))) == (0, 1, 2, 3), "out_indices")
        _check(bb.get("frozen_stages", -1) == 1, "frozen_stages must be 1")
        self.pretrained = bb.get("pretrained", pretrained)
        _check(nk.get("type", "FPN") == "FPN" and list(nk.get("in_channels")) == [256, 512, 1024, 2048]
               and nk.get("out_channels") == 256 and nk.get("start_level", 0) == 1 and nk.get("num_outs") == 5
               and nk.get("add_extra_convs") == "on_output" and nk.get("relu_before_extra_convs", False),
               "FPN config differs from configs/fcos_semi (start_level=1, on_output, 5 outs, relu_before_extra_convs)")
        self.depth = bb["depth"]
        self.head_cfg = FCOSHead.parse_cfg(hd)
        self.num_classes = self.head_cfg["num_classes"]
        self.train_cfg, self.test_cfg = train_cfg, dict(test_cfg or {})
        bb_spec = rla_resnet_spec(RESNET_BLOCKS[self.depth]) if self.backbone_kind == "rla" else resnet_spec(self.depth)
        store = ParamStore(bb_spec + fpn_spec() + head_spec(self.num_classes), "cpu").init_reference(0)
        self._bind_store(store)
        self._trainable = self.trainable_parameters()
        self._nets = OrderedDict()   # (B, H, W, train) -> FCOSNet
        self._posts = OrderedDict()
        self.check_grad_outputs = False
        self.cur_iter = 0            # FCOSHead.cur_iter (soft warm-up counter, fcos_head.py:100,325-327)
        self.fp16_enabled = False

    # ---- reference API surface ------------------------------------------------------------------------------
    @property
    def with_neck(self):
        return True

    @property
    def with_bbox(self):
        return True

    def init_weights(self):
        """Reference init_cfg (synthetic, no checkpoint offline): see ParamStore.init_reference. A backbone `pretrained`
        path that names an existing checkpoint file is loaded into `backbone.*` non-strictly, as
        RLA_ResNet.init_weights does (resnet_rla.py:379-388; single_stage.py:32-36 hands `pretrained` to the backbone)."""
        import os
        self.store.init_reference(0)
        if isinstance(self.pretrained, str) and os.path.isfile(self.pretrained):
            ck = torch.load(self.pretrained, map_location="cpu")
            ck = ck.get("state_dict", ck)
            self.store.load_state_dict({"backbone." + k: v for k, v in ck.items()
                                        if not k.endswith("num_batches_tracked")}, strict=False)
        self._dirty()

    def _on_store_moved(self):
        self._nets.clear()
        self._posts.clear()

    def _dirty(self):
        for net in self._nets.values():
            net._stale = True

    def load_state_dict(self, sd, strict=True):
        sd = {k: v for k, v in sd.items() if not k.endswith("num_batches_tracked")}
        missing = self.store.load_state_dict(sd, strict=strict)
        self._dirty()
        return missing

    def _net(self, B, H, W, train):
        from .engine import FCOSNet
        key = (B, H, W, bool(train))
        if key not in self._nets:
            if len(self._nets) >= 4:   # activation plans are GBs at 800x1344: keep a small LRU of shapes
                self._nets.popitem(last=False)
            hc = self.head_cfg
            self._nets[key] = FCOSNet(B, H, W, depth=self.depth, num_classes=self.num_classes, train=train,
                                      store=self.store, device=self.store.device, loss_weight=hc["loss_weight"],
                                      soft_weight=hc["soft_weight"], center_sampling=hc["center_sampling"],
                                      radius=hc["center_sample_radius"], norm_on_bbox=hc["norm_on_bbox"],
                                      strides=hc["strides"], regress_ranges=hc["regress_ranges"],
                                      backbone=self.backbone_kind,
                                      head_precision="bf16" if train else self.eval_head_precision)
        else:
            self._nets.move_to_end(key)
        return self._nets[key]

    # inference precision of the FCOSHead: "bf16" (default) or "bf16x3" = split-bf16 operands with fp32 tower maps, which
    # holds the reference's fp32 outputs (fcos_head.py:118-168) to 1e-3 at 3x the tower FLOPs (engine._build_head_split).
    # Also settable for a whole process with DSLB_HEAD_PRECISION=bf16x3.
    eval_head_precision = __import__("os").environ.get("DSLB_HEAD_PRECISION", "bf16")

    def set_eval_head_precision(self, precision):
        assert precision in ("bf16", "bf16x3")
        self.eval_head_precision = precision
        self._nets.clear()

    def extract_feat(self, img):
        """FPN outputs as the reference returns them: 5 x (B, 256, h, w) fp32 NCHW (single_stage.py:136-141)."""
        from . import _lib as L
        _need_cuda(img, "FCOS.extract_feat")
        B, _, H, W = img.shape
        net = self._net(B, H, W, train=False)
        with torch.no_grad():
            net.repack(everything=True)
            net.img.copy_(img)
            for op in net.fwd_ops[:net.head_op_start]:
                op()
            outs = []
            for l, (h, w) in enumerate(net.psize):
                o = torch.empty(B, 256, h, w, dtype=torch.float32, device=img.device)
                L.check(L.lib.dslb_nhwc_to_nchw_f32(L.ptr(net.p[l]), L.ptr(o), B, 256, h, w, 256, 0, L.cur_stream()),
                        "nhwc_to_nchw")
                outs.append(o)
        return tuple(outs)

    def forward(self, img, img_metas, return_loss=True, **kwargs):
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        return self.forward_test(img, img_metas, **kwargs)

    def forward_train(self, img, img_metas, gt_bboxes, gt_labels, gt_bboxes_ignore=None):
        """single_stage.py:152-180 + base_dense_head.py:22-59 + FCOSHead.loss (fcos_head.py:170-338)."""
        _need_cuda(img, "FCOS.forward_train")
        B, _, H, W = img.shape
        assert len(gt_bboxes) == B and len(gt_labels) == B
        hc = self.head_cfg
        if hc["loss_weight"] != 1.0 and gt_bboxes_ignore is None:
            # the reference evaluates len(gt_bboxes_ignore) here (fcos_head.py:224)
            raise TypeError("object of type 'NoneType' has no len()")
        net = self._net(B, H, W, train=True)
        net.repack(everything=getattr(net, "_stale", True))
        net._stale = False
        net.img.copy_(img)
        net.si_weight = 0.0
        if B % 2 == 1 and hc["soft_weight"] != 0.0:     # scale-invariant soft loss branch (fcos_head.py:312-333)
            if hc["soft_warm_up"] >= self.cur_iter:
                net.si_weight = hc["soft_weight"] / 1000.0
                self.cur_iter += 1
            else:
                net.si_weight = hc["soft_weight"]
        with torch.no_grad():
            net.forward()
            net.set_targets(gt_bboxes, gt_labels, gt_bboxes_ignore)
            net.run_targets()
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                net.world_size = float(dist.get_world_size())
                dist.all_reduce(net.counts)          # ONE packed all-reduce for num_pos and sum(centerness targets)
            net.run_loss()
        outs = _TrainStep.apply(self, net, *[p for _, p in self._trainable])
        return OrderedDict(zip(self._loss_keys, outs))

    def forward_test(self, imgs, img_metas, **kwargs):
        if isinstance(imgs, (list, tuple)):
            assert len(imgs) == 1, "test-time augmentation is not part of the DSL path"
            imgs, img_metas = imgs[0], img_metas[0]
        return self.simple_test(imgs, img_metas, **kwargs)

    def _post(self, net):
        from .postprocess import TeacherPost
        key = (net.B, tuple(net.psize))
        if key not in self._posts:
            tc = self.test_cfg
            self._posts[key] = TeacherPost(net.B, net.psize, self.head_cfg["strides"], self.num_classes,
                                           self.store.device, nms_pre=tc.get("nms_pre", 1000),
                                           score_thr=tc.get("score_thr", 0.05),
                                           iou_thr=dict(tc.get("nms", {})).get("iou_threshold", 0.5),
                                           max_per_img=tc.get("max_per_img", 100))
        return self._posts[key]

    @torch.no_grad()
    def simple_test(self, img, img_metas, rescale=False):
        """single_stage.py:182-203: forward (eval: bbox x stride) -> get_bboxes -> bbox2result per image."""
        _need_cuda(img, "FCOS.simple_test")
        B, _, H, W = img.shape
        net = self._net(B, H, W, train=False)
        net.repack(everything=True)
        net.img.copy_(img)
        net.forward()
        post = self._post(net)
        post.set_meta([m["img_shape"] for m in img_metas],
                      [np.asarray(m["scale_factor"]).reshape(-1).tolist() for m in img_metas] if rescale else None)
        post.decode(net.cls_out, net.rc_out)
        post.nms()
        out = []
        for dets, labels in post.results():
            d, l = dets.numpy(), labels.numpy()
            out.append([d[l == i, :] for i in range(self.num_classes)])   # core/bbox/transforms.py:101-116
        return out

    def _parse_losses(self, losses):
        """base.py:175-208, with the per-key all-reduces packed into one."""
        from .dist_ops import reduce_log_vars
        log_vars = OrderedDict((k, v.mean()) for k, v in losses.items())
        loss = sum(v for k, v in log_vars.items() if "loss" in k)
        log_vars["loss"] = loss
        return loss, reduce_log_vars(log_vars)

    def train_step(self, data, optimizer):
        losses = self(**data)
        loss, log_vars = self._parse_losses(losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data["img_metas"]))

    def val_step(self, data, optimizer=None):
        return self.train_step(data, optimizer)


# ====================================================================================================== backbone / neck
def _to_nchw(t_nhwc, B, C, h, w, dev):
    from . import _lib as L
    o = torch.empty(B, C, h, w, dtype=torch.float32, device=dev)
    L.check(L.lib.dslb_nhwc_to_nchw_f32(L.ptr(t_nhwc), L.ptr(o), B, C, h, w, C, 0, L.cur_stream()), "nhwc_to_nchw")
    return o


def _to_nhwc(t_nchw, dst, B, C, h, w):
    from . import _lib as L
    L.check(L.lib.dslb_nchw_to_nhwc_bf16(L.ptr(t_nchw.contiguous().float()), L.ptr(dst), B, C, h, w, C, L.cur_stream()),
            "nchw_to_nhwc")


class _PartFn(torch.autograd.Function):
    """Autograd bridge of the standalone ResNet / FPN modules: forward already ran on the plan `net`; backward copies the
    upstream gradients into the plan's seed buffers, runs its CUDA backward and hands back input + parameter gradients."""

    @staticmethod
    def forward(ctx, owner, net, outs, seeds, in_grads, n_in, *inputs_and_params):
        ctx.owner, ctx.net, ctx.seeds, ctx.in_grads, ctx.n_in = owner, net, seeds, in_grads, n_in
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        net, owner = ctx.net, ctx.owner
        for g, (buf, B, C, h, w) in zip(gouts, ctx.seeds):
            if buf is None:
                continue
            if g is None:
                buf.zero_()
            else:
                _to_nhwc(g, buf, B, C, h, w)
        net.backward()
        gin = [None if e is None else _to_nchw(e[0], *e[1:], owner.store.device) for e in ctx.in_grads]
        grads = []
        for p, _ in owner._trainable:
            o, n = net.store.offsets[p.name]
            grads.append(net.grad[o:o + n].view(p.shape).clone())
        return (None, None, None, None, None, None) + tuple(gin) + tuple(grads)


class ResNet(_StoreModule):
    """BACKBONES['ResNet'] (mmdet/models/backbones/resnet.py:304-656) for the configuration the fcos_semi configs use:
    depth 50 / 101, caffe style, frozen BatchNorm (norm_eval, requires_grad=False), frozen_stages=1. NCHW fp32 in,
    tuple of NCHW fp32 stage outputs (C2..C5 at `out_indices`) out, autograd-connected for layers 2-4."""

    def __init__(self, depth, in_channels=3, num_stages=4, strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1),
                 out_indices=(0, 1, 2, 3), style="pytorch", frozen_stages=-1, norm_cfg=None, norm_eval=True,
                 init_cfg=None, pretrained=None, **kwargs):
        super().__init__()
        _check(depth in (50, 101), f"ResNet depth {depth}")
        _check(in_channels == 3 and num_stages == 4 and tuple(strides) == (1, 2, 2, 2) and
               tuple(dilations) == (1, 1, 1, 1), "only the standard 4-stage stride-(1,2,2,2) layout")
        _check(style == "caffe", "ResNet style must be 'caffe' (stride on conv1)")
        _check(frozen_stages == 1, "frozen_stages must be 1")
        _check(norm_eval and not dict(norm_cfg or {}).get("requires_grad", True),
               "BatchNorm must be frozen (norm_eval=True, requires_grad=False)")
        _check(not kwargs.get("dcn") and not kwargs.get("plugins") and not kwargs.get("deep_stem") and
               not kwargs.get("avg_down") and not kwargs.get("with_cp"), "dcn / plugins / deep_stem / avg_down / with_cp")
        self.depth, self.out_indices = depth, tuple(out_indices)
        self._bind_store(ParamStore(resnet_spec(depth, prefix=""), "cpu").init_reference(0))
        self._trainable = self.trainable_parameters()
        self._nets = OrderedDict()

    def _on_store_moved(self):
        self._nets.clear()

    def init_weights(self):
        self.store.init_reference(0)
        self._nets.clear()

    def load_state_dict(self, sd, strict=True):
        missing = self.store.load_state_dict({k: v for k, v in sd.items() if not k.endswith("num_batches_tracked")},
                                             strict=strict)
        for n in self._nets.values():
            n._stale = True
        return missing

    def _net(self, B, H, W, train):
        from .engine import FCOSNet
        key = (B, H, W, bool(train))
        if key not in self._nets:
            if len(self._nets) >= 4:
                self._nets.popitem(last=False)
            self._nets[key] = FCOSNet(B, H, W, depth=self.depth, train=train, store=self.store,
                                      device=self.store.device, parts="backbone",
                                      backbone=getattr(self, "backbone_kind", "resnet"))
        return self._nets[key]

    def forward(self, x):
        _need_cuda(x, "ResNet.forward")
        B, _, H, W = x.shape
        train = torch.is_grad_enabled() and self.training
        net = self._net(B, H, W, train)
        with torch.no_grad():
            net.repack(everything=True)     # parameters may have been stepped by any optimizer since the last call
            net.img.copy_(x)
            net.forward()
            outs = [_to_nchw(t, B, c, h, w, x.device) for (t, h, w, c) in net.stage_out]
        if train:
            seeds = [(None, 0, 0, 0, 0)] + [(net.gc[i], B, c, h, w) for i, (_, h, w, c) in enumerate(net.stage_out[1:])]
            outs = _PartFn.apply(self, net, outs, seeds, [], 0, *[p for _, p in self._trainable])
        return tuple(outs[i] for i in self.out_indices)


class FPN(_StoreModule):
    """NECKS['FPN'] (mmdet/models/necks/fpn.py:9-202) for the fcos_semi layout: in_channels [256, 512, 1024, 2048],
    start_level=1, add_extra_convs='on_output', num_outs=5, relu_before_extra_convs=True. Tuple of NCHW fp32 C2..C5 in
    (C2 is unused, as in the reference with start_level=1), tuple of five NCHW fp32 maps out, autograd-connected."""

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False,
                 relu_before_extra_convs=False, no_norm_on_lateral=False, conv_cfg=None, norm_cfg=None, act_cfg=None,
                 upsample_cfg=None, init_cfg=None, **kwargs):
        super().__init__()
        _check(list(in_channels) == [256, 512, 1024, 2048] and out_channels == 256 and start_level == 1 and num_outs == 5
               and add_extra_convs == "on_output" and relu_before_extra_convs and end_level in (-1, 4)
               and norm_cfg is None and act_cfg is None and conv_cfg is None,
               "FPN config differs from configs/fcos_semi (start_level=1, on_output, 5 outs, relu_before_extra_convs)")
        _check(dict(upsample_cfg or dict(mode="nearest")).get("mode") == "nearest", "FPN upsample mode must be nearest")
        self.in_channels, self.out_channels, self.num_outs = list(in_channels), out_channels, num_outs
        self._bind_store(ParamStore(fpn_spec(prefix=""), "cpu").init_reference(0))
        self._trainable = self.trainable_parameters()
        self._nets = OrderedDict()

    def _on_store_moved(self):
        self._nets.clear()

    def init_weights(self):
        self.store.init_reference(0)
        self._nets.clear()

    def load_state_dict(self, sd, strict=True):
        return self.store.load_state_dict(sd, strict=strict)

    def _net(self, B, sizes, train):
        from .engine import FCOSNet
        key = (B, tuple(sizes), bool(train))
        if key not in self._nets:
            if len(self._nets) >= 4:
                self._nets.popitem(last=False)
            self._nets[key] = FCOSNet(B, 0, 0, train=train, store=self.store, device=self.store.device, parts="neck",
                                      level_sizes=list(sizes))
        return self._nets[key]

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        feats = list(inputs[1:])
        _need_cuda(feats[0], "FPN.forward")
        B = feats[0].shape[0]
        sizes = [tuple(f.shape[-2:]) for f in feats]
        train = torch.is_grad_enabled() and self.training
        net = self._net(B, sizes, train)
        with torch.no_grad():
            net.repack(everything=True)
            for f, (buf, h, w, c) in zip(feats, net.stage_out[1:]):
                _to_nhwc(f, buf, B, c, h, w)
            net.forward()
            outs = [_to_nchw(net.p[l], B, 256, h, w, feats[0].device) for l, (h, w) in enumerate(net.psize)]
        if train:
            seeds = [(net.dp[l], B, 256, h, w) for l, (h, w) in enumerate(net.psize)]
            in_grads = [None] + [(net.gc[i], B, c, h, w) for i, (_, h, w, c) in enumerate(net.stage_out[1:])]
            outs = _PartFn.apply(self, net, outs, seeds, in_grads, len(inputs), *inputs,
                                 *[p for _, p in self._trainable])
        return tuple(outs)


# ====================================================================================================== head
class FCOSHead(_StoreModule):
    """FCOSHead (mmdet/models/dense_heads/fcos_head.py:14-726; tower layout anchor_free_head.py:89-139) as a
    standalone HEADS-registry module: NCHW fp32 tensors in and out, like the reference; libdslb.so underneath."""

    @staticmethod
    def parse_cfg(cfg):
        c = dict(cfg)
        c.pop("type", None)
        out = dict(num_classes=c.pop("num_classes"), in_channels=c.pop("in_channels"),
                   feat_channels=c.pop("feat_channels", 256), stacked_convs=c.pop("stacked_convs", 4),
                   strides=tuple(c.pop("strides", (4, 8, 16, 32, 64))),
                   regress_ranges=tuple(tuple(r) for r in c.pop("regress_ranges", _DEFAULT_RANGES)),
                   center_sampling=c.pop("center_sampling", False),
                   center_sample_radius=c.pop("center_sample_radius", 1.5),
                   norm_on_bbox=c.pop("norm_on_bbox", False), centerness_on_reg=c.pop("centerness_on_reg", False),
                   loss_weight=float(c.pop("loss_weight", 1.0)), soft_weight=float(c.pop("soft_weight", 0.0)),
                   soft_warm_up=c.pop("soft_warm_up", 0))
        fl = _loss_cfg(c.pop("loss_cls", None), "FocalLoss", use_sigmoid=True)
        out["gamma"], out["alpha"] = float(fl.get("gamma", 2.0)), float(fl.get("alpha", 0.25))
        _loss_cfg(c.pop("loss_bbox", dict(type="GIoULoss")), "GIoULoss")
        _loss_cfg(c.pop("loss_centerness", None), "CrossEntropyLoss", use_sigmoid=True)
        _check(out["in_channels"] == 256 and out["feat_channels"] == 256 and out["stacked_convs"] == 4,
               "FCOSHead tower must be 4 stacked 256-channel convs (TMA tiles are 64 channels wide)")
        _check(out["centerness_on_reg"], "centerness_on_reg=False")
        _check(not c.pop("dcn_on_last_conv", False), "dcn_on_last_conv")
        _check(c.pop("conv_bias", "auto") in (True, "auto"), "conv_bias=False")
        nc = dict(c.pop("norm_cfg", dict(type="GN", num_groups=32, requires_grad=True)))
        _check(nc.get("type") == "GN" and nc.get("num_groups") == 32, "norm_cfg must be GN(32)")
        _check(len(out["strides"]) == len(out["regress_ranges"]) == 5, "five levels")
        _check(out["num_classes"] % 4 == 0, "num_classes must be a multiple of 4 (COCO 80, VOC 20: float4 rows of logits)")
        out["train_cfg"], out["test_cfg"] = c.pop("train_cfg", None), dict(c.pop("test_cfg", None) or {})
        c.pop("init_cfg", None)
        c.pop("conv_cfg", None)
        if c:
            raise TypeError(f"FCOSHead got unexpected keyword arguments {sorted(c)}")
        return out

    def __init__(self, num_classes, in_channels, **kwargs):
        super().__init__()
        self.cfg = self.parse_cfg(dict(num_classes=num_classes, in_channels=in_channels, **kwargs))
        for k, v in self.cfg.items():
            setattr(self, k, v)
        self.cls_out_channels = num_classes
        self.cur_iter = 0
        store = ParamStore(head_spec(num_classes), "cpu").init_reference(0)
        self._bind_store(store, prefix_strip="bbox_head.")
        self._trainable = self.trainable_parameters()
        self._nets = OrderedDict()
        self._posts = OrderedDict()
        self.check_grad_outputs = False
        self.fp16_enabled = False

    def init_weights(self):
        self.store.init_reference(0)

    def _on_store_moved(self):
        self._nets.clear()
        self._posts.clear()

    def load_state_dict(self, sd, strict=True):
        return self.store.load_state_dict({"bbox_head." + k: v for k, v in sd.items()}, strict=strict)

    def _net(self, B, sizes, train):
        from .engine import FCOSNet
        key = (B, tuple(sizes), bool(train))
        if key not in self._nets:
            if len(self._nets) >= 4:
                self._nets.popitem(last=False)
            c = self.cfg
            self._nets[key] = FCOSNet(B, 0, 0, num_classes=c["num_classes"], train=train, store=self.store,
                                      device=self.store.device, loss_weight=c["loss_weight"],
                                      soft_weight=c["soft_weight"], center_sampling=c["center_sampling"],
                                      radius=c["center_sample_radius"], norm_on_bbox=c["norm_on_bbox"], parts="head",
                                      level_sizes=sizes, strides=c["strides"], regress_ranges=c["regress_ranges"],
                                      parity_outputs=True)
        else:
            self._nets.move_to_end(key)
        return self._nets[key]

    def forward(self, feats):
        """fcos_head.py:118-168: tuple of 5 (B, 256, h, w) maps -> (cls_scores, bbox_preds, centernesses), lists of
        (B, C, h, w) / (B, 4, h, w) / (B, 1, h, w) fp32. bbox_pred is x stride in eval mode only (:161-166)."""
        from . import _lib as L
        assert len(feats) == 5
        _need_cuda(feats[0], "FCOSHead.forward")
        if torch.is_grad_enabled() and any(f.requires_grad for f in feats):
            # a detector that composes this head with its own trainable backbone would get grad-less outputs: the
            # gradient of the fused head comes out of the loss kernel (FCOS.forward_train), not of these tensors
            raise NotImplementedError("dsl_b200 FCOSHead.forward does not record autograd history: train through "
                                      "dsl_b200.plugin.FCOS (fused loss + backward) or call it under torch.no_grad()")
        B = feats[0].shape[0]
        sizes = [tuple(f.shape[-2:]) for f in feats]
        net = self._net(B, sizes, self.training)
        s = L.cur_stream()
        with torch.no_grad():
            net.repack(everything=True)
            for l, f in enumerate(feats):
                assert f.shape[1] == 256
                L.check(L.lib.dslb_nchw_to_nhwc_bf16(L.ptr(f.contiguous().float()), L.ptr(net.p[l]), B, 256, sizes[l][0],
                                                     sizes[l][1], 256, s), "nchw_to_nhwc")
            net.forward_head()
            cls, box, ctr = [], [], []
            C = self.cfg["num_classes"]
            for l, (h, w) in enumerate(sizes):
                c = torch.empty(B, C, h, w, device=feats[0].device)
                b = torch.empty(B, 4, h, w, device=feats[0].device)
                t = torch.empty(B, 1, h, w, device=feats[0].device)
                L.check(L.lib.dslb_nhwc_to_nchw_f32(L.ptr(net.cls_out[l]), L.ptr(c), B, C, h, w, C, 1, s), "to_nchw")
                L.check(L.lib.dslb_nhwc_to_nchw_f32(L.ptr(net.rc_out[l]), L.ptr(b), B, 4, h, w, 8, 1, s), "to_nchw")
                L.check(L.lib.dslb_nhwc_to_nchw_f32(L.ptr(net.rc_out[l][..., 4:]), L.ptr(t), B, 1, h, w, 8, 1, s),
                        "to_nchw")
                cls.append(c)
                box.append(b)
                ctr.append(t)
        self._last = (net, sizes)
        return cls, box, ctr

    def _load_outputs(self, net, cls_scores, bbox_preds, centernesses):
        for l in range(5):
            net.cls_out[l].copy_(cls_scores[l].permute(0, 2, 3, 1))
            net.rc_out[l][..., :4].copy_(bbox_preds[l].permute(0, 2, 3, 1))
            net.rc_out[l][..., 4].copy_(centernesses[l][:, 0])

    def get_targets(self, points, gt_bboxes_list, gt_labels_list):
        """fcos_head.py:562-621: (labels per level, bbox_targets per level), concatenated over images, bit-exact.
        `points` (list of (h*w, 2) per level, from get_points) only fixes the level sizes."""
        net, sizes = self._last
        assert len(points) == len(sizes) and all(p.shape[0] == h * w for p, (h, w) in zip(points, sizes))
        net.set_targets(gt_bboxes_list, gt_labels_list, None)
        net.run_targets()
        n = [net.B * h * w for (h, w) in sizes]
        return list(net.labels.split(n)), list(net.bbox_targets.split(n))

    def get_points(self, featmap_sizes, dtype=torch.float32, device="cuda", flatten=True):
        """anchor_free_head.py:287-321 + fcos_head.py:550-560: point = index * stride + stride // 2."""
        out = []
        for (h, w), s in zip(featmap_sizes, self.cfg["strides"]):
            y, x = torch.meshgrid(torch.arange(h, device=device, dtype=dtype),
                                  torch.arange(w, device=device, dtype=dtype), indexing="ij")
            out.append(torch.stack((x.reshape(-1) * s, y.reshape(-1) * s), dim=-1) + s // 2)
        return out

    def loss(self, cls_scores, bbox_preds, centernesses, gt_bboxes, gt_labels, img_metas, gt_bboxes_ignore=None):
        """fcos_head.py:170-338 -> dict(loss_cls, loss_bbox, loss_centerness[, loss_sisoft]) of 0-dim fp32 tensors.
        (Gradients flow to parameters through FCOS.forward_train; this standalone entry point evaluates the losses on
        the tensors it is given, like the reference's unit test does.)"""
        assert len(cls_scores) == len(bbox_preds) == len(centernesses)
        _need_cuda(cls_scores[0], "FCOSHead.loss")
        c = self.cfg
        if c["loss_weight"] != 1.0 and gt_bboxes_ignore is None:
            raise TypeError("object of type 'NoneType' has no len()")   # fcos_head.py:224
        B = cls_scores[0].shape[0]
        sizes = [tuple(t.shape[-2:]) for t in cls_scores]
        net = self._net(B, sizes, True)
        self._last = (net, sizes)
        with torch.no_grad():
            self._load_outputs(net, cls_scores, bbox_preds, centernesses)
            net.si_weight = 0.0
            if B % 2 == 1 and c["soft_weight"] != 0.0:
                if c["soft_warm_up"] >= self.cur_iter:
                    net.si_weight = c["soft_weight"] / 1000.0
                    self.cur_iter += 1
                else:
                    net.si_weight = c["soft_weight"]
            net.set_targets(gt_bboxes, gt_labels, gt_bboxes_ignore)
            net.run_targets()
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                net.world_size = float(dist.get_world_size())
                dist.all_reduce(net.counts)
            net.run_loss()
            return OrderedDict((k, v.clone()) for k, v in net.losses().items())

    @torch.no_grad()
    def get_bboxes(self, cls_scores, bbox_preds, centernesses, img_metas, cfg=None, rescale=False, with_nms=True):
        """fcos_head.py:340-548: [(dets (n,5), labels (n,))] per image (bbox_preds already x stride, eval mode)."""
        from .postprocess import TeacherPost
        assert len(cls_scores) == len(bbox_preds) == len(centernesses)
        _need_cuda(cls_scores[0], "FCOSHead.get_bboxes")
        _check(with_nms, "get_bboxes(with_nms=False)")
        B = cls_scores[0].shape[0]
        sizes = [tuple(t.shape[-2:]) for t in cls_scores]
        net = self._net(B, sizes, False)
        self._load_outputs(net, cls_scores, bbox_preds, centernesses)
        tc = dict(cfg or self.cfg["test_cfg"])
        key = (B, tuple(sizes))
        if key not in self._posts:
            self._posts[key] = TeacherPost(B, sizes, self.cfg["strides"], self.cfg["num_classes"], self.store.device,
                                           nms_pre=tc.get("nms_pre", 1000), score_thr=tc.get("score_thr", 0.05),
                                           iou_thr=dict(tc.get("nms", {})).get("iou_threshold", 0.5),
                                           max_per_img=tc.get("max_per_img", 100))
        post = self._posts[key]
        post.set_meta([m["img_shape"] for m in img_metas],
                      [np.asarray(m["scale_factor"]).reshape(-1).tolist() for m in img_metas] if rescale else None)
        post.decode(net.cls_out, net.rc_out)
        post.nms()
        return post.results()


# ====================================================================================================== EMA hook
def ema_update_(teacher, student, keep_rate):
    """SemiEpochBasedRunner.EMA body (semi_epoch_based_runner.py:392-406): teacher = student*(1-k) + teacher*k over
    every state_dict entry, in place, ONE launch over the flat buffers (bit-exact with the reference expression)."""
    from . import _lib as L
    ts, ss = teacher.store, student.store
    assert ts.numel == ss.numel, "teacher and student must be built from the same config"
    _need_cuda(ts.flat, "ema_update_")
    c_s = float(torch.tensor(1 - keep_rate, dtype=torch.float32))
    c_t = float(torch.tensor(keep_rate, dtype=torch.float32))
    L.check(L.lib.dslb_ema_update(L.ptr(ts.flat), L.ptr(ss.flat), ss.numel, c_s, c_t, L.cur_stream()), "ema")
    if hasattr(teacher, "_dirty"):
        teacher._dirty()


from .hooks import EMAOWNHook, _unwrap  # noqa: E402,F401  (the reference's hook; an mmcv Hook when mmcv is importable)


def scale_invariant_input(img, gt_bboxes, gt_labels, gt_bboxes_ignore, img_metas):
    """SemiEpochBasedRunner.train SI block (semi_epoch_based_runner.py:186-204): append a bilinear half-resolution copy
    of the LAST image (zero padded to the batch H x W), with its boxes, ignore boxes and meta sizes halved.
    Returns (img, gt_bboxes, gt_labels, gt_bboxes_ignore, img_metas) of the batch of B + 1."""
    from . import _lib as L
    _need_cuda(img, "scale_invariant_input(img)")
    h, w = img.shape[-2:]
    src = img[-1].contiguous().float()
    tmp = torch.empty_like(src)
    L.check(L.lib.dslb_si_half_image(L.ptr(src), L.ptr(tmp), int(src.shape[0]), int(h), int(w), L.cur_stream()),
            "si_half_image")
    tmp = tmp.to(img.dtype)[None]
    meta = dict(img_metas[-1])
    for k in ("img_shape", "pad_shape"):
        if k in meta:
            meta[k] = (int(meta[k][0] / 2), int(meta[k][1] / 2), meta[k][2])
    if "scale_factor" in meta:
        meta["scale_factor"] = np.asarray(meta["scale_factor"]) / 2
    ig = gt_bboxes_ignore[-1].clone()
    if len(ig) > 0:
        ig = ig / 2
    return (torch.cat((img, tmp), 0), list(gt_bboxes) + [gt_bboxes[-1].clone() / 2],
            list(gt_labels) + [gt_labels[-1].clone()], list(gt_bboxes_ignore) + [ig], list(img_metas) + [meta])


# ====================================================================================================== registry
class RLA_ResNet(ResNet):
    """BACKBONES['RLA_ResNet'] (mmdet/models/backbones/resnet_rla.py:140-400), the backbone of the shipped DSL configs
    (configs/fcos_semi/RLA_*.py:3-13): same constructor keywords; supported = what those configs use (Bottleneck blocks,
    layers (3,4,6,3) or (3,4,23,3), rla_channel 32, no SE / ECA, frozen_stages=1, norm_eval=True). NCHW fp32 in, the four
    stage outputs NCHW fp32 out (:312-313), autograd-connected for stages 2-4 incl. the trainable BatchNorm affines."""

    @staticmethod
    def check_cfg(cfg):
        """Validate an RLA_ResNet config dict / kwargs; returns the equivalent ResNet depth (block counts)."""
        layers = tuple(cfg.get("layers", (3, 4, 6, 3)))
        depth = {v: k for k, v in RESNET_BLOCKS.items()}.get(layers)
        _check(depth is not None, f"RLA_ResNet layers {layers}")
        _check(cfg.get("block") is None and cfg.get("rla_channel", 32) == 32 and not cfg.get("SE", False)
               and cfg.get("ECA") is None, "only RLA_Bottleneck, rla_channel=32, no SE / ECA")
        _check(cfg.get("groups", 1) == 1 and cfg.get("width_per_group", 64) == 64 and
               not any(cfg.get("replace_stride_with_dilation") or ()), "groups / width_per_group / dilation")
        _check(cfg.get("norm_eval", True) and cfg.get("norm_layer") is None, "norm_eval=True with nn.BatchNorm2d")
        _check(cfg.get("style", "pytorch") == "pytorch", "RLA_ResNet is PyTorch style (stride on conv2)")
        return depth

    def __init__(self, block=None, layers=(3, 4, 6, 3), num_classes=1000, rla_channel=32, SE=False, ECA=None,
                 frozen_stages=-1, norm_eval=True, style="pytorch", zero_init_last_bn=True, groups=1, width_per_group=64,
                 replace_stride_with_dilation=None, norm_layer=None, pretrained=None):
        _StoreModule.__init__(self)
        self.depth = self.check_cfg(dict(block=block, layers=layers, rla_channel=rla_channel, SE=SE, ECA=ECA,
                                         norm_eval=norm_eval, style=style, groups=groups,
                                         width_per_group=width_per_group, norm_layer=norm_layer,
                                         replace_stride_with_dilation=replace_stride_with_dilation))
        _check(frozen_stages == 1, "frozen_stages must be 1")
        self.backbone_kind = "rla"
        self.out_indices = (0, 1, 2, 3)
        self.pretrained = pretrained
        self._bind_store(ParamStore(rla_resnet_spec(tuple(layers), prefix=""), "cpu").init_reference(0))
        self._trainable = self.trainable_parameters()
        self._nets = OrderedDict()

    def init_weights(self):
        """resnet_rla.py:379-388: load `pretrained` (non-strict) when it names a checkpoint file."""
        import os
        if isinstance(self.pretrained, str) and os.path.isfile(self.pretrained):
            ck = torch.load(self.pretrained, map_location="cpu")
            self.load_state_dict(ck.get("state_dict", ck), strict=False)
        else:
            ResNet.init_weights(self)


def register(force=True, runner=False):
    """Register under the reference's registry keys. Returns the list of keys registered ([] when mmdet / mmcv are not
    importable, e.g. on a bare GPU box: the classes are then used directly).

    Importing dsl_b200.plugin swaps the MODEL-side classes (detector / backbone / neck / head / losses) and the
    reference's own EMAOWNHook; the fused RUNNERS['SemiEpochBasedRunner'] changes how a whole iteration executes, so it
    is registered only on request: `custom_imports=dict(imports=['dsl_b200.plugin', 'dsl_b200.plugin_runner'])` or
    register(runner=True)."""
    done = []
    try:
        from mmdet.models.builder import DETECTORS, HEADS
    except Exception:
        return done
    DETECTORS.register_module(name="FCOS", force=force, module=FCOS)
    HEADS.register_module(name="FCOSHead", force=force, module=FCOSHead)
    done += ["DETECTORS.FCOS", "HEADS.FCOSHead"]
    try:
        from mmdet.models.builder import BACKBONES, NECKS
        BACKBONES.register_module(name="ResNet", force=force, module=ResNet)
        BACKBONES.register_module(name="RLA_ResNet", force=force, module=RLA_ResNet)
        NECKS.register_module(name="FPN", force=force, module=FPN)
        done += ["BACKBONES.ResNet", "BACKBONES.RLA_ResNet", "NECKS.FPN"]
    except Exception:
        pass
    try:
        from mmcv.runner import HOOKS
        HOOKS.register_module(name="EMAOWNHook", force=force, module=EMAOWNHook)
        done.append("HOOKS.EMAOWNHook")
    except Exception:
        pass
    try:   # LOSSES.{FocalLoss, GIoULoss, CrossEntropyLoss} (need libdslb.so)
        from . import losses
        done += losses.register(force)
    except Exception:
        pass
    if runner:
        from . import runner as _runner
        done += _runner.register(force)
    return done


REGISTERED = register()
