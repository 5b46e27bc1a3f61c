"""TEST INFRASTRUCTURE ONLY — a minimal stand-in for the un-vendored third-party dependency `mmcv-full`
(pinned >=1.3.8,<=1.4.0 by /root/reference/mmdet/__init__.py:19-27; README says 1.3.10), just enough for the
reference's OWN hot-path source files to execute in place (see oracle/ref_loader.py).

Only thin torch.nn wrappers and no-op decorators are restated here; no detector arithmetic lives in this file.
What each piece stands in for (mmcv 1.3.x public behaviour):
  Registry / build_from_cfg  -- `type`-keyed class registry, `register_module(force=...)`, `build(cfg)`
  ConvModule                 -- conv -> norm -> activation, bias='auto' means bias = (norm_cfg is None)
  Scale                      -- learnable scalar multiply
  build_conv_layer / build_norm_layer -- nn.Conv2d / (name, nn.BatchNorm2d|nn.GroupNorm); requires_grad flag
  force_fp32 / auto_fp16 / jit -- identity decorators (fp16_enabled is False on this path)
  ops.nms / ops.batched_nms  -- torchvision NMS (IoU > thr suppressed, offset 0), class-offset trick
  runner.Hook / HOOKS / RUNNERS / build_runner / build_optimizer / OptimizerHook / EpochBasedRunner,
  parallel.MMDataParallel    -- shells so that mmdet/apis/train.py::train_detector executes in place (no training logic)
  ops.sigmoid_focal_loss     -- deliberately NOT provided: on CPU the reference takes its own
                                py_sigmoid_focal_loss branch (mmdet/models/losses/focal_loss.py:162-168)
"""
import inspect
import sys
import types

import torch
import torch.nn as nn


class Registry:
    def __init__(self, name, build_func=None, parent=None, scope=None):
        self._name = name
        self._module_dict = {}
        self.parent = parent
        self.build_func = build_func or build_from_cfg
        if parent is not None and build_func is None:
            self.build_func = parent.build_func

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        if key in self._module_dict:
            return self._module_dict[key]
        if self.parent is not None:
            return self.parent.get(key)
        return None

    def _register(self, cls, name=None, force=False):
        name = name or cls.__name__
        if not force and name in self._module_dict:
            raise KeyError(f"{name} is already registered in {self._name}")
        self._module_dict[name] = cls

    def register_module(self, name=None, force=False, module=None):
        if module is not None:
            self._register(module, name, force)
            return module

        def deco(cls):
            self._register(cls, name, force)
            return cls

        return deco

    def build(self, *args, **kwargs):
        return self.build_func(*args, **kwargs, registry=self)


def build_from_cfg(cfg, registry, default_args=None):
    args = dict(cfg)
    if default_args:
        for k, v in default_args.items():
            args.setdefault(k, v)
    t = args.pop("type")
    cls = registry.get(t) if isinstance(t, str) else t
    if cls is None:
        raise KeyError(f"{t} is not in the {registry._name} registry")
    return cls(**args)


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg
        self._is_init = False

    def init_weights(self):
        pass


class Sequential(BaseModule, nn.Sequential):
    def __init__(self, *args, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.Sequential.__init__(self, *args)


class ModuleList(BaseModule, nn.ModuleList):
    def __init__(self, modules=None, init_cfg=None):
        BaseModule.__init__(self, init_cfg)
        nn.ModuleList.__init__(self, modules)


class Scale(nn.Module):
    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

    def forward(self, x):
        return x * self.scale


def build_conv_layer(cfg, *args, **kwargs):
    assert cfg is None or cfg.get("type", "Conv2d") in ("Conv2d", "Conv"), cfg
    return nn.Conv2d(*args, **kwargs)


def build_norm_layer(cfg, num_features, postfix=""):
    cfg = dict(cfg)
    t = cfg.pop("type")
    requires_grad = cfg.pop("requires_grad", True)
    cfg.setdefault("eps", 1e-5)
    if t == "BN":
        name, layer = "bn" + str(postfix), nn.BatchNorm2d(num_features, **cfg)
    elif t == "GN":
        name, layer = "gn" + str(postfix), nn.GroupNorm(num_channels=num_features, **cfg)
    else:
        raise KeyError(t)
    for p in layer.parameters():
        p.requires_grad = requires_grad
    return name, layer


def build_plugin_layer(*a, **k):
    raise NotImplementedError("plugins are not on the DSL path")


class ConvModule(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias="auto", conv_cfg=None, norm_cfg=None, act_cfg=dict(type="ReLU"), inplace=True,
                 with_spectral_norm=False, padding_mode="zeros", order=("conv", "norm", "act")):
        super().__init__()
        assert order == ("conv", "norm", "act")
        self.with_norm = norm_cfg is not None
        self.with_activation = act_cfg is not None
        if bias == "auto":
            bias = not self.with_norm
        self.conv = build_conv_layer(conv_cfg, in_channels, out_channels, kernel_size, stride=stride,
                                     padding=padding, dilation=dilation, groups=groups, bias=bias)
        if self.with_norm:
            self.norm_name, norm = build_norm_layer(norm_cfg, out_channels)
            self.add_module(self.norm_name, norm)
        if self.with_activation:
            assert act_cfg["type"] == "ReLU"
            self.activate = nn.ReLU(inplace=inplace)

    @property
    def norm(self):
        return getattr(self, self.norm_name) if self.with_norm else None

    def forward(self, x, activate=True, norm=True):
        x = self.conv(x)
        if norm and self.with_norm:
            x = self.norm(x)
        if activate and self.with_activation:
            x = self.activate(x)
        return x


def _identity_decorator(*dargs, **dkwargs):
    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]

    def deco(fn):
        return fn

    return deco


def batched_nms(boxes, scores, idxs, nms_cfg, class_agnostic=False):
    from torchvision.ops import nms as tv_nms
    cfg = dict(nms_cfg)
    assert cfg.pop("type", "nms") == "nms"
    thr = cfg.pop("iou_threshold")
    if class_agnostic:
        b = boxes
    else:
        max_coordinate = boxes.max()
        offsets = idxs.to(boxes) * (max_coordinate + torch.tensor(1).to(boxes))
        b = boxes + offsets[:, None]
    keep = tv_nms(b, scores, thr)
    return torch.cat([boxes[keep], scores[keep, None]], -1), keep


def nms(boxes, scores, iou_threshold, offset=0, score_threshold=0, max_num=-1):
    from torchvision.ops import nms as tv_nms
    is_np = not isinstance(boxes, torch.Tensor)
    if is_np:
        boxes, scores = torch.from_numpy(boxes), torch.from_numpy(scores)
    if score_threshold > 0:
        valid = scores > score_threshold
        inds0 = valid.nonzero().squeeze(1)
        boxes, scores = boxes[valid], scores[valid]
    keep = tv_nms(boxes, scores, iou_threshold)
    if max_num > 0:
        keep = keep[:max_num]
    dets = torch.cat([boxes[keep], scores[keep, None]], -1)
    if score_threshold > 0:
        keep = inds0[keep]
    if is_np:
        dets, keep = dets.numpy(), keep.numpy()
    return dets, keep


def install():
    """Put the stub package tree into sys.modules as `mmcv` (idempotent)."""
    if "mmcv" in sys.modules and getattr(sys.modules["mmcv"], "__dslb_stub__", False):
        return sys.modules["mmcv"]
    assert "mmcv" not in sys.modules, "a real mmcv is already imported"

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mmcv = mod("mmcv", __dslb_stub__=True, __version__="1.3.10", jit=_identity_decorator,
               is_tuple_of=lambda seq, t: isinstance(seq, tuple) and all(isinstance(s, t) for s in seq),
               is_list_of=lambda seq, t: isinstance(seq, list) and all(isinstance(s, t) for s in seq),
               is_str=lambda x: isinstance(x, str), build_from_cfg=build_from_cfg)
    MODELS = Registry("model")
    mmcv.cnn = mod("mmcv.cnn", MODELS=MODELS, ConvModule=ConvModule, Scale=Scale,
                   build_conv_layer=build_conv_layer, build_norm_layer=build_norm_layer,
                   build_plugin_layer=build_plugin_layer)
    mmcv.utils = mod("mmcv.utils", Registry=Registry, build_from_cfg=build_from_cfg)

    def _no_checkpoint(*a, **k):   # resnet_rla.py imports these; the oracle never loads a checkpoint file
        raise RuntimeError("checkpoint loading is not part of the oracle")

    # ---- runner-side shells (mmcv/runner/*): just enough for mmdet/apis/train.py::train_detector to execute in place
    # against a runner class registered under RUNNERS (tests/test_train_detector.py). No training logic lives here.
    class Hook:   # mmcv/runner/hooks/hook.py: every stage a no-op, *_train_* / *_val_* fall through to the generic ones
        def before_run(self, runner): pass
        def after_run(self, runner): pass
        def before_epoch(self, runner): pass
        def after_epoch(self, runner): pass
        def before_iter(self, runner): pass
        def after_iter(self, runner): pass
        def before_train_epoch(self, runner): self.before_epoch(runner)
        def before_val_epoch(self, runner): self.before_epoch(runner)
        def after_train_epoch(self, runner): self.after_epoch(runner)
        def after_val_epoch(self, runner): self.after_epoch(runner)
        def before_train_iter(self, runner): self.before_iter(runner)
        def before_val_iter(self, runner): self.before_iter(runner)
        def after_train_iter(self, runner): self.after_iter(runner)
        def after_val_iter(self, runner): self.after_iter(runner)
        def every_n_epochs(self, runner, n): return (runner.epoch + 1) % n == 0 if n > 0 else False
        def every_n_inner_iters(self, runner, n): return (runner.inner_iter + 1) % n == 0 if n > 0 else False
        def every_n_iters(self, runner, n): return (runner.iter + 1) % n == 0 if n > 0 else False
        def end_of_epoch(self, runner): return runner.inner_iter + 1 == len(runner.data_loader)
        def is_last_epoch(self, runner): return runner.epoch + 1 == runner._max_epochs
        def is_last_iter(self, runner): return runner.iter + 1 == runner._max_iters

    class OptimizerHook(Hook):   # also subclassed by mmdet/core/utils/dist_utils.py at import time
        def __init__(self, grad_clip=None, *a, **k):
            self.grad_clip = grad_clip

    class Fp16OptimizerHook(OptimizerHook):
        pass

    class DistSamplerSeedHook(Hook):
        def before_epoch(self, runner):
            s = getattr(runner.data_loader, "sampler", None)
            if hasattr(s, "set_epoch"):
                s.set_epoch(runner.epoch)

    class EpochBasedRunner:   # isinstance() target of train_detector (:167)
        pass

    HOOKS = Registry("hook")
    RUNNERS = Registry("runner")

    def build_runner(cfg, default_args=None):   # mmcv/runner/builder.py
        return build_from_cfg(cfg, RUNNERS, default_args=default_args)

    def build_optimizer(model, cfg):
        """mmcv DefaultOptimizerConstructor for what the fcos_semi configs use (configs/fcos_semi/*.py:179-182):
        torch.optim.<type> with one param group per parameter when paramwise_cfg is given; bias_lr_mult /
        bias_decay_mult apply to `.bias` of non-norm layers; parameters with requires_grad=False keep the defaults."""
        cfg = dict(cfg)
        paramwise = cfg.pop("paramwise_cfg", None)
        cfg.pop("constructor", None)
        cls = getattr(torch.optim, cfg.pop("type"))
        if hasattr(model, "module"):
            model = model.module
        if not paramwise:
            return cls(model.parameters(), **cfg)
        groups = []
        for name, p in model.named_parameters():
            g = {"params": [p]}
            is_norm = any(t in name for t in (".gn.", ".bn", "downsample.1.", "stage_bns."))
            if p.requires_grad and name.endswith(".bias") and not is_norm:
                g["lr"] = cfg["lr"] * paramwise.get("bias_lr_mult", 1.0)
                if cfg.get("weight_decay") is not None:
                    g["weight_decay"] = cfg["weight_decay"] * paramwise.get("bias_decay_mult", 1.0)
            groups.append(g)
        return cls(groups, **cfg)

    mmcv.runner = mod("mmcv.runner", BaseModule=BaseModule, Sequential=Sequential, ModuleList=ModuleList,
                      force_fp32=_identity_decorator, auto_fp16=_identity_decorator, OptimizerHook=OptimizerHook,
                      load_checkpoint=_no_checkpoint, load_state_dict=_no_checkpoint, Hook=Hook, HOOKS=HOOKS,
                      RUNNERS=RUNNERS, build_runner=build_runner, build_optimizer=build_optimizer,
                      EpochBasedRunner=EpochBasedRunner, Fp16OptimizerHook=Fp16OptimizerHook,
                      DistSamplerSeedHook=DistSamplerSeedHook)
    sys.modules["mmcv.runner.builder"] = mod("mmcv.runner.builder", RUNNERS=RUNNERS, build_runner=build_runner)

    class _Wrap(nn.Module):   # mmcv/parallel: MMDataParallel / MMDistributedDataParallel keep the model in .module
        def __init__(self, module, device_ids=None, **kw):
            super().__init__()
            self.module = module
            self.device_ids = device_ids

        def train_step(self, *a, **k):
            return self.module.train_step(*a, **k)

        def forward(self, *a, **k):
            return self.module(*a, **k)

    class DataContainer:
        def __init__(self, data, stack=False, padding_value=0, cpu_only=False, pad_dims=2):
            self.data, self.stack, self.cpu_only = data, stack, cpu_only

    mmcv.parallel = mod("mmcv.parallel", MMDataParallel=_Wrap, MMDistributedDataParallel=_Wrap,
                        DataContainer=DataContainer, is_module_wrapper=lambda m: isinstance(m, _Wrap))

    def _no_cuda_focal(*a, **k):
        raise RuntimeError("mmcv.ops.sigmoid_focal_loss is CUDA-only; the CPU oracle uses py_sigmoid_focal_loss")

    mmcv.ops = mod("mmcv.ops", sigmoid_focal_loss=_no_cuda_focal, batched_nms=batched_nms, nms=nms)
    sys.modules["mmcv.ops.nms"] = mod("mmcv.ops.nms", batched_nms=batched_nms, nms=nms)
    mmcv.ops.nms_mod = sys.modules["mmcv.ops.nms"]
    return mmcv


_ = inspect  # keep import (used by downstream debugging)
