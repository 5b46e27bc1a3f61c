"""SemiEpochBasedRunner on the fused engine: the reference's RUNNERS-registry runner
(mmdet/runner/hooks/semi_epoch_based_runner.py:49-511) restated over `DSLEngine`, with the part of mmcv's
BaseRunner / EpochBasedRunner interface that `mmdet/apis/train.py::train_detector` (:104-217) drives:

    build_runner(cfg.runner, default_args=dict(model, optimizer, work_dir, logger, meta, ema_model, scale_invariant))
    runner.ema_flag / .ITER / .timestamp = ...                                            (:149-152)
    runner.register_training_hooks(lr_config, optimizer_config, ema_config, checkpoint_config, log_config,
                                   momentum_config)                                         (:164-166; runner :470-511)
    runner.register_hook(hook[, priority])          (DistSamplerSeedHook, EvalHook, UnlabelPredHook, custom hooks)
    runner.resume(cfg.resume_from) / runner.load_checkpoint(cfg.load_from)                  (:214-217; runner :350-366)
    runner.run(data_loaders, cfg.workflow)

What the reference spreads over one iteration — the scale-invariant extra input (:186-204), `run_iter` ->
`model.train_step` (:142-167), mmcv's OptimizerHook (clip 35 + SGD), `EMAOWNHook` -> `runner.EMA()` (:368-409), the LR
hook's param_group writes and `UnlabelPredHook`'s teacher inference (unlabel_pred_hook.py:512-562) — is ONE captured
CUDA-graph step here:
  * the optimizer hook is not called: its `grad_clip` configures the fused clip + SGD (`register_optimizer_hook`);
  * the LR hook IS called (mmcv's own when importable, else dsl_b200.hooks.StepLrUpdaterHook): before every step
    param_groups[0]['lr'] / initial_lr is written to the device-side LR scalar the captured SGD kernels read;
  * EMAOWNHook(mode='iteration', interval=1) — the shipped ema_config — runs inside the step from `start_point` on;
    other modes reach `runner.EMA()` from the hook's stages;
  * with `enable_device_pseudo_labels(...)` the teacher's boxes of iteration i become the unlabeled images' GT of
    iteration i + lag on the device (burn-in until `start_point`, as UnlabelPredHook.after_train_iter :455-469).

Batches are dicts with the reference's keys: `img` (B,3,H,W), `img_metas`, `gt_bboxes`, `gt_labels`,
`gt_bboxes_ignore`, plus optionally `teacher_img` (the weak-aug images the EMA teacher labels; the reference's hook reads
them from disk); mmcv DataContainers (`.data[0]`) are unwrapped.

CUDA only; the model / ema_model must be `dsl_b200.plugin.FCOS` modules (their ParamStores are trained in place).
"""
import logging
import os
import os.path as osp
import re
import time
from collections import OrderedDict

import torch

from . import _lib as L
from . import hooks as H
from .trainer import DSLEngine


def _unwrap(v):
    """mmcv DataContainer -> its payload of the first (only) GPU."""
    d = getattr(v, "data", None)
    if d is not None and not isinstance(v, torch.Tensor):
        return d[0]
    return v


def _dist_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class SemiEpochBasedRunner:
    def __init__(self, model, batch_processor=None, optimizer=None, work_dir=None, logger=None, meta=None,
                 max_iters=None, max_epochs=None, ema_model=None, scale_invariant=False):
        if batch_processor is not None:
            raise NotImplementedError("dsl_b200 SemiEpochBasedRunner: batch_processor is deprecated in the reference and "
                                      "not supported; the model's fused train step is used")
        m = model.module if hasattr(model, "module") else model
        e = ema_model.module if (ema_model is not None and hasattr(ema_model, "module")) else ema_model
        from .plugin import FCOS
        if not isinstance(m, FCOS) or (e is not None and not isinstance(e, FCOS)):
            raise TypeError("dsl_b200 SemiEpochBasedRunner drives dsl_b200.plugin.FCOS models")
        if optimizer is not None and not isinstance(optimizer, torch.optim.Optimizer):
            raise TypeError(f"optimizer must be a torch.optim.Optimizer object or None, but got {type(optimizer)}")
        if logger is not None and not isinstance(logger, logging.Logger):
            raise TypeError(f"logger must be a logging.Logger object, but got {type(logger)}")
        if meta is not None and not isinstance(meta, dict):
            raise TypeError(f"meta must be a dict or None, but got {type(meta)}")
        if max_epochs is not None and max_iters is not None:
            raise ValueError("Only one of `max_epochs` or `max_iters` can be set.")
        self.model, self.ema_model = m, e
        self.optimizer = optimizer
        self.logger = logger or logging.getLogger("dsl_b200")
        self.meta = meta
        if isinstance(work_dir, str):
            self.work_dir = osp.abspath(work_dir)
            os.makedirs(self.work_dir, exist_ok=True)
        elif work_dir is None:
            self.work_dir = None
        else:
            raise TypeError('"work_dir" must be a str or None')
        self._model_name = m.__class__.__name__
        self._rank, self._world_size = _dist_info()
        self.timestamp = time.strftime("%Y%m%d_%H%M%S", time.localtime())
        # reference: train_detector sets ema_flag = False (apis/train.py:149) and runner.EMA() raises it on its first
        # call (:408); the teacher is used by the hooks only afterwards. Here it tells whether an EMA has been applied.
        self.ema_flag = False
        self.ITER = None
        self.scale_invariant = bool(scale_invariant)
        self.mode = None
        self._hooks = []
        self._epoch = self._iter = self._inner_iter = 0
        self._max_epochs, self._max_iters = max_epochs, max_iters
        self.outputs = None
        self.log_buffer = H.LogBuffer()
        self.imagefiles = []
        self.data_loader = None
        self.engine = None                   # the engine of the current batch shape
        # multi-scale training (BASELINE configs[4]: Resize [(1333,640),(1333,800)] + Pad(32) gives a handful of padded
        # shapes): one engine (plan + CUDA graph) per (B, H, W), least recently used first; all of them train the SAME
        # parameter stores and share the optimizer / adathres state (see _share_state)
        self._engines = OrderedDict()
        self.max_cached_shapes = 8
        self._si_iter = 0                    # FCOSHead.cur_iter of the SI-soft warm-up, across shapes
        # fused EMA: cfg ema_config (configs/fcos_semi/*.py:199: interval=1, mode="iteration", ratio=0.99,
        # start_point=1). None = no EMA inside the step (no ema hook registered, or an epoch-mode hook calls EMA()).
        self.fused_ema = dict(ratio=0.99, start_point=1) if e is not None else None
        self.max_grad_norm = 35.0            # cfg optimizer_config.grad_clip (:183-185) until register_optimizer_hook
        self._lr_scale_host = 1.0
        # device-side pseudo-label loop (enable_device_pseudo_labels)
        self.pl_start_iter = None
        self.pl_lag = 1
        self._pl_ring = []
        self._pl_steps = 0

    # ---- counters (mmcv BaseRunner properties) ----------------------------------------------------------------
    epoch = property(lambda self: self._epoch)
    iter = property(lambda self: self._iter)
    inner_iter = property(lambda self: self._inner_iter)
    max_epochs = property(lambda self: self._max_epochs)
    max_iters = property(lambda self: self._max_iters)
    rank = property(lambda self: self._rank)
    world_size = property(lambda self: self._world_size)
    model_name = property(lambda self: self._model_name)
    hooks = property(lambda self: self._hooks)

    # ---- hooks (mmcv/runner/base_runner.py: register_hook keeps the list sorted by priority, stable) --------------
    def register_hook(self, hook, priority="NORMAL"):
        if not H.is_hook(hook) and not any(hasattr(hook, s) for s in H._Hook.stages):
            raise TypeError(f"hook must be an mmcv.runner.Hook (or expose its stage methods), got {type(hook)}")
        if hasattr(hook, "priority"):
            raise ValueError('"priority" is a reserved attribute for hooks')
        hook.priority = H.get_priority(priority)
        if isinstance(hook, H.EMAOWNHook):
            self._configure_ema(hook)
        inserted = False
        for i in range(len(self._hooks) - 1, -1, -1):
            if hook.priority >= self._hooks[i].priority:
                self._hooks.insert(i + 1, hook)
                inserted = True
                break
        if not inserted:
            self._hooks.insert(0, hook)

    def call_hook(self, fn_name):
        for h in self._hooks:
            fn = getattr(h, fn_name, None)
            if fn is not None:
                fn(self)

    def get_hook_info(self):
        lines = []
        for stage in H._Hook.stages:
            names = [f"({h.priority:<3}) {type(h).__name__}" for h in self._hooks if hasattr(h, stage)]
            if names:
                lines.append(f"{stage}:\n" + "\n".join(names) + "\n -------------------- ")
        return "\n".join(lines)

    def _configure_ema(self, hook):
        """EMAOWNHook(mode='iteration', interval=1): the EMA runs inside the captured step (from start_point on); any
        other trigger rule: not inside the step, the hook's stages call runner.EMA()."""
        if self.ema_model is None:      # the reference's runner.EMA would fail on a missing teacher; here: nothing to do
            hook.fused, self.fused_ema = False, None
            return
        if hook.mode == "iteration" and hook.interval == 1:
            hook.fused = True
            self.fused_ema = dict(ratio=float(hook.ratio), start_point=int(hook.start_point))
        else:
            hook.fused = False
            self.fused_ema = None
        for eng in self._engines.values():
            self._apply_ema_cfg(eng)

    def _apply_ema_cfg(self, eng):
        want = self.fused_ema is not None and self.fused_ema["start_point"] <= self._iter + 1
        keep = self.fused_ema["ratio"] if self.fused_ema is not None else eng.ema_keep
        if eng.ema_in_step != want or (want and eng.ema_keep != keep):
            eng.ema_in_step, eng.ema_keep = want, keep
            eng.graphs = None            # the EMA launch and its coefficients are constants of the captured step

    def set_ema_ratio(self, ratio):
        """EMAOWNHook.step_decay changed the ratio (ema.py:24-26)."""
        if self.fused_ema is not None:
            self.fused_ema["ratio"] = float(ratio)

    def register_lr_hook(self, lr_config):
        """mmcv BaseRunner.register_lr_hook: policy 'step' -> StepLrUpdaterHook, priority VERY_HIGH."""
        if lr_config is None:
            return
        if isinstance(lr_config, dict):
            assert "policy" in lr_config
            cfg = dict(lr_config)
            policy = cfg.pop("policy")
            if policy == policy.lower():
                policy = policy.title()
            cfg["type"] = policy + "LrUpdaterHook"
            hook = H.build_hook(cfg)
        else:
            hook = lr_config
        self.register_hook(hook, priority="VERY_HIGH")

    def register_momentum_hook(self, momentum_config):
        if momentum_config is None:
            return
        raise NotImplementedError("dsl_b200 runner: momentum schedules (momentum is a launch constant of the fused SGD; "
                                  "the fcos_semi configs set none)")

    def register_optimizer_hook(self, optimizer_config):
        """The fused step IS the optimizer hook (backward, clip_grad_norm_, SGD): only grad_clip is taken from the
        config / hook object, nothing is registered (mmcv would call loss.backward() + optimizer.step())."""
        if optimizer_config is None:
            return
        self.max_grad_norm = H.grad_clip_of(optimizer_config)
        for eng in self._engines.values():
            if eng.max_grad_norm != self.max_grad_norm:
                eng.max_grad_norm, eng.graphs = self.max_grad_norm, None

    def register_ema_hook(self, ema_config):
        """semi_epoch_based_runner.py:460-468."""
        if ema_config is None:
            return
        hook = H.build_hook(ema_config, default_type="EMAOWNHook") if isinstance(ema_config, dict) else ema_config
        self.register_hook(hook, priority=45)

    def register_checkpoint_hook(self, checkpoint_config):
        if checkpoint_config is None:
            return
        hook = H.build_hook(checkpoint_config, default_type="CheckpointHook") if isinstance(checkpoint_config, dict) \
            else checkpoint_config
        self.register_hook(hook, priority="NORMAL")

    def register_timer_hook(self, timer_config):
        if timer_config is None:
            return
        hook = H.build_hook(timer_config) if isinstance(timer_config, dict) else timer_config
        self.register_hook(hook, priority="LOW")

    def register_logger_hooks(self, log_config):
        if log_config is None:
            return
        interval = log_config["interval"]
        for info in log_config["hooks"]:
            hook = H.build_hook(dict(info, interval=interval))
            self.register_hook(hook, priority="VERY_LOW")

    def register_custom_hooks(self, custom_config):
        if custom_config is None:
            return
        for item in (custom_config if isinstance(custom_config, list) else [custom_config]):
            if isinstance(item, dict):
                item = dict(item)
                priority = item.pop("priority", "NORMAL")
                self.register_hook(H.build_hook(item), priority=priority)
            else:
                self.register_hook(item, priority="NORMAL")

    def register_training_hooks(self, lr_config, optimizer_config=None, ema_config=None, checkpoint_config=None,
                                log_config=None, momentum_config=None, timer_config=dict(type="IterTimerHook"),
                                custom_hooks_config=None):
        """semi_epoch_based_runner.py:470-511 (same positional order: train_detector passes six positionals)."""
        self.register_lr_hook(lr_config)
        self.register_momentum_hook(momentum_config)
        self.register_optimizer_hook(optimizer_config)
        if ema_config is not None:
            self.register_ema_hook(ema_config)
        self.register_checkpoint_hook(checkpoint_config)
        self.register_timer_hook(timer_config)
        self.register_logger_hooks(log_config)
        self.register_custom_hooks(custom_hooks_config)

    # ---- device-side pseudo-label loop ---------------------------------------------------------------------------
    def enable_device_pseudo_labels(self, start_point=0, lag=1, num_unlabeled=None):
        """Close the teacher -> student loop on the device (replaces UnlabelPredHook's JSON round trip,
        unlabel_pred_hook.py:455-469, 512-562 + semicoco.py:220-269): from iteration `start_point * iters_per_epoch` on
        (the hook's burn-in rule), the last `num_unlabeled` images of a batch (default: the teacher batch) take the
        pseudo GT / ignore boxes the EMA teacher produced `lag` iterations earlier for `teacher_img` — mapped into the
        strong view by the batch's `img_metas` (geometry.view_from_meta) — instead of the dataloader's boxes. The batch
        must then carry `teacher_img` = the weak views of the images the student meets `lag` iterations later
        (the reference's hook runs `preload` + 1 iterations ahead of the dataloader for the same reason). Until then,
        and for the first `lag` iterations after it, the dataloader's boxes are used."""
        self.pl_start_point = float(start_point)
        self.pl_start_iter = None   # resolved at the first train() call (needs len(data_loader))
        self.pl_lag = int(lag)
        assert self.pl_lag >= 1
        self.pl_num_unlabeled = num_unlabeled
        self.pl_enabled = True

    def _pl_push(self, eng):
        """Keep the teacher's lists of this step for the step `lag` iterations ahead (lag 1: the engine's own buffers are
        read before the next step overwrites them, no copy)."""
        self._pl_steps += 1
        if self.pl_lag == 1:
            return
        if not self._pl_ring:
            names = ("pl_gt_boxes", "pl_gt_labels", "pl_gt_off", "pl_ig_boxes", "pl_ig_off")
            self._pl_ring = [tuple(torch.zeros_like(getattr(eng, n)) for n in names) for _ in range(self.pl_lag)]
        slot = self._pl_ring[(self._pl_steps - 1) % self.pl_lag]
        for dst, n in zip(slot, ("pl_gt_boxes", "pl_gt_labels", "pl_gt_off", "pl_ig_boxes", "pl_ig_off")):
            dst.copy_(getattr(eng, n), non_blocking=True)

    def _pl_source(self, eng):
        if self.pl_lag == 1:
            return None     # the engine's own pl_* buffers (shared by every shape's engine)
        return self._pl_ring[(self._pl_steps - self.pl_lag) % self.pl_lag]

    # ---- engine ------------------------------------------------------------------------------------------------
    @staticmethod
    def _share_state(src, dst):
        """Everything a step carries over to the next one besides the weights must not depend on the batch shape: the
        SGD momentum buffer, the LR-schedule scalar, the epoch's adaptive-threshold statistics / thresholds and the
        teacher's latest pseudo-label lists are ONE set of tensors used by every shape's engine (assigned before the
        new engine captures its graph)."""
        dst.mom, dst.lr_scale = src.mom, src.lr_scale
        for name in ("stat_cnt", "stat_cum", "stat_prev", "thr_class", "class_weight", "have_prev", "cand_overflow"):
            setattr(dst.post, name, getattr(src.post, name))
        if dst.teacher.B == src.teacher.B:
            for name in ("pl_gt_boxes", "pl_gt_labels", "pl_gt_off", "pl_ig_boxes", "pl_ig_off"):
                setattr(dst, name, getattr(src, name))

    def _optimizer_hparams(self):
        """cfg optimizer (:182): SGD lr .01 momentum .9 wd 1e-4, paramwise bias_lr_mult 2 / bias_decay_mult 0. The BASE
        lr is the group's initial_lr (what the LR hook scales), not whatever the schedule has made of it by now."""
        kw = dict(lr=0.01, momentum=0.9, weight_decay=1e-4)
        if self.optimizer is not None:
            g = self.optimizer.param_groups[0]
            d = self.optimizer.defaults
            kw = dict(lr=float(g.get("initial_lr", d.get("lr", g["lr"]))), momentum=float(d.get("momentum", 0.0)),
                      weight_decay=float(d.get("weight_decay", 0.0)))
        return kw

    def _engine_for(self, B, H_, W):
        key = (B, H_, W)
        if key in self._engines:
            self._engines.move_to_end(key)
            self.engine = self._engines[key]
            return self.engine
        m, hc = self.model, self.model.head_cfg
        if m.store.device.type != "cuda":
            raise RuntimeError("dsl_b200 SemiEpochBasedRunner: move the models to CUDA first (model.cuda()); there is no "
                               "CPU fallback")
        head_kwargs = dict(center_sampling=hc["center_sampling"], radius=hc["center_sample_radius"],
                           norm_on_bbox=hc["norm_on_bbox"], strides=hc["strides"], regress_ranges=hc["regress_ranges"])
        ema = self.ema_model is not None
        self.engine = DSLEngine(B, H_, W, depth=m.depth, num_classes=m.num_classes, device=m.store.device,
                                loss_weight=hc["loss_weight"], ema_keep=0.99, student_store=m.store,
                                teacher_store=self.ema_model.store if ema else None, max_grad_norm=self.max_grad_norm,
                                scale_invariant=self.scale_invariant, soft_weight=hc["soft_weight"],
                                soft_warm_up=hc["soft_warm_up"], head_kwargs=head_kwargs,
                                teacher_B=getattr(self, "pl_num_unlabeled", None),
                                backbone=getattr(m, "backbone_kind", "resnet"), **self._optimizer_hparams())
        self.engine.ema_in_step = False
        self._apply_ema_cfg(self.engine)
        if self._engines:
            self._share_state(next(reversed(self._engines.values())), self.engine)
        else:
            self.engine.lr_scale.fill_(self._lr_scale_host)
            if getattr(self, "_pending_mom", None) is not None:      # resume() ran before the first engine existed
                self.engine.mom.copy_(self._pending_mom.to(self.engine.mom.device))
        self._engines[key] = self.engine
        while len(self._engines) > self.max_cached_shapes:
            self._engines.popitem(last=False)
        m._dirty()
        if ema:
            self.ema_model._dirty()
        return self.engine

    def _sync_lr(self, eng):
        """LR schedule -> the device scalar the captured SGD kernels multiply their base LR with. The LR hook (mmcv's
        LrUpdaterHook or hooks.StepLrUpdaterHook) has just written param_group['lr'] in before_train_epoch / _iter; its
        warm-up and step factors multiply every group alike, so group 0's lr / initial_lr is THE factor."""
        if self.optimizer is None:
            return
        g = self.optimizer.param_groups[0]
        base = float(g.get("initial_lr", eng.lr))
        scale = float(g["lr"]) / base if base != 0.0 else 0.0
        if scale != self._lr_scale_host:
            eng.lr_scale.fill_(scale)       # shared by every shape's engine (_share_state)
            self._lr_scale_host = scale

    def current_lr(self):
        """mmcv BaseRunner.current_lr: the LR of every param group."""
        if self.optimizer is not None:
            return [group["lr"] for group in self.optimizer.param_groups]
        return [self.engine.lr * self._lr_scale_host] if self.engine is not None else []

    def current_momentum(self):
        if self.optimizer is not None:
            return [group.get("momentum", 0.0) for group in self.optimizer.param_groups]
        return [self.engine.momentum] if self.engine is not None else []

    def run_iter(self, data_batch, train_mode=True, **kwargs):
        if not train_mode:
            raise NotImplementedError("dsl_b200 SemiEpochBasedRunner: validation goes through FCOS.simple_test")
        img = _unwrap(data_batch["img"])
        metas = _unwrap(data_batch["img_metas"])
        gts = [g.float() for g in _unwrap(data_batch["gt_bboxes"])]
        labels = list(_unwrap(data_batch["gt_labels"]))
        ign = data_batch.get("gt_bboxes_ignore")
        ign = [g.float() for g in _unwrap(ign)] if ign is not None else None
        B, _, H_, W = img.shape
        eng = self._engine_for(B, H_, W)
        self._apply_ema_cfg(eng)
        self._sync_lr(eng)
        teacher_img = data_batch.get("teacher_img")
        teacher_img = _unwrap(teacher_img) if teacher_img is not None else None
        t_metas = data_batch.get("teacher_img_metas")
        if t_metas is not None:
            # the weak (test-pipeline) views the teacher sees: clip range, rescale to original-image coordinates
            # (simple_test(rescale=True), apis/test.py) and the image size the dataset rule checks (semicoco.py:236-246)
            t_metas = _unwrap(t_metas)
            eng.post.set_meta([m["img_shape"] for m in t_metas],
                              [[float(v) for v in m["scale_factor"]] for m in t_metas],
                              [m.get("ori_shape", m["img_shape"]) for m in t_metas])
        eng.cur_iter = self._si_iter
        use_pl = (getattr(self, "pl_enabled", False) and self.pl_start_iter is not None
                  and self._iter >= self.pl_start_iter and self._pl_steps >= self.pl_lag)
        if use_pl:
            from .geometry import view_from_meta
            tB = eng.teacher.B
            if ign is None:
                ign = [g.new_zeros((0, 4)) for g in gts]
            views = [view_from_meta(m) for m in metas[B - tB:]]
            eng.set_inputs_with_pseudo_labels(img, gts[:B - tB], labels[:B - tB], ign[:B - tB], views,
                                              teacher_img=teacher_img, pl=self._pl_source(eng))
        else:
            t_img = teacher_img if teacher_img is not None else img[B - eng.teacher.B:]
            eng.set_inputs(img, gts, labels, ign, teacher_img=t_img)
        self._si_iter = eng.cur_iter
        losses = eng.step()
        if getattr(self, "pl_enabled", False):
            self._pl_push(eng)
        if eng.ema_in_step:
            self.ema_flag = True
        log_vars = {k: float(v) for k, v in losses.items()}     # the reference's .item() per logged value (base.py:206)
        log_vars["loss"] = sum(v for k, v in log_vars.items() if "loss" in k)
        self.outputs = dict(loss=log_vars["loss"], log_vars=log_vars, num_samples=len(metas))
        self.log_buffer.update(log_vars, len(metas))

    def train(self, data_loader, **kwargs):
        self.mode = "train"
        self.data_loader = data_loader
        if self._max_epochs is not None:
            self._max_iters = self._max_epochs * len(data_loader)
        self.iter_tol_epoch = len(data_loader)
        if getattr(self, "pl_enabled", False) and self.pl_start_iter is None:
            self.pl_start_iter = int(self.pl_start_point * self.iter_tol_epoch)
        self.call_hook("before_train_epoch")
        for i, data_batch in enumerate(data_loader):
            self._inner_iter = i
            self.imagefiles = [m.get("filename") for m in _unwrap(data_batch["img_metas"])]
            self.call_hook("before_train_iter")
            self.run_iter(data_batch, train_mode=True, **kwargs)
            self.call_hook("after_train_iter")
            self._iter += 1
        # per-epoch adaptive thresholds (UnlabelPredHook.before_train_epoch -> adathres, unlabel_pred_hook.py:447-449)
        if self.engine is not None:
            self.engine.end_epoch()
            for eng in self._engines.values():     # the other shapes' engines: same history flag, capture again
                if eng is not self.engine and eng.post.have_prev != self.engine.post.have_prev:
                    eng.post.have_prev = self.engine.post.have_prev
                    eng.graphs = None
        self.call_hook("after_train_epoch")
        self._epoch += 1

    def run(self, data_loaders, workflow, max_epochs=None, **kwargs):
        assert isinstance(data_loaders, list) and len(data_loaders) == len(workflow)
        assert all(isinstance(f, tuple) for f in workflow)
        if max_epochs is not None:
            self._max_epochs = max_epochs
        assert self._max_epochs is not None, "max_epochs must be specified during instantiation"
        for i, (mode, _) in enumerate(workflow):
            if mode == "train":
                self._max_iters = self._max_epochs * len(data_loaders[i])
                break
        self.logger.info("Start running, work_dir: %s", self.work_dir if self.work_dir is not None else "NONE")
        self.logger.info("Hooks will be executed in the following order:\n%s", self.get_hook_info())
        self.logger.info("workflow: %s, max: %d epochs", workflow, self._max_epochs)
        self.call_hook("before_run")
        while self.epoch < self._max_epochs:
            for i, (mode, epochs) in enumerate(workflow):
                if not isinstance(mode, str):
                    raise TypeError("mode in workflow must be a str, but got {}".format(type(mode)))
                if mode != "train":
                    raise NotImplementedError("dsl_b200 SemiEpochBasedRunner: only the 'train' workflow is fused "
                                              "(validation: an EvalHook over FCOS.simple_test)")
                for _ in range(epochs):
                    if self.epoch >= self._max_epochs:
                        break
                    self.train(data_loaders[i], **kwargs)
        time.sleep(0)
        self.call_hook("after_run")

    def EMA(self, keep_rate=None, mode=None, start_point=None, **kwargs):
        """runner.EMA() (:368-409) as an explicit call: T <- (1 - k) S + k T over every state_dict entry. The fused
        step already does this every iteration when the EMA hook is iteration-mode / interval 1; this entry point
        serves epoch-mode hooks and scripts that call it themselves."""
        if self.ema_model is None:
            return
        k = keep_rate if keep_rate is not None else (self.fused_ema["ratio"] if self.fused_ema else 0.99)
        if self.engine is not None:
            self.engine.ema(k)          # + refresh of the teacher plan's derived operands
            self.ema_model._dirty()
        else:
            from .plugin import ema_update_
            ema_update_(self.ema_model, self.model, k)
        self.ema_flag = True

    def save_adathres(self, path, cat_names):
        """Write the epoch's per-class thresholds / class weights as the reference's adathres.json
        (unlabel_pred_hook.py:344-367), e.g. for its SemiCOCODataset(thres=path) or to continue the run there."""
        from . import formats
        assert self.engine is not None, "no step has run yet"
        torch.cuda.synchronize()
        formats.save_adathres(path, *self.engine.post.adathres_state(), cat_names)

    def load_adathres(self, path, cat_names, absent_thr=0.3):
        """Resume from an adathres.json the reference (or save_adathres) wrote: thresholds + next epoch's counting
        gate. The state is shared by every shape's engine; graphs captured without a history are dropped."""
        from . import formats
        assert self.engine is not None, "build an engine first (run one iteration or call _engine_for)"
        thr, counted = formats.adathres_from_json(path, cat_names, absent_thr)
        self.engine.post.load_adathres(thr, counted)
        for eng in self._engines.values():
            if eng.post is not self.engine.post:
                eng.post.have_prev = True
            eng.graphs = None

    # ---- checkpoints ---------------------------------------------------------------------------------------------
    def _param_offsets(self):
        """id(nn.Parameter) -> (offset, numel) inside the flat trainable range, for the optimizer-state mapping."""
        st = self.model.store
        return {id(p): st.offsets[spec.name] for spec, p in self.model.trainable_parameters()}

    def _optimizer_state_dict(self):
        """torch.optim.SGD.state_dict() layout with the fused momentum buffer scattered back to its parameters, so the
        reference (or plain torch) can resume from a checkpoint written here."""
        sd = self.optimizer.state_dict()
        off = self._param_offsets()
        state, idx = {}, 0
        mom = self.engine.mom.detach().cpu() if self.engine is not None else None
        for group in self.optimizer.param_groups:
            for p in group["params"]:
                if mom is not None and id(p) in off:
                    o, n = off[id(p)]
                    state[idx] = dict(momentum_buffer=mom[o:o + n].view(p.shape).clone())
                idx += 1
        sd["state"] = state
        return sd

    def _load_optimizer_state(self, osd):
        """Accepts both layouts: the flat one of save_checkpoint(optimizer=None) and torch's SGD state_dict."""
        if osd is None:
            return
        eng_moms = [e.mom for e in self._engines.values()]
        if "momentum_buffer" in osd:          # flat layout
            self._pending_mom = osd["momentum_buffer"]
        else:
            off = self._param_offsets()
            flat = torch.zeros(self.model.store.n_train, dtype=torch.float32)
            idx = 0
            params = [p for g in self.optimizer.param_groups for p in g["params"]] if self.optimizer is not None else []
            for p in params:
                st = osd.get("state", {}).get(idx)
                if st is not None and st.get("momentum_buffer") is not None and id(p) in off:
                    o, n = off[id(p)]
                    flat[o:o + n] = st["momentum_buffer"].reshape(-1).float().cpu()
                idx += 1
            self._pending_mom = flat
            if self.optimizer is not None:
                # param_groups carry lr / initial_lr of the schedule; the per-parameter state stays in the fused buffer
                groups = osd.get("param_groups")
                if groups is not None and len(groups) == len(self.optimizer.param_groups):
                    for g, saved in zip(self.optimizer.param_groups, groups):
                        for k, v in saved.items():
                            if k != "params":
                                g[k] = v
        for m in eng_moms:
            m.copy_(self._pending_mom.to(m.device))

    def load_checkpoint(self, filename, map_location="cpu", strict=False, revise_keys=((r"^module\.", ""),)):
        """semi_epoch_based_runner.py:350-366: the SAME file is loaded into the EMA teacher first and then into the
        student (mmcv load_checkpoint semantics: `state_dict` key optional, `revise_keys` regex renames, non-strict by
        default with the mismatches logged)."""
        self.logger.info("load checkpoint from %s", filename)
        ck = torch.load(filename, map_location=map_location, weights_only=False)
        sd = ck["state_dict"] if isinstance(ck, dict) and "state_dict" in ck else ck
        for pat, rep in revise_keys:
            sd = OrderedDict((re.sub(pat, rep, k), v) for k, v in sd.items())
        for name, mdl in (("ema_model", self.ema_model), ("model", self.model)):
            if mdl is None:
                continue
            res = mdl.load_state_dict(sd, strict=strict)
            missing = list(getattr(res, "missing_keys", []) or [])
            unexpected = list(getattr(res, "unexpected_keys", []) or [])
            if missing or unexpected:
                self.logger.warning("%s: missing keys %s, unexpected keys %s", name, missing[:8], unexpected[:8])
            mdl._dirty()
        for eng in self._engines.values():      # derived bf16 operands of every cached plan
            eng.student.repack(everything=True)
            eng.teacher.repack(everything=True)
        return ck if isinstance(ck, dict) else dict(state_dict=ck)

    def resume(self, checkpoint, resume_optimizer=True, map_location="default", ema_checkpoint=None):
        """mmcv BaseRunner.resume over this runner's load_checkpoint: epoch / iter / hook messages from `meta`, optimizer
        state (momentum buffer + the LR schedule's param_groups). Like the reference, the teacher is loaded from the
        SAME file as the student (:350-366) unless `ema_checkpoint` names the `<file>_ema` written next to it."""
        if map_location == "default":
            map_location = "cpu"    # the flat parameter buffers copy to the device themselves
        ck = self.load_checkpoint(checkpoint, map_location=map_location)
        if ema_checkpoint is not None and self.ema_model is not None:
            eck = torch.load(ema_checkpoint, map_location=map_location, weights_only=False)
            self.ema_model.load_state_dict(eck.get("state_dict", eck), strict=False)
            self.ema_model._dirty()
            for eng in self._engines.values():
                eng.teacher.repack(everything=True)
        self._epoch = ck["meta"]["epoch"]
        self._iter = ck["meta"]["iter"]
        if self.meta is None:
            self.meta = {}
        self.meta.setdefault("hook_msgs", {})
        self.meta["hook_msgs"].update(ck["meta"].get("hook_msgs", {}))
        if "optimizer" in ck and resume_optimizer:
            self._load_optimizer_state(ck["optimizer"])
        self.logger.info("resumed epoch %d, iter %d", self.epoch, self.iter)

    def save_checkpoint(self, out_dir, filename_tmpl="epoch_{}.pth", save_optimizer=True, meta=None,
                        create_symlink=True):
        """semi_epoch_based_runner.py:411-458: `<name>` for the student and `<name>_ema` for the teacher, both with the
        reference's state_dict names (loadable by the reference's load_checkpoint); `optimizer` in torch.optim.SGD's
        own state_dict layout when the runner has an optimizer object."""
        meta = dict(meta or {})
        if self.meta is not None:
            meta.update(self.meta)
        meta.update(epoch=self.epoch + 1, iter=self.iter)
        os.makedirs(out_dir, exist_ok=True)
        filename = filename_tmpl.format(self.epoch + 1)
        path = osp.join(out_dir, filename)
        torch.cuda.synchronize()
        ck = dict(meta=meta, state_dict={k: v.detach().cpu() for k, v in self.model.state_dict().items()})
        if save_optimizer and self.optimizer is not None:
            ck["optimizer"] = self._optimizer_state_dict()
        elif save_optimizer and self.engine is not None:
            ck["optimizer"] = dict(momentum_buffer=self.engine.mom.detach().cpu(), lr=self.engine.lr,
                                   momentum=self.engine.momentum, weight_decay=self.engine.wd)
        torch.save(ck, path)
        if self.ema_model is not None:
            torch.save(dict(meta=meta, state_dict={k: v.detach().cpu() for k, v in self.ema_model.state_dict().items()}),
                       path + "_ema")
        if create_symlink:
            dst = osp.join(out_dir, "latest.pth")
            if osp.lexists(dst):
                os.remove(dst)
            os.symlink(filename, dst)
        return path


def register(force=True):
    """RUNNERS['SemiEpochBasedRunner'] -> this class. Called by dsl_b200.plugin.register(runner=True) — i.e. only when
    the config asks for it (custom_imports of dsl_b200.plugin_runner): replacing the runner changes how a step is
    executed (fused), which a user who only wants the model classes must not get implicitly."""
    try:
        from mmcv.runner import RUNNERS
    except Exception:
        try:
            from mmcv.runner.builder import RUNNERS
        except Exception:
            return []
    RUNNERS.register_module(name="SemiEpochBasedRunner", force=force, module=SemiEpochBasedRunner)
    return ["RUNNERS.SemiEpochBasedRunner"]


_ = L
