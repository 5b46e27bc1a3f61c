"""SemiEpochBasedRunner on the fused engine: the reference's RUNNERS-registry runner
(mmdet/runner/hooks/semi_epoch_based_runner.py:49-458) restated over `DSLEngine`.

What the reference spreads over one iteration — the scale-invariant extra input (:186-204), `run_iter` ->
`model.train_step` (:142-167), mmcv's OptimizerHook (clip 35 + SGD), `EMAOWNHook` -> `runner.EMA()` (:368-409) and
`UnlabelPredHook`'s teacher inference (unlabel_pred_hook.py:512-562) — is ONE captured CUDA-graph step here. The class
keeps the reference's constructor arguments, counters (`epoch`, `iter`, `inner_iter`, `max_epochs`, `max_iters`),
hook protocol (`register_hook` / `call_hook` with the mmcv stage names), `train` / `run` / `EMA` / `save_checkpoint`
(`epoch_N.pth` + `epoch_N.pth_ema`, reference state_dict names) so a training script written against the reference
drives it unchanged. Batches are dicts with the reference's keys: `img` (B,3,H,W), `img_metas`, `gt_bboxes`,
`gt_labels`, `gt_bboxes_ignore`, plus optionally `teacher_img` (the weak-aug images of the unlabeled samples the EMA
teacher labels; the reference's hook reads them from disk); mmcv DataContainers (`.data[0]`) are unwrapped.

CUDA only; the model / ema_model must be `dsl_b200.plugin.FCOS` modules (their ParamStores are trained in place).
"""
import logging
import os
import os.path as osp
import time
from collections import OrderedDict

import torch

from . import _lib as L
from .trainer import DSLEngine


def _unwrap(v):
    """mmcv DataContainer -> its payload of the first (only) GPU."""
    d = getattr(v, "data", None)
    if d is not None and not isinstance(v, torch.Tensor):
        return d[0]
    return v


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
        self.ema_flag = e is not None        # reference: self.ema_flag (:121-125)
        self.scale_invariant = bool(scale_invariant)
        self.mode = None
        self._hooks = []
        self._epoch = self._iter = self._inner_iter = 0
        self._max_epochs, self._max_iters = max_epochs, max_iters
        self.outputs = None
        self.log_buffer = []
        self.imagefiles = []
        self.engine = None                   # the engine of the current batch shape
        # multi-scale training (BASELINE configs[4]: Resize [(1333,640),(1333,800)] + Pad(32) gives a handful of padded
        # shapes): one engine (plan + CUDA graph) per (B, H, W), least recently used first; all of them train the SAME
        # parameter stores and share the optimizer / adathres state (see _share_state)
        self._engines = OrderedDict()
        self.max_cached_shapes = 8
        self._si_iter = 0                    # FCOSHead.cur_iter of the SI-soft warm-up, across shapes
        self.ema_keep = 0.99                 # cfg ema_config ratio (configs/fcos_semi/*.py:199)

    # ---- counters (mmcv BaseRunner properties) ----------------------------------------------------------------
    epoch = property(lambda self: self._epoch)
    iter = property(lambda self: self._iter)
    inner_iter = property(lambda self: self._inner_iter)
    max_epochs = property(lambda self: self._max_epochs)
    max_iters = property(lambda self: self._max_iters)

    # ---- hooks -------------------------------------------------------------------------------------------------
    def register_hook(self, hook, priority="NORMAL"):
        """EMAOWNHook instances only configure the fused EMA (ratio); every other hook is called at its stages."""
        from .plugin import EMAOWNHook
        if isinstance(hook, EMAOWNHook):
            self.ema_keep = float(hook.ratio)
            for eng in self._engines.values():
                eng.ema_keep = self.ema_keep
                eng.graphs = None
            return
        self._hooks.append(hook)

    def call_hook(self, fn_name):
        for h in self._hooks:
            getattr(h, fn_name, lambda r: None)(self)

    # ---- engine ------------------------------------------------------------------------------------------------
    @staticmethod
    def _share_state(src, dst):
        """Everything a step carries over to the next one besides the weights must not depend on the batch shape: the
        SGD momentum buffer, the LR-schedule scalar and the epoch's adaptive-threshold statistics / thresholds are ONE set
        of tensors used by every shape's engine (assigned before the new engine captures its graph)."""
        dst.mom, dst.lr_scale = src.mom, src.lr_scale
        for name in ("stat_cnt", "stat_cum", "stat_prev", "thr_class", "class_weight", "have_prev"):
            setattr(dst.post, name, getattr(src.post, name))

    def _engine_for(self, B, H, W):
        key = (B, H, W)
        if key in self._engines:
            self._engines.move_to_end(key)
            self.engine = self._engines[key]
            return self.engine
        m, hc = self.model, self.model.head_cfg
        if m.store.device.type != "cuda":
            raise RuntimeError("dsl_b200 SemiEpochBasedRunner: move the models to CUDA first (model.cuda()); there is no "
                               "CPU fallback")
        kw = dict(lr=0.01, momentum=0.9, weight_decay=1e-4)
        if self.optimizer is not None:      # cfg optimizer (:182): SGD lr .01 momentum .9 wd 1e-4, bias lr x2 / decay x0
            g = self.optimizer.param_groups[0]
            kw = dict(lr=float(g["lr"]), momentum=float(g.get("momentum", 0.0)),
                      weight_decay=float(g.get("weight_decay", 0.0)))
        head_kwargs = dict(center_sampling=hc["center_sampling"], radius=hc["center_sample_radius"],
                           norm_on_bbox=hc["norm_on_bbox"], strides=hc["strides"], regress_ranges=hc["regress_ranges"])
        self.engine = DSLEngine(B, H, W, depth=m.depth, num_classes=m.num_classes, device=m.store.device,
                                loss_weight=hc["loss_weight"], ema_keep=self.ema_keep, student_store=m.store,
                                teacher_store=self.ema_model.store if self.ema_flag else None,
                                scale_invariant=self.scale_invariant, soft_weight=hc["soft_weight"],
                                soft_warm_up=hc["soft_warm_up"], head_kwargs=head_kwargs,
                                backbone=getattr(m, "backbone_kind", "resnet"), **kw)
        self.engine.ema_keep = self.ema_keep
        if self._engines:
            self._share_state(next(reversed(self._engines.values())), self.engine)
        self._engines[key] = self.engine
        while len(self._engines) > self.max_cached_shapes:
            self._engines.popitem(last=False)
        m._dirty()
        if self.ema_flag:
            self.ema_model._dirty()
        return self.engine

    def run_iter(self, data_batch, train_mode=True, **kwargs):
        if not train_mode:
            raise NotImplementedError("dsl_b200 SemiEpochBasedRunner: validation goes through FCOS.simple_test")
        img = _unwrap(data_batch["img"])
        metas = _unwrap(data_batch["img_metas"])
        gts = [g.float() for g in _unwrap(data_batch["gt_bboxes"])]
        labels = list(_unwrap(data_batch["gt_labels"]))
        ign = data_batch.get("gt_bboxes_ignore")
        ign = [g.float() for g in _unwrap(ign)] if ign is not None else None
        B, _, H, W = img.shape
        eng = self._engine_for(B, H, W)
        teacher_img = data_batch.get("teacher_img")
        eng.cur_iter = self._si_iter
        eng.set_inputs(img, gts, labels, ign, teacher_img=_unwrap(teacher_img) if teacher_img is not None else img)
        self._si_iter = eng.cur_iter
        losses = eng.step()
        log_vars = {k: float(v) for k, v in losses.items()}     # the reference's .item() per logged value (base.py:206)
        log_vars["loss"] = sum(v for k, v in log_vars.items() if "loss" in k)
        self.outputs = dict(loss=log_vars["loss"], log_vars=log_vars, num_samples=len(metas))
        self.log_buffer.append(log_vars)

    def train(self, data_loader, **kwargs):
        self.mode = "train"
        self.data_loader = data_loader
        if self._max_epochs is not None:
            self._max_iters = self._max_epochs * len(data_loader)
        self.call_hook("before_train_epoch")
        self.iter_tol_epoch = len(data_loader)
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
        if max_epochs is not None:
            self._max_epochs = max_epochs
        assert self._max_epochs is not None, "max_epochs must be specified during instantiation"
        self.logger.info("workflow: %s, max: %d epochs", workflow, self._max_epochs)
        self.call_hook("before_run")
        while self.epoch < self._max_epochs:
            for i, (mode, epochs) in enumerate(workflow):
                if mode != "train":
                    raise NotImplementedError("dsl_b200 SemiEpochBasedRunner: only the 'train' workflow is fused")
                for _ in range(epochs):
                    if self.epoch >= self._max_epochs:
                        break
                    self.train(data_loaders[i], **kwargs)
        time.sleep(0)
        self.call_hook("after_run")

    def EMA(self, keep_rate=None):
        """runner.EMA() (:368-409) as an explicit call: T <- (1 - k) S + k T over every state_dict entry. The fused
        step already does this every iteration; this entry point serves scripts that call it themselves."""
        if not self.ema_flag:
            return
        from .plugin import ema_update_
        ema_update_(self.ema_model, self.model, self.ema_keep if keep_rate is None else keep_rate)

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

    def current_lr(self):
        return [self.engine.lr] if self.engine is not None else []

    def save_checkpoint(self, out_dir, filename_tmpl="epoch_{}.pth", save_optimizer=True, meta=None,
                        create_symlink=True):
        """semi_epoch_based_runner.py:411-458: `<name>` for the student and `<name>_ema` for the teacher, both with the
        reference's state_dict names (loadable by the reference's load_checkpoint)."""
        meta = dict(meta or {})
        if self.meta is not None:
            meta.update(self.meta)
        meta.update(epoch=self.epoch + 1, iter=self.iter)
        os.makedirs(out_dir, exist_ok=True)
        filename = filename_tmpl.format(self.epoch + 1)
        path = osp.join(out_dir, filename)
        torch.cuda.synchronize()
        ck = dict(meta=meta, state_dict={k: v.detach().cpu() for k, v in self.model.state_dict().items()})
        if save_optimizer and self.engine is not None:
            ck["optimizer"] = dict(momentum_buffer=self.engine.mom.detach().cpu(), lr=self.engine.lr,
                                   momentum=self.engine.momentum, weight_decay=self.engine.wd)
        torch.save(ck, path)
        if self.ema_flag:
            torch.save(dict(meta=meta, state_dict={k: v.detach().cpu() for k, v in self.ema_model.state_dict().items()}),
                       path + "_ema")
        if create_symlink:
            dst = osp.join(out_dir, "latest.pth")
            if osp.lexists(dst):
                os.remove(dst)
            os.symlink(filename, dst)
        return path


def register(force=True):
    try:
        from mmcv.runner import RUNNERS
    except Exception:
        return []
    RUNNERS.register_module(name="SemiEpochBasedRunner", force=force, module=SemiEpochBasedRunner)
    return ["RUNNERS.SemiEpochBasedRunner"]


_ = L
