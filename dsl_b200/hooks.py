"""Hook side of the drop-in boundary.

`mmdet/apis/train.py:149-217` drives the runner through mmcv's hook machinery: `register_training_hooks(lr_config,
optimizer_config, ema_config, checkpoint_config, log_config, momentum_config)`, `register_hook(hook, priority)` (which
asserts `isinstance(hook, mmcv.runner.Hook)`), the stage calls `before_run ... after_train_iter ...`. mmcv-full (pinned
>=1.3.8,<=1.4.0, mmdet/__init__.py:19-27) is a third-party dependency that is not in the reference tree. When it is
importable every class here derives from / defers to mmcv's own (`Hook`, `HOOKS.build`, priorities); when it is not
(the GPU box), the small part of its published 1.3.x behaviour the fcos_semi configs use is restated below so the same
config keys keep working:

  Hook                      stage methods + every_n_epochs / every_n_iters / every_n_inner_iters / end_of_epoch /
                            is_last_epoch / is_last_iter                             (mmcv/runner/hooks/hook.py)
  LrUpdaterHook, StepLrUpdaterHook   policy='step', step=[..], gamma, min_lr, warmup 'constant' | 'linear' | 'exp',
                            warmup_iters / warmup_ratio / warmup_by_epoch, by_epoch   (mmcv/runner/hooks/lr_updater.py)
                            = configs/fcos_semi/RLA_*.py:188-194 (linear warm-up 500 it at 1/3, step [20, 26])
  OptimizerHook             grad_clip holder (the fused step does backward + clip + SGD itself)
  CheckpointHook            interval / by_epoch / save_optimizer / out_dir / max_keep_ckpts -> runner.save_checkpoint
  IterTimerHook, TextLoggerHook      time / data_time and a plain-text line every `interval` iterations
  DistSamplerSeedHook       sampler.set_epoch(runner.epoch)

EMAOWNHook (mmdet/runner/hooks/ema.py:4-42) is the reference's own hook and lives here too.
"""
import datetime
import logging
import os
import os.path as osp
import time
from collections import OrderedDict

try:   # a real mmcv wins: its Hook is what mmcv's register_hook asserts on
    from mmcv.runner import Hook as _MMCVHook
    from mmcv.runner import HOOKS as _MMCV_HOOKS
    HAVE_MMCV = not getattr(__import__("mmcv"), "__dslb_stub__", False) or hasattr(_MMCVHook, "before_run")
except Exception:   # noqa: BLE001  (absent or a partial stub)
    _MMCVHook, _MMCV_HOOKS, HAVE_MMCV = None, None, False


# mmcv/runner/priority.py
PRIORITY = dict(HIGHEST=0, VERY_HIGH=10, HIGH=30, ABOVE_NORMAL=40, NORMAL=50, BELOW_NORMAL=60, LOW=70, VERY_LOW=90,
                LOWEST=100)


def get_priority(priority):
    if isinstance(priority, int):
        if not 0 <= priority <= 100:
            raise ValueError("priority must be between 0 and 100")
        return priority
    if isinstance(priority, str):
        return PRIORITY[priority.upper()]
    if hasattr(priority, "value"):
        return int(priority.value)
    raise TypeError("priority must be an integer or Priority enum value")


class _Hook:
    """mmcv.runner.Hook (hooks/hook.py) restated: every stage is a no-op, the *_train_* / *_val_* stages fall through
    to the generic ones."""
    stages = ("before_run", "before_train_epoch", "before_train_iter", "after_train_iter", "after_train_epoch",
              "before_val_epoch", "before_val_iter", "after_val_iter", "after_val_epoch", "after_run")

    def before_run(self, runner):
        pass

    def after_run(self, runner):
        pass

    def before_epoch(self, runner):
        pass

    def after_epoch(self, runner):
        pass

    def before_iter(self, runner):
        pass

    def after_iter(self, runner):
        pass

    def before_train_epoch(self, runner):
        self.before_epoch(runner)

    def before_val_epoch(self, runner):
        self.before_epoch(runner)

    def after_train_epoch(self, runner):
        self.after_epoch(runner)

    def after_val_epoch(self, runner):
        self.after_epoch(runner)

    def before_train_iter(self, runner):
        self.before_iter(runner)

    def before_val_iter(self, runner):
        self.before_iter(runner)

    def after_train_iter(self, runner):
        self.after_iter(runner)

    def after_val_iter(self, runner):
        self.after_iter(runner)

    def every_n_epochs(self, runner, n):
        return (runner.epoch + 1) % n == 0 if n > 0 else False

    def every_n_inner_iters(self, runner, n):
        return (runner.inner_iter + 1) % n == 0 if n > 0 else False

    def every_n_iters(self, runner, n):
        return (runner.iter + 1) % n == 0 if n > 0 else False

    def end_of_epoch(self, runner):
        return runner.inner_iter + 1 == len(runner.data_loader)

    def is_last_epoch(self, runner):
        return runner.epoch + 1 == runner._max_epochs

    def is_last_iter(self, runner):
        return runner.iter + 1 == runner._max_iters


Hook = _MMCVHook if _MMCVHook is not None and hasattr(_MMCVHook, "before_run") else _Hook


def is_hook(obj):
    """What register_hook accepts: an mmcv Hook, or one of the restated hooks (mmcv asserts isinstance(hook, Hook))."""
    return isinstance(obj, (Hook, _Hook))


# ---------------------------------------------------------------------------------------------------- LR schedule
class LrUpdaterHook(_Hook):
    """mmcv/runner/hooks/lr_updater.py::LrUpdaterHook (1.3.x): writes param_group['lr'] of runner.optimizer; the fused
    runner turns param_groups[0]['lr'] / initial_lr into the device-side LR scalar of the captured SGD kernels."""

    def __init__(self, by_epoch=True, warmup=None, warmup_iters=0, warmup_ratio=0.1, warmup_by_epoch=False):
        if warmup is not None:
            if warmup not in ("constant", "linear", "exp"):
                raise ValueError(f'"{warmup}" is not a supported type for warming up, valid types are "constant" and '
                                 '"linear"')
            assert warmup_iters > 0, '"warmup_iters" must be a positive integer'
            assert 0 < warmup_ratio <= 1.0, '"warmup_ratio" must be in range (0,1]'
        self.by_epoch, self.warmup, self.warmup_iters, self.warmup_ratio = by_epoch, warmup, warmup_iters, warmup_ratio
        self.warmup_by_epoch = warmup_by_epoch
        if self.warmup_by_epoch:
            self.warmup_epochs, self.warmup_iters = self.warmup_iters, None
        else:
            self.warmup_epochs = None
        self.base_lr, self.regular_lr = [], []

    def _set_lr(self, runner, lr_groups):
        for group, lr in zip(runner.optimizer.param_groups, lr_groups):
            group["lr"] = lr

    def get_lr(self, runner, base_lr):
        raise NotImplementedError

    def get_regular_lr(self, runner):
        return [self.get_lr(runner, b) for b in self.base_lr]

    def get_warmup_lr(self, cur_iters):
        if self.warmup == "constant":
            return [lr * self.warmup_ratio for lr in self.regular_lr]
        if self.warmup == "linear":
            k = (1 - cur_iters / self.warmup_iters) * (1 - self.warmup_ratio)
            return [lr * (1 - k) for lr in self.regular_lr]
        k = self.warmup_ratio ** (1 - cur_iters / self.warmup_iters)
        return [lr * k for lr in self.regular_lr]

    def before_run(self, runner):
        for group in runner.optimizer.param_groups:
            group.setdefault("initial_lr", group["lr"])
        self.base_lr = [group["initial_lr"] for group in runner.optimizer.param_groups]

    def before_train_epoch(self, runner):
        if self.warmup_iters is None:
            self.warmup_iters = self.warmup_epochs * len(runner.data_loader)
        if not self.by_epoch:
            return
        self.regular_lr = self.get_regular_lr(runner)
        self._set_lr(runner, self.regular_lr)

    def before_train_iter(self, runner):
        cur_iter = runner.iter
        if not self.by_epoch:
            self.regular_lr = self.get_regular_lr(runner)
            if self.warmup is None or cur_iter >= self.warmup_iters:
                self._set_lr(runner, self.regular_lr)
            else:
                self._set_lr(runner, self.get_warmup_lr(cur_iter))
        elif self.warmup is not None and cur_iter <= self.warmup_iters:
            if cur_iter == self.warmup_iters:
                self._set_lr(runner, self.regular_lr)
            else:
                self._set_lr(runner, self.get_warmup_lr(cur_iter))


class FixedLrUpdaterHook(LrUpdaterHook):
    def get_lr(self, runner, base_lr):
        return base_lr


class StepLrUpdaterHook(LrUpdaterHook):
    """policy='step' (lr_updater.py::StepLrUpdaterHook): lr = base * gamma ** (number of passed steps), floor min_lr."""

    def __init__(self, step, gamma=0.1, min_lr=None, **kwargs):
        if isinstance(step, list):
            assert all(isinstance(s, int) and s > 0 for s in step)
        elif isinstance(step, int):
            assert step > 0
        else:
            raise TypeError('"step" must be a list or integer')
        self.step, self.gamma, self.min_lr = step, gamma, min_lr
        super().__init__(**kwargs)

    def get_lr(self, runner, base_lr):
        progress = runner.epoch if self.by_epoch else runner.iter
        if isinstance(self.step, int):
            exp = progress // self.step
        else:
            exp = len(self.step)
            for i, s in enumerate(self.step):
                if progress < s:
                    exp = i
                    break
        lr = base_lr * (self.gamma ** exp)
        if self.min_lr is not None:
            lr = max(lr, self.min_lr)
        return lr


# ---------------------------------------------------------------------------------------------------- optimizer shell
class OptimizerHook(_Hook):
    """mmcv's OptimizerHook carries only `grad_clip`; its after_train_iter (zero_grad / backward / clip / step) is what
    the fused CUDA-graph step already does, so the runner reads grad_clip from it and never calls it."""

    def __init__(self, grad_clip=None, **kwargs):
        self.grad_clip = grad_clip


def grad_clip_of(optimizer_config):
    """(max_norm or None) from cfg.optimizer_config (a dict, an mmcv / restated OptimizerHook, or None)."""
    if optimizer_config is None:
        return None
    if isinstance(optimizer_config, dict):
        t = optimizer_config.get("type", "OptimizerHook")
        if t not in ("OptimizerHook", "DistOptimizerHook"):
            raise NotImplementedError(f"dsl_b200 runner: optimizer hook type {t} (the fused step implements OptimizerHook: "
                                      "backward, clip_grad_norm_, SGD step)")
        gc = optimizer_config.get("grad_clip")
    else:
        if type(optimizer_config).__name__ not in ("OptimizerHook", "DistOptimizerHook"):
            raise NotImplementedError(f"dsl_b200 runner: optimizer hook {type(optimizer_config).__name__}")
        gc = getattr(optimizer_config, "grad_clip", None)
    if gc is None:
        return None
    if float(gc.get("norm_type", 2)) != 2.0:
        raise NotImplementedError("dsl_b200 runner: grad_clip.norm_type must be 2 (dslb_sq_norm / dslb_clip_coef)")
    return float(gc["max_norm"])


# ---------------------------------------------------------------------------------------------------- bookkeeping hooks
class CheckpointHook(_Hook):
    """mmcv/runner/hooks/checkpoint.py (1.3.x) for by_epoch / by-iteration saves through runner.save_checkpoint."""

    def __init__(self, interval=-1, by_epoch=True, save_optimizer=True, out_dir=None, max_keep_ckpts=-1, save_last=True,
                 **kwargs):
        self.interval, self.by_epoch, self.save_optimizer = interval, by_epoch, save_optimizer
        self.out_dir, self.max_keep_ckpts, self.save_last, self.args = out_dir, max_keep_ckpts, save_last, kwargs

    def before_run(self, runner):
        if not self.out_dir:
            self.out_dir = runner.work_dir

    def _save(self, runner):
        if getattr(runner, "rank", 0) != 0 or not self.out_dir:
            return
        runner.save_checkpoint(self.out_dir, save_optimizer=self.save_optimizer, **self.args)
        if self.max_keep_ckpts > 0:
            cur = (runner.epoch if self.by_epoch else runner.iter) + 1
            tmpl = self.args.get("filename_tmpl", "epoch_{}.pth" if self.by_epoch else "iter_{}.pth")
            for k in range(cur - self.max_keep_ckpts * self.interval, 0, -self.interval):
                path = osp.join(self.out_dir, tmpl.format(k))
                if not osp.exists(path):
                    break
                os.remove(path)
                if osp.exists(path + "_ema"):
                    os.remove(path + "_ema")

    def after_train_epoch(self, runner):
        if not self.by_epoch:
            return
        if self.every_n_epochs(runner, self.interval) or (self.save_last and self.is_last_epoch(runner)):
            runner.logger.info(f"Saving checkpoint at {runner.epoch + 1} epochs")
            self._save(runner)

    def after_train_iter(self, runner):
        if self.by_epoch:
            return
        if self.every_n_iters(runner, self.interval) or (self.save_last and self.is_last_iter(runner)):
            runner.logger.info(f"Saving checkpoint at {runner.iter + 1} iterations")
            self._save(runner)


class IterTimerHook(_Hook):
    def before_epoch(self, runner):
        self.t = time.time()

    def before_iter(self, runner):
        runner.log_buffer.update({"data_time": time.time() - self.t})

    def after_iter(self, runner):
        runner.log_buffer.update({"time": time.time() - self.t})
        self.t = time.time()


class LogBuffer:
    """mmcv/runner/log_buffer.py: running histories, `average(n)` over the last n entries into `output`."""

    def __init__(self):
        self.val_history, self.n_history, self.output, self.ready = OrderedDict(), OrderedDict(), OrderedDict(), False

    def clear(self):
        self.val_history.clear()
        self.n_history.clear()
        self.clear_output()

    def clear_output(self):
        self.output.clear()
        self.ready = False

    def update(self, vars, count=1):
        assert isinstance(vars, dict)
        for k, v in vars.items():
            self.val_history.setdefault(k, []).append(v)
            self.n_history.setdefault(k, []).append(count)

    def average(self, n=0):
        assert n >= 0
        for k, vals in self.val_history.items():
            v, c = vals[-n:], self.n_history[k][-n:]
            self.output[k] = sum(a * b for a, b in zip(v, c)) / max(sum(c), 1)
        self.ready = True


class TextLoggerHook(_Hook):
    """A plain-text line every `interval` iterations (mmcv/runner/hooks/logger/text.py, without the JSON log file)."""

    def __init__(self, interval=10, ignore_last=True, reset_flag=False, by_epoch=True, **kwargs):
        self.interval, self.ignore_last, self.reset_flag, self.by_epoch = interval, ignore_last, reset_flag, by_epoch

    def before_run(self, runner):
        self.start_iter = runner.iter

    def before_epoch(self, runner):
        runner.log_buffer.clear()

    def after_train_iter(self, runner):
        if self.every_n_inner_iters(runner, self.interval) or (self.end_of_epoch(runner) and not self.ignore_last):
            runner.log_buffer.average(self.interval)
        if runner.log_buffer.ready:
            out = runner.log_buffer.output
            lr = runner.current_lr()
            msg = f"Epoch [{runner.epoch + 1}][{runner.inner_iter + 1}/{len(runner.data_loader)}]\tlr: {lr[0]:.3e}, "
            if "time" in out:
                done = runner.iter - self.start_iter + 1
                eta = out["time"] * (runner.max_iters - runner.iter - 1)
                msg += f"eta: {datetime.timedelta(seconds=int(eta))}, time: {out['time']:.3f}, "
                msg += f"data_time: {out.get('data_time', 0.0):.3f}, "
                _ = done
            msg += ", ".join(f"{k}: {v:.4f}" for k, v in out.items() if k not in ("time", "data_time"))
            runner.logger.info(msg)
            runner.log_buffer.clear_output()

    def after_train_epoch(self, runner):
        runner.log_buffer.clear_output()


class DistSamplerSeedHook(_Hook):
    def before_epoch(self, runner):
        dl = runner.data_loader
        if hasattr(getattr(dl, "sampler", None), "set_epoch"):
            dl.sampler.set_epoch(runner.epoch)
        elif hasattr(getattr(getattr(dl, "batch_sampler", None), "sampler", None), "set_epoch"):
            dl.batch_sampler.sampler.set_epoch(runner.epoch)


LOCAL_HOOKS = dict(StepLrUpdaterHook=StepLrUpdaterHook, FixedLrUpdaterHook=FixedLrUpdaterHook, OptimizerHook=OptimizerHook,
                   CheckpointHook=CheckpointHook, IterTimerHook=IterTimerHook, TextLoggerHook=TextLoggerHook,
                   DistSamplerSeedHook=DistSamplerSeedHook)


def build_hook(cfg, default_type=None):
    """dict(type=...) -> hook. mmcv's HOOKS registry answers when mmcv is importable (so every hook mmcv or mmdet
    registers works); otherwise the restated table above, and an unknown type is an error — never silently dropped."""
    if not isinstance(cfg, dict):
        return cfg
    cfg = dict(cfg)
    if default_type is not None:
        cfg.setdefault("type", default_type)
    t = cfg["type"]
    if t == "EMAOWNHook":
        cfg.pop("type")
        return EMAOWNHook(**cfg)
    if _MMCV_HOOKS is not None and _MMCV_HOOKS.get(t) is not None:
        cls = _MMCV_HOOKS.get(t)
        cfg.pop("type")
        return cls(**cfg)
    if t in LOCAL_HOOKS:
        cfg.pop("type")
        return LOCAL_HOOKS[t](**cfg)
    raise KeyError(f"dsl_b200 runner: hook type {t!r} is neither in mmcv's HOOKS registry (mmcv importable: "
                   f"{_MMCV_HOOKS is not None}) nor one of {sorted(LOCAL_HOOKS)}")


# ---------------------------------------------------------------------------------------------------- EMA hook
def _unwrap(m):
    return m.module if hasattr(m, "module") else m


class EMAOWNHook(Hook):
    """EMAOWNHook (mmdet/runner/hooks/ema.py:4-42): same constructor keywords and trigger rules. With the fused runner
    the EMA of `mode="iteration", interval=1` is part of the captured step (runner.register_hook reads ratio /
    start_point from the hook); every other setting reaches runner.EMA() from the stages below, which for dsl_b200
    modules is ONE dslb_ema_update launch over the flat parameter buffers (bit-exact with the reference expression).
    No barriers are needed because nothing touches the file system."""

    def __init__(self, interval=-1, mode="epoch", ratio=0.99, start_point=-1, step_decay=None, decay_ratio=0.1,
                 **kwargs):
        self.interval, self.mode, self.start_point, self.ratio = interval, mode, start_point, ratio
        self.args, self.step_decay, self.decay_ratio = kwargs, step_decay, decay_ratio
        self.fused = False   # set by the fused runner when the EMA runs inside its captured step

    def every_n_epochs(self, runner, n):
        return (runner.epoch + 1) % n == 0 if n > 0 else False

    def every_n_iters(self, runner, n):
        return (runner.iter + 1) % n == 0 if n > 0 else False

    def _ema(self, runner):
        from .plugin import _StoreModule, ema_update_
        s, t = _unwrap(runner.model), _unwrap(runner.ema_model)
        if getattr(runner, "fused_ema", None) is not None:
            runner.EMA(keep_rate=self.ratio)
        elif isinstance(s, _StoreModule) and isinstance(t, _StoreModule):
            ema_update_(t, s, self.ratio)
            runner.ema_flag = True
        else:
            runner.EMA(keep_rate=self.ratio, mode=self.mode, start_point=self.start_point, **self.args)

    def after_train_epoch(self, runner):
        if self.step_decay is not None and runner.epoch + 1 in self.step_decay:
            self.ratio = max(1.0 - (1.0 - self.ratio) / self.decay_ratio, 0.01)
            runner.logger.info("[INFO] ema ratio changes to %f", self.ratio)
            if self.fused:
                runner.set_ema_ratio(self.ratio)
        if self.mode != "epoch" or self.interval == -1 or self.start_point > runner.epoch + 1:
            return
        if self.every_n_epochs(runner, self.interval):
            self._ema(runner)

    def after_train_iter(self, runner):
        if self.fused or self.mode != "iteration" or self.interval == -1 or self.start_point > runner.iter + 1:
            return
        if self.every_n_iters(runner, self.interval):
            self._ema(runner)


_ = logging
