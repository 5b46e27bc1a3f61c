"""The drop-in claim, executed: the reference's OWN `mmdet/apis/train.py::train_detector` (loaded in place from the
reference tree, under the mmcv stub's runner-side shells) drives `dsl_b200.runner.SemiEpochBasedRunner` +
`dsl_b200.plugin.FCOS` through its whole call sequence — MMDataParallel wrapping, build_optimizer with the config's
paramwise_cfg, build_runner with the reference's default_args, `ema_flag / ITER / timestamp`, register_training_hooks
with six positionals, custom hooks, load_checkpoint / resume — up to the first `run_iter`, which on this GPU-less box
must stop with the loud "move the models to CUDA first" error (there is no CPU path). The GPU twin of this test runs
the iterations for real. Skipped where the reference tree is absent (the GPU box).

Also pinned here without any reference code: the LR schedule of configs/fcos_semi/RLA_*.py:188-194 as the restated
StepLrUpdaterHook produces it (closed form of mmcv's lr_updater.py), checkpoint save -> resume round trips in both
optimizer-state layouts, hook priorities.
"""
import importlib
import logging
import os
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

from tests.test_plugin import MODEL_CFG, _build


def _same_weights(a, b):
    """Every state_dict entry equal (the flat buffers also hold alignment gaps no checkpoint carries)."""
    sa, sb = a.state_dict(), b.state_dict()
    return sa.keys() == sb.keys() and all(torch.equal(sa[k], sb[k]) for k in sa)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Cfg(dict):
    """mmcv.Config / ConfigDict stand-in: attribute access over nested dicts."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return Cfg(v) if isinstance(v, dict) and not isinstance(v, Cfg) else v

    def __setattr__(self, k, v):
        self[k] = v

    def get(self, k, default=None):
        v = dict.get(self, k, default)
        return Cfg(v) if isinstance(v, dict) and not isinstance(v, Cfg) else v


def shipped_cfg(work_dir, **over):
    """The schedule / hook part of configs/fcos_semi/RLA_r50_caffe_mslonger_tricks_0.Xdata_unlabel_dynamic_lw_nofuse_
    iterlabel_si-soft_singlestage.py:179-216 (values verbatim)."""
    c = Cfg(
        log_level="INFO", gpu_ids=[0], seed=0, work_dir=work_dir,
        data=dict(samples_per_gpu=2, workers_per_gpu=2),
        optimizer=dict(type="SGD", lr=0.01, momentum=0.9, weight_decay=0.0001,
                       paramwise_cfg=dict(bias_lr_mult=2., bias_decay_mult=0.)),
        optimizer_config=dict(grad_clip=dict(max_norm=35, norm_type=2)),
        lr_config=dict(policy="step", warmup="linear", warmup_iters=500, warmup_ratio=1.0 / 3, step=[20, 26]),
        runner=dict(type="SemiEpochBasedRunner", max_epochs=28),
        checkpoint_config=dict(interval=1),
        ema_config=dict(interval=1, mode="iteration", ratio=0.99, start_point=1),
        scale_invariant=True,
        log_config=dict(interval=10, hooks=[dict(type="TextLoggerHook")]),
        custom_hooks=[dict(type="ProbeHook", priority="LOW")],
        resume_from=None, load_from=None, workflow=[("train", 1)])
    c.update(over)
    return c


def _batch(B, H, W, seed):
    from tests.golden import inputs as GI
    rng = np.random.RandomState(seed)
    gts, labels, ignores = GI.make_gt(seed, B, H, W, with_ignore=True)
    metas = [dict(filename=f"im{seed}_{b}.jpg", img_shape=(H, W, 3), pad_shape=(H, W, 3), ori_shape=(H, W, 3),
                  scale_factor=np.ones(4, np.float32), flip=False) for b in range(B)]
    return dict(img=GI.make_tensor(rng, B, 3, H, W, scale=50.0), img_metas=metas, gt_bboxes=gts, gt_labels=labels,
                gt_bboxes_ignore=ignores)


@pytest.fixture()
def reference_train(monkeypatch):
    """mmdet.apis.train of the reference, imported in place with shells for what it imports besides mmcv."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present (GPU box)")
    ref_loader.load()
    import mmcv
    import mmcv.runner as mr
    from dsl_b200 import hooks as H
    from dsl_b200 import plugin, runner as R
    assert "RUNNERS.SemiEpochBasedRunner" not in plugin.REGISTERED      # importing the plugin does not swap the runner
    mr.RUNNERS.register_module(name="SemiEpochBasedRunner", force=True, module=R.SemiEpochBasedRunner)
    plugin.register(runner=True)

    built = []
    orig = mr.build_runner

    def build_runner(cfg, default_args=None):
        r = orig(cfg, default_args=default_args)
        built.append(r)
        return r

    class ProbeHook(H._Hook):       # a user hook from cfg.custom_hooks, built through HOOKS like the reference does
        calls = []

        def before_run(self, runner):
            ProbeHook.calls.append("before_run")

        def before_train_epoch(self, runner):
            ProbeHook.calls.append("before_train_epoch")

        def before_train_iter(self, runner):
            ProbeHook.calls.append(("before_train_iter", runner.current_lr()[0]))

    ProbeHook.calls = []
    mr.HOOKS.register_module(name="ProbeHook", force=True, module=ProbeHook)
    monkeypatch.setattr(mr, "build_runner", build_runner)

    def shell(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        monkeypatch.setitem(sys.modules, name, m)
        return m

    loaders = {}
    shell("mmdet.runner", DistSamplerSeedHook_semi=type("DistSamplerSeedHook_semi", (mr.Hook,), {}))
    shell("mmdet.runner.hooks", UnlabelPredHook=type("UnlabelPredHook", (mr.Hook,), {}),
          SemiEpochBasedRunner=R.SemiEpochBasedRunner)
    shell("mmdet.datasets", build_dataloader=lambda ds, *a, **k: loaders.setdefault(id(ds), list(ds.batches)),
          build_dataset=None, build_multi_dataloader=None, replace_ImageToTensor=None)
    core = sys.modules["mmdet.core"]
    monkeypatch.setattr(core, "DistEvalHook", type("DistEvalHook", (mr.Hook,), {}), raising=False)
    monkeypatch.setattr(core, "EvalHook", type("EvalHook", (mr.Hook,), {}), raising=False)
    ref_loader._shell("mmdet.apis")
    monkeypatch.delitem(sys.modules, "mmdet.apis.train", raising=False)
    train = importlib.import_module("mmdet.apis.train")
    assert train.__file__.startswith(ref_loader.REF_ROOT)
    _ = mmcv
    return types.SimpleNamespace(train_detector=train.train_detector, built=built, probe=ProbeHook, H=H, R=R)


def _models_on_cpu_pretending_cuda(monkeypatch):
    student, teacher = _build(), _build()
    for m in (student, teacher):
        monkeypatch.setattr(m, "cuda", lambda *a, _m=m, **k: _m, raising=False)   # no GPU here: .cuda() is a no-op
    return student, teacher


def test_reference_train_detector_drives_the_plugin_runner(reference_train, monkeypatch, tmp_path):
    T = reference_train
    student, teacher = _models_on_cpu_pretending_cuda(monkeypatch)
    # cfg.load_from: a checkpoint in the reference's format (DataParallel-prefixed keys, as mmcv's revise_keys expects)
    src = _build()
    with torch.no_grad():
        src.store.flat.normal_(0, 0.02)
    ck_path = str(tmp_path / "init.pth")
    torch.save(dict(meta=dict(epoch=3, iter=12), state_dict={"module." + k: v.clone() for k, v in
                                                             src.state_dict().items()}), ck_path)
    cfg = shipped_cfg(str(tmp_path / "work"), load_from=ck_path)
    dataset = types.SimpleNamespace(batches=[_batch(2, 64, 96, 1), _batch(2, 64, 96, 2)], CLASSES=("c",) * 80)
    with pytest.raises(RuntimeError, match="move the models to CUDA first"):
        T.train_detector(student, dataset, cfg, distributed=False, validate=False, timestamp="20261017_000000",
                         meta=dict(seed=0), ema_model=teacher)
    # --- everything train_detector did before the first fused step ---
    (runner,) = T.built
    assert isinstance(runner, T.R.SemiEpochBasedRunner)
    assert runner.model is student and runner.ema_model is teacher          # MMDataParallel wrappers unwrapped
    assert runner.scale_invariant is True and runner.max_epochs == 28 and runner.max_iters == 28 * 2
    assert runner.timestamp == "20261017_000000" and runner.ema_flag is False and runner.ITER is None
    assert runner.work_dir == str(tmp_path / "work") and os.path.isdir(runner.work_dir)
    # hooks: priority order of mmcv's register_training_hooks table; the optimizer hook is NOT a hook here
    names = [(type(h).__name__, h.priority) for h in runner.hooks]
    assert names == [("StepLrUpdaterHook", 10), ("EMAOWNHook", 45), ("CheckpointHook", 50), ("IterTimerHook", 70),
                     ("ProbeHook", 70), ("TextLoggerHook", 90)], names
    assert runner.max_grad_norm == 35.0
    assert runner.fused_ema == dict(ratio=0.99, start_point=1) and runner.hooks[1].fused
    # optimizer built by the reference's call on the plugin's flat-view parameters, paramwise bias rule applied
    opt = runner.optimizer
    named = dict(student.named_parameters())
    by_id = {id(p): n for n, p in named.items()}
    lrs = {by_id[id(g["params"][0])]: (g["initial_lr"], g["weight_decay"]) for g in opt.param_groups}
    assert lrs["backbone.layer2.0.conv1.weight"] == (0.01, 0.0001)
    assert lrs["bbox_head.conv_cls.bias"] == (0.02, 0.0) and lrs["neck.lateral_convs.0.conv.bias"] == (0.02, 0.0)
    assert lrs["bbox_head.cls_convs.0.gn.bias"] == (0.01, 0.0001)            # norm layers keep the base rule
    # the stage calls reached the hooks in order, and the LR hook had written the warm-up LR of iteration 0
    assert T.probe.calls[:2] == ["before_run", "before_train_epoch"]
    tag, lr0 = T.probe.calls[2]
    assert tag == "before_train_iter" and abs(lr0 - 0.01 / 3) < 1e-12
    assert runner.current_lr()[0] == lr0 and runner._hooks[0].base_lr[0] == 0.01
    # load_from went into BOTH models (semi_epoch_based_runner.py:350-366), "module." prefixes stripped
    for m in (student, teacher):
        assert _same_weights(m, src)
    assert (runner.epoch, runner.iter) == (0, 0)                            # load_from does not resume counters


def test_reference_train_detector_resume_path(reference_train, monkeypatch, tmp_path):
    """cfg.resume_from: counters, hook messages, the schedule's param_groups and the momentum buffers of a checkpoint
    in torch.optim.SGD's own state_dict layout (what the reference's save_checkpoint writes) reach the fused runner."""
    T = reference_train
    student, teacher = _models_on_cpu_pretending_cuda(monkeypatch)
    src = _build()
    with torch.no_grad():
        src.store.flat.normal_(0, 0.02)
    import mmcv.runner as mr
    ref_opt = mr.build_optimizer(src, shipped_cfg("x").optimizer)
    g = torch.Generator().manual_seed(5)
    for grp in ref_opt.param_groups:
        grp["initial_lr"] = grp["lr"]
        grp["lr"] = grp["lr"] * 0.1
        p = grp["params"][0]
        if p.requires_grad:
            ref_opt.state[p]["momentum_buffer"] = torch.randn(p.shape, generator=g)
    ck_path = str(tmp_path / "epoch_21.pth")
    torch.save(dict(meta=dict(epoch=21, iter=42, hook_msgs=dict(last_ckpt="epoch_21.pth")),
                    state_dict=src.state_dict(), optimizer=ref_opt.state_dict()), ck_path)
    cfg = shipped_cfg(str(tmp_path / "work"), resume_from=ck_path)
    dataset = types.SimpleNamespace(batches=[_batch(2, 64, 96, 1), _batch(2, 64, 96, 2)], CLASSES=("c",) * 80)
    with pytest.raises(RuntimeError, match="move the models to CUDA first"):
        T.train_detector(student, dataset, cfg, distributed=False, validate=False, timestamp="t", meta=None,
                         ema_model=teacher)
    (runner,) = T.built
    assert (runner.epoch, runner.iter) == (21, 42) and runner.meta["hook_msgs"]["last_ckpt"] == "epoch_21.pth"
    assert _same_weights(student, src) and _same_weights(teacher, src)
    # epoch 21 is past step 20: the LR hook restarts from initial_lr and applies gamma once; warm-up is over (iter 42
    # < 500 would still warm up in mmcv too: cur_iter <= warmup_iters -> warm-up LR of iteration 42 on the stepped LR)
    k = (1 - 42 / 500) * (1 - 1.0 / 3)
    assert abs(runner.current_lr()[0] - 0.01 * 0.1 * (1 - k)) < 1e-12
    # momentum buffers: scattered from per-parameter state into the flat buffer the fused SGD uses
    st = student.store
    mom = runner._pending_mom
    assert mom.shape == (st.n_train,)
    for spec, p in student.trainable_parameters()[:5] + student.trainable_parameters()[-5:]:
        o, n = st.offsets[spec.name]
        ref_p = dict(src.named_parameters())[spec.name]
        assert torch.equal(mom[o:o + n].view(p.shape), ref_opt.state[ref_p]["momentum_buffer"]), spec.name


def test_emaownhook_is_an_mmcv_hook_when_mmcv_is_importable():
    """With an importable `mmcv.runner.Hook` (here: the stub's shell) the EMA hook derives from it, so mmcv's own
    `register_hook` assertion `isinstance(hook, Hook)` holds (mmdet/runner/hooks/ema.py:1-6). Own process: the base
    class is chosen when dsl_b200.hooks is first imported."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from oracle import mmcv_stub; mmcv_stub.install()\n"
            "import mmcv.runner as mr\n"
            "from dsl_b200 import hooks as H\n"
            "assert H.Hook is mr.Hook and issubclass(H.EMAOWNHook, mr.Hook)\n"
            "h = H.EMAOWNHook(interval=1, mode='iteration', ratio=0.99, start_point=1)\n"
            "assert isinstance(h, mr.Hook) and h.ratio == 0.99 and h.every_n_iters(type('R', (), dict(iter=3))(), 2)\n"
            "print('ok')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def _expected_lr(it, epoch, base=0.01, warm=500, ratio=1.0 / 3, steps=(20, 26), gamma=0.1):
    """Closed form of mmcv StepLrUpdaterHook (by_epoch) + linear warm-up, lr_updater.py (1.3.x)."""
    reg = base * gamma ** sum(1 for s in steps if epoch >= s)
    if it < warm:
        return reg * (1 - (1 - it / warm) * (1 - ratio))
    return reg


def test_step_lr_hook_matches_mmcv_closed_form_over_the_whole_schedule():
    """28 epochs x 40 iterations through the restated hook: every iteration's LR equals the closed form (warm-up 500
    iterations at 1/3, x0.1 at epochs 20 and 26; configs/fcos_semi/RLA_*.py:188-194), for both param groups."""
    from dsl_b200 import hooks as H
    w = [torch.nn.Parameter(torch.zeros(1)), torch.nn.Parameter(torch.zeros(1))]
    opt = torch.optim.SGD([dict(params=[w[0]]), dict(params=[w[1]], lr=0.02)], lr=0.01, momentum=0.9)
    hook = H.build_hook(dict(type="StepLrUpdaterHook", warmup="linear", warmup_iters=500, warmup_ratio=1.0 / 3,
                             step=[20, 26]))
    r = types.SimpleNamespace(optimizer=opt, epoch=0, iter=0, data_loader=[0] * 40)
    hook.before_run(r)
    for epoch in range(28):
        r.epoch = epoch
        hook.before_train_epoch(r)
        for _ in range(40):
            hook.before_train_iter(r)
            e = _expected_lr(r.iter, epoch)
            assert abs(opt.param_groups[0]["lr"] - e) < 1e-15 and abs(opt.param_groups[1]["lr"] - 2 * e) < 1e-15
            r.iter += 1
    assert abs(opt.param_groups[0]["lr"] - 1e-4) < 1e-18


def test_runner_hook_priorities_and_optimizer_config_variants():
    from dsl_b200 import hooks as H
    from dsl_b200.runner import SemiEpochBasedRunner
    m, e = _build(), _build()
    r = SemiEpochBasedRunner(m, logger=logging.getLogger("t"), max_epochs=1, ema_model=e)
    assert r.fused_ema == dict(ratio=0.99, start_point=1)
    r.register_training_hooks(dict(policy="step", step=[1]), H.OptimizerHook(grad_clip=dict(max_norm=10, norm_type=2)),
                              dict(interval=2, mode="epoch", ratio=0.9, start_point=1), dict(interval=1),
                              dict(interval=5, hooks=[dict(type="TextLoggerHook")]), None)
    assert r.max_grad_norm == 10.0
    assert r.fused_ema is None                         # epoch-mode EMA: not inside the step, hook calls runner.EMA()
    assert [type(h).__name__ for h in r.hooks] == ["StepLrUpdaterHook", "EMAOWNHook", "CheckpointHook", "IterTimerHook",
                                                    "TextLoggerHook"]
    r.register_optimizer_hook(dict(grad_clip=None))
    assert r.max_grad_norm is None
    with pytest.raises(NotImplementedError):
        r.register_optimizer_hook(dict(type="Fp16OptimizerHook", loss_scale=512.))
    with pytest.raises(NotImplementedError):
        r.register_optimizer_hook(dict(grad_clip=dict(max_norm=1, norm_type=1)))
    with pytest.raises(NotImplementedError):
        r.register_momentum_hook(dict(policy="cyclic"))
    with pytest.raises(KeyError):
        r.register_custom_hooks([dict(type="NoSuchHook")])
    with pytest.raises(TypeError):
        r.register_hook(object())


def test_checkpoint_round_trip_both_optimizer_layouts(tmp_path):
    """save_checkpoint -> load_checkpoint / resume on the host side (no engine): student + teacher from one file as the
    reference does, or the teacher from `<file>_ema` on request; the flat optimizer layout of optimizer=None runners."""
    from dsl_b200.runner import SemiEpochBasedRunner
    m, e = _build(), _build()
    with torch.no_grad():
        m.store.flat.normal_(0, 0.02)
        e.store.flat.normal_(0, 0.02)
    r = SemiEpochBasedRunner(m, logger=logging.getLogger("t"), max_epochs=3, ema_model=e, meta=dict(seed=7))
    r._epoch, r._iter = 1, 20
    r.engine = types.SimpleNamespace(mom=torch.arange(m.store.n_train, dtype=torch.float32), lr=0.01, momentum=0.9,
                                     wd=1e-4)
    monkey_sync = torch.cuda.synchronize
    torch.cuda.synchronize = lambda *a, **k: None      # host-side round trip on a GPU-less box
    try:
        path = r.save_checkpoint(str(tmp_path))
    finally:
        torch.cuda.synchronize = monkey_sync
    assert os.path.basename(path) == "epoch_2.pth" and os.path.islink(str(tmp_path / "latest.pth"))
    m2, e2 = _build(), _build()
    r2 = SemiEpochBasedRunner(m2, logger=logging.getLogger("t"), max_epochs=3, ema_model=e2)
    r2.resume(path)
    assert (r2.epoch, r2.iter) == (2, 20) and r2.meta["hook_msgs"] == {}
    assert _same_weights(m2, m) and _same_weights(e2, m)   # same file for both
    assert torch.equal(r2._pending_mom, r.engine.mom)
    r3 = SemiEpochBasedRunner(_build(), logger=logging.getLogger("t"), max_epochs=3, ema_model=_build())
    r3.resume(path, ema_checkpoint=path + "_ema")
    assert _same_weights(r3.ema_model, e) and _same_weights(r3.model, m)


# ------------------------------------------------------------------------------------------------------------ GPU twins
@pytest.mark.gpu
def test_fused_runner_lr_scalar_follows_the_schedule_for_600_iterations():
    """The LR the captured SGD kernels apply: the device scalar after every iteration of 600 (warm-up 500 + 100) and
    across the step epochs equals the closed form of mmcv's StepLrUpdaterHook; and the update really scales with it."""
    from dsl_b200.runner import SemiEpochBasedRunner
    cfg = {k: v for k, v in MODEL_CFG.items() if k != "type"}
    from dsl_b200 import plugin
    model, ema = plugin.FCOS(**cfg).cuda(), plugin.FCOS(**cfg).cuda()
    ema.load_state_dict(model.state_dict())
    opt = torch.optim.SGD([p for _, p in model._trainable], lr=0.01, momentum=0.9, weight_decay=1e-4)
    runner = SemiEpochBasedRunner(model, optimizer=opt, logger=logging.getLogger("t"), max_epochs=30, ema_model=ema)
    runner.register_training_hooks(dict(policy="step", warmup="linear", warmup_iters=500, warmup_ratio=1.0 / 3,
                                        step=[20, 26]), dict(grad_clip=dict(max_norm=35, norm_type=2)),
                                   dict(interval=1, mode="iteration", ratio=0.99, start_point=1), None, None, None,
                                   timer_config=None)
    seen = []

    class Probe:
        def after_train_iter(self, r):
            seen.append((r.iter, r.epoch, float(r.engine.lr_scale.item())))

    runner.register_hook(Probe())
    loader = [_batch(2, 64, 96, 3 + i) for i in range(20)]
    runner.run([loader], [("train", 1)])                    # 30 epochs x 20 iterations = 600
    assert len(seen) == 600
    for it, ep, scale in seen:
        assert abs(scale - np.float32(_expected_lr(it, ep) / 0.01)) < 1e-7, (it, ep, scale)
    assert abs(seen[-1][2] - 0.01) < 1e-8 and abs(seen[0][2] - 1.0 / 3) < 1e-7
    # the applied update scales with the scalar: first step (zero momentum) from identical states at two LR factors
    deltas = []
    for factor in (1.0, 0.25):
        m2 = plugin.FCOS(**cfg).cuda()
        with torch.no_grad():
            m2.store.flat.copy_(model.store.flat)
        m2._dirty()
        o2 = torch.optim.SGD([p for _, p in m2._trainable], lr=0.01, momentum=0.9, weight_decay=1e-4)
        r2 = SemiEpochBasedRunner(m2, optimizer=o2, logger=logging.getLogger("t"), max_epochs=1)
        for g in o2.param_groups:
            g["initial_lr"] = 0.01
            g["lr"] = 0.01 * factor
        before = m2.store.flat.clone()
        r2.run([[loader[0]]], [("train", 1)])
        torch.cuda.synchronize()
        deltas.append((m2.store.flat - before)[:m2.store.n_train].double())
    ratio = (deltas[1].norm() / deltas[0].norm()).item()
    assert abs(ratio - 0.25) < 1e-4, ratio


@pytest.mark.gpu
def test_runner_closes_the_teacher_student_loop_on_the_device():
    """UnlabelPredHook's iteration mode (unlabel_pred_hook.py:455-469, 512-562) without the JSON round trip: from the
    start point on, the unlabeled image of iteration i + 1 is labelled by the teacher's detections of iteration i —
    hook gate + per-class NMS + dataset rule (oracle.hook_pseudo_labels), carried into the strong view by the batch's
    img_metas (oracle.view_boxes) — bit-exact in the student's target buffers; with the scale-invariant extra sample
    (semi_epoch_based_runner.py:186-204) appended on the device."""
    from dsl_b200 import plugin
    from dsl_b200.runner import SemiEpochBasedRunner
    from oracle import fcos_oracle as O
    cfg = {k: v for k, v in MODEL_CFG.items() if k != "type"}
    cfg["bbox_head"] = dict(cfg["bbox_head"], loss_weight=3.0, soft_weight=1.0, soft_warm_up=1)
    model, ema = plugin.FCOS(**cfg).cuda(), plugin.FCOS(**cfg).cuda()
    with torch.no_grad():
        model.store["bbox_head.conv_cls.bias"][:3] = 1.5         # three confident classes: scores ~0.8 x centerness
    model._dirty()
    ema.load_state_dict(model.state_dict())
    runner = SemiEpochBasedRunner(model, logger=logging.getLogger("t"), max_epochs=1, ema_model=ema, scale_invariant=True)
    runner.enable_device_pseudo_labels(start_point=0, lag=1, num_unlabeled=1)
    B, H, W = 2, 128, 160
    rng = np.random.RandomState(7)
    loader = []
    for i in range(4):
        b = _batch(B, H, W, 30 + i)
        # strong view of the unlabeled image (last of the batch): Resize 0.5 of a 256 x 320 original, PatchShuffle, flip
        b["img_metas"][1].update(img_shape=(H - 4 * i, W - 6, 3), scale_factor=np.array([0.48 + 0.01 * i, 0.5] * 2, np.float32),
                                 flip=bool(i % 2), flip_direction="horizontal", PS=bool(i >= 2), PS_mode="flip" if i == 2 else "flop",
                                 PS_place=0.3 + 0.1 * i)
        from tests.golden import inputs as GI
        b["teacher_img"] = GI.make_tensor(rng, 1, 3, H, W, scale=50.0)
        b["teacher_img_metas"] = [dict(img_shape=(H, W - 8, 3), scale_factor=np.array([0.5] * 4, np.float32),
                                       ori_shape=(2 * H, 2 * (W - 8), 3))]
        loader.append(b)
    rec = []

    class Probe:
        def after_train_iter(self, r):
            torch.cuda.synchronize()
            e = r.engine
            st = e.student
            rec.append(dict(dets=e.post.results(), thr=e.post.thr_class.cpu().numpy(),
                            gt=st.gt_boxes.cpu().numpy().copy(), gl=st.gt_labels.cpu().numpy().copy(),
                            go=st.gt_off.cpu().numpy().copy(), ig=st.ig_boxes.cpu().numpy().copy(),
                            io=st.ig_off.cpu().numpy().copy(), loss=r.outputs["log_vars"]["loss"]))

    runner.register_hook(Probe())
    runner.run([loader], [("train", 1)])
    from dsl_b200.geometry import view_from_meta
    assert len(rec) == 4 and all(np.isfinite(x["loss"]) for x in rec)
    n_pl = 0
    for i in range(1, 4):
        (dets, labels), = rec[i - 1]["dets"]                      # the teacher's detections one iteration earlier
        assert len(dets) > 5, "the confident teacher must detect something"
        gt, gl, ig = O.hook_pseudo_labels(dets.numpy(), labels.numpy(), 80, 2 * (W - 8), 2 * H, rec[i - 1]["thr"])
        v = view_from_meta(loader[i]["img_metas"][1])
        kw = dict(sx=np.float32(v.sx), sy=np.float32(v.sy), img_w=v.img_w, img_h=v.img_h, clip=bool(v.clip),
                  ps_mode=v.ps_mode, ps_crop=v.ps_crop, flip=bool(v.flip))
        wb, wl = O.view_boxes(gt, gl, **kw)
        wi, _ = O.view_boxes(ig, None, **kw)
        r = rec[i]
        go, io = r["go"].tolist(), r["io"].tolist()
        nl = len(loader[i]["gt_bboxes"][0])
        assert go[1] == nl and np.array_equal(r["gt"][:nl], loader[i]["gt_bboxes"][0].numpy())     # labeled image: dataloader
        assert np.array_equal(r["gt"][go[1]:go[2]], wb) and np.array_equal(r["gl"][go[1]:go[2]], wl), i
        assert np.array_equal(r["ig"][io[1]:io[2]], wi), i
        # scale-invariant extra sample: the unlabeled image's lists, halved
        assert go[3] - go[2] == go[2] - go[1] and np.array_equal(r["gt"][go[2]:go[3]], wb / 2)
        assert np.array_equal(r["gl"][go[2]:go[3]], wl) and np.array_equal(r["ig"][io[2]:io[3]], wi / 2)
        n_pl += len(wb) + len(wi)
    assert n_pl > 0, "pseudo GT / ignore boxes must have reached the student"
    # burn-in: iteration 0 used the dataloader's boxes for both images
    go0 = rec[0]["go"].tolist()
    assert go0[2] - go0[1] == len(loader[0]["gt_bboxes"][1])
