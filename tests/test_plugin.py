"""The drop-in boundary (dsl_b200/plugin.py): config schema, registry keys, state_dict names — on CPU; and the
reference-style unit tests of the modules through the CUDA path — on the GPU."""
import os

import numpy as np
import pytest
import torch

from tests.golden import inputs as GI

# the model dict of configs/fcos_semi/r50_caffe_mslonger_tricks_0.Xdata.py:2-62 (baseline) with the DSL head keys of
# configs/fcos_semi/RLA_r50_..._singlestage.py:35-37 added
MODEL_CFG = dict(
    type="FCOS",
    backbone=dict(type="ResNet", depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                  norm_cfg=dict(type="BN", requires_grad=False), norm_eval=True, style="caffe",
                  init_cfg=dict(type="Pretrained", checkpoint="open-mmlab://detectron2/resnet50_caffe")),
    neck=dict(type="FPN", in_channels=[256, 512, 1024, 2048], out_channels=256, start_level=1,
              add_extra_convs="on_output", num_outs=5, relu_before_extra_convs=True),
    bbox_head=dict(type="FCOSHead", num_classes=80, in_channels=256, stacked_convs=4, feat_channels=256,
                   strides=[8, 16, 32, 64, 128], norm_on_bbox=True, centerness_on_reg=True, dcn_on_last_conv=False,
                   center_sampling=True, conv_bias=True, loss_weight=3.0, soft_weight=1.0, soft_warm_up=5000,
                   loss_cls=dict(type="FocalLoss", use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
                   loss_bbox=dict(type="GIoULoss", loss_weight=1.0),
                   loss_centerness=dict(type="CrossEntropyLoss", use_sigmoid=True, loss_weight=1.0)),
    train_cfg=dict(assigner=dict(type="MaxIoUAssigner", pos_iou_thr=0.5, neg_iou_thr=0.4, min_pos_iou=0,
                                 ignore_iof_thr=-1), allowed_border=-1, pos_weight=-1, debug=False),
    test_cfg=dict(nms_pre=1000, min_bbox_size=0, score_thr=0.05, nms=dict(type="nms", iou_threshold=0.5),
                  max_per_img=100))


def _build(cfg=MODEL_CFG):
    from dsl_b200 import plugin
    c = {k: v for k, v in cfg.items() if k != "type"}
    return plugin.FCOS(**c)


def test_plugin_builds_from_reference_config_schema():
    m = _build()
    sd = m.state_dict()
    n_all = sum(p.numel() for p in m.parameters())
    n_train = sum(p.numel() for p in m.parameters() if p.requires_grad)
    # ResNet-50 23 508 032 + FPN 3 868 672 + head 4 920 666
    assert n_all == 23508032 + 3868672 + 4920666
    assert n_train == n_all - sum(p.numel() for n, p in m.named_parameters() if not p.requires_grad)
    for k in ("backbone.conv1.weight", "backbone.layer4.2.bn3.running_var", "backbone.layer1.0.downsample.1.weight",
              "neck.lateral_convs.0.conv.bias", "neck.fpn_convs.4.conv.weight", "bbox_head.cls_convs.3.gn.weight",
              "bbox_head.conv_centerness.bias", "bbox_head.scales.4.scale",
              "backbone.bn1.num_batches_tracked"):
        assert k in sd, k
    assert sd["backbone.layer2.0.conv2.weight"].shape == (128, 128, 3, 3)   # OIHW fp32, as the reference stores it
    assert not dict(m.named_parameters())["backbone.layer1.0.conv1.weight"].requires_grad   # frozen_stages=1
    assert not dict(m.named_parameters())["backbone.layer3.0.bn1.weight"].requires_grad     # frozen BatchNorm
    assert dict(m.named_parameters())["backbone.layer2.0.conv1.weight"].requires_grad
    # parameters are views of ONE flat buffer: an optimizer-style in-place update shows up in the store
    p = dict(m.named_parameters())["bbox_head.conv_cls.bias"]
    with torch.no_grad():
        p.add_(1.0)
    assert torch.equal(m.store["bbox_head.conv_cls.bias"], p.detach())


def test_plugin_rejects_unsupported_options_loudly():
    from dsl_b200 import plugin
    bad = dict(MODEL_CFG, bbox_head=dict(MODEL_CFG["bbox_head"], dcn_on_last_conv=True))
    with pytest.raises(NotImplementedError):
        _build(bad)
    bad = dict(MODEL_CFG, backbone=dict(MODEL_CFG["backbone"], style="pytorch"))
    with pytest.raises(NotImplementedError):
        _build(bad)
    with pytest.raises(TypeError):
        plugin.FCOSHead(80, 256, centerness_on_reg=True, no_such_kwarg=1)
    m = _build()
    with pytest.raises(RuntimeError, match="no CPU"):   # no silent CPU fallback
        m.forward_train(torch.zeros(2, 3, 64, 64), [{}, {}], [torch.zeros(0, 4)] * 2, [torch.zeros(0).long()] * 2,
                        [torch.zeros(0, 4)] * 2)


def test_plugin_registers_under_reference_registry_keys():
    """With the reference's own registry loaded (source tree + mmcv stub; build container only), importing the plugin
    answers the keys the fcos_semi configs name, and build_detector(cfg.model) returns the B200 classes with the
    reference's state_dict names."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present (GPU box)")
    R = ref_loader.load()
    ref_model = R.builder.build_detector(dict(MODEL_CFG, backbone=dict(MODEL_CFG["backbone"], init_cfg=None)))
    ref_keys = {k: tuple(v.shape) for k, v in ref_model.state_dict().items()}
    from dsl_b200 import plugin
    reg = R.builder.DETECTORS        # one registry, aliased BACKBONES / NECKS / HEADS / LOSSES / DETECTORS (builder.py:6-14)
    names = ("FCOS", "FCOSHead", "ResNet", "FPN", "FocalLoss", "GIoULoss", "CrossEntropyLoss")
    originals = {n: reg.get(n) for n in names}
    keys = plugin.register(force=True)
    try:
        assert {"DETECTORS.FCOS", "HEADS.FCOSHead", "BACKBONES.ResNet", "NECKS.FPN", "LOSSES.FocalLoss",
                "LOSSES.GIoULoss", "LOSSES.CrossEntropyLoss"} <= set(keys), keys
        m = R.builder.build_detector(dict(MODEL_CFG))
        assert isinstance(m, plugin.FCOS)
        ours = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert ours == ref_keys
        # checkpoints travel both ways
        m.load_state_dict(ref_model.state_dict())
        assert torch.equal(m.store["bbox_head.conv_reg.weight"], ref_model.state_dict()["bbox_head.conv_reg.weight"])
        ref_model.load_state_dict(m.state_dict())
        # the standalone modules build from the reference's config dicts through the same registry, with its names
        from dsl_b200 import losses
        bb = R.builder.build_backbone(dict(MODEL_CFG["backbone"], init_cfg=None))
        nk = R.builder.build_neck(dict(MODEL_CFG["neck"]))
        assert isinstance(bb, plugin.ResNet) and isinstance(nk, plugin.FPN)
        assert set(bb.state_dict()) == {k[len("backbone."):] for k in ref_keys if k.startswith("backbone.")}
        assert set(nk.state_dict()) == {k[len("neck."):] for k in ref_keys if k.startswith("neck.")}
        assert isinstance(R.builder.build_loss(dict(MODEL_CFG["bbox_head"]["loss_cls"])), losses.FocalLoss)
        assert isinstance(R.builder.build_loss(dict(MODEL_CFG["bbox_head"]["loss_bbox"])), losses.GIoULoss)
        assert isinstance(R.builder.build_loss(dict(MODEL_CFG["bbox_head"]["loss_centerness"])), losses.CrossEntropyLoss)
    finally:   # put the reference's own classes back for the other tests
        for n, cls in originals.items():
            if cls is not None:
                reg.register_module(name=n, force=True, module=cls)


def test_c_abi_exports_every_declared_symbol():
    """libdslb.so loads without a GPU and exports every function include/dslb.h declares (no compute calls here)."""
    import ctypes
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "dslb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(dslb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) > 50, names
    lib = ctypes.CDLL(os.path.join(root, "dsl_b200", "libdslb.so"))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    lib.dslb_version.restype = ctypes.c_int
    assert lib.dslb_version() >= 100


@pytest.mark.gpu
@pytest.mark.parametrize("hw", [(64, 96), (70, 101), (33, 47)])
def test_scale_invariant_input_matches_oracle(hw):
    """SI extra input (semi_epoch_based_runner.py:186-204): the half-resolution kernel vs F.interpolate on the CPU."""
    from dsl_b200 import plugin
    from oracle import fcos_oracle as O
    H, W = hw
    rng = np.random.RandomState(0)
    img = GI.make_tensor(rng, 2, 3, H, W)
    gts, lbs, igs = GI.make_gt(3, 2, H, W, with_ignore=True)
    metas = [dict(img_shape=(H, W, 3), pad_shape=(H, W, 3), scale_factor=np.ones(4, np.float32))] * 2
    a = plugin.scale_invariant_input(img.cuda(), [g.cuda() for g in gts], lbs, [i.cuda() for i in igs], metas)
    b = O.scale_invariant_input(img, gts, igs)
    assert a[0].shape == b[0].shape
    assert torch.equal(a[0][:2].cpu(), b[0][:2])
    assert (a[0][2].cpu() - b[0][2]).abs().max().item() <= 1e-5      # fp32 bilinear weights, fma contraction may differ
    assert torch.equal(a[0][2, :, H // 2:].cpu(), torch.zeros(3, H - H // 2, W))    # zero padding outside the copy
    assert torch.equal(a[1][-1].cpu(), b[1]) and torch.equal(a[3][-1].cpu(), b[2])
    assert len(a[1]) == len(a[2]) == len(a[3]) == len(a[4]) == 3 and torch.equal(a[2][-1], lbs[-1])
    assert a[4][-1]["img_shape"][:2] == (H // 2, W // 2)


# ------------------------------------------------------------------------------------------------------- GPU
def _data(B, H, W, seed, dev="cuda"):
    rng = np.random.RandomState(seed)
    img = GI.make_tensor(rng, B, 3, H, W, scale=50.0).to(dev)
    gts, labels, ignores = GI.make_gt(seed + 1, B, H, W, with_ignore=True)
    metas = [dict(img_shape=(H, W, 3), pad_shape=(H, W, 3), scale_factor=np.ones(4, np.float32), filename=f"{i}.jpg")
             for i in range(B)]
    return dict(img=img, img_metas=metas, gt_bboxes=[g.to(dev) for g in gts], gt_labels=[l.to(dev) for l in labels],
                gt_bboxes_ignore=[i.to(dev) for i in ignores])


@pytest.mark.gpu
def test_plugin_train_step_backward_and_sgd():
    """detector.train_step -> loss.backward() -> torch.optim.SGD.step(), as tools/train.py drives it: losses match the
    fp32 oracle on the same weights (bf16 forward: 2e-2), parameter gradients arrive on the nn.Parameters, and an
    optimizer step on them changes what the kernels compute (parameters are views of the flat store)."""
    from oracle import fcos_oracle as O
    from tests.test_gpu_parity import _oracle_state
    m = _build().cuda()
    m.train()
    data = _data(2, 256, 320, 11)
    out = m.train_step(data, None)
    assert set(out) == {"loss", "log_vars", "num_samples"} and out["num_samples"] == 2
    assert set(out["log_vars"]) == {"loss_cls", "loss_bbox", "loss_centerness", "loss"}
    net = next(iter(m._nets.values()))
    bb, neck, head = _oracle_state(net)
    with torch.no_grad():
        ps = O.fpn_forward(neck, O.resnet_forward(bb, data["img"].cpu(), 50))
        cls, box, ctr = O.fcos_head_forward(head, ps, training=True)
        ref = O.fcos_loss(cls, box, ctr, [g.cpu() for g in data["gt_bboxes"]], [l.cpu() for l in data["gt_labels"]],
                          [i.cpu() for i in data["gt_bboxes_ignore"]], loss_weight=3.0)
    for k, v in ref.items():
        got = out["log_vars"][k]
        print(k, got, float(v))
        assert abs(got - float(v)) <= 2e-2 * abs(float(v)) + 1e-4
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=0.01, momentum=0.9)
    opt.zero_grad()
    out["loss"].backward()
    named = dict(m.named_parameters())
    g = named["bbox_head.conv_cls.weight"].grad
    assert g is not None and g.shape == (80, 256, 3, 3) and float(g.abs().sum()) > 0
    assert named["backbone.layer1.0.conv1.weight"].grad is None
    o, n = m.store.offsets["bbox_head.conv_cls.weight"]
    assert torch.equal(g.flatten(), net.grad[o:o + n])
    before = m.store["bbox_head.conv_cls.weight"].clone()
    opt.step()
    assert not torch.equal(before, m.store["bbox_head.conv_cls.weight"])
    out2 = m.train_step(data, None)
    assert out2["log_vars"]["loss"] != out["log_vars"]["loss"]


@pytest.mark.gpu
def test_plugin_head_reference_unit_test_properties():
    """The reference's own FCOSHead unit test (tests/test_models/test_dense_heads/test_fcos_head.py:7-63), restated for
    256 input channels: empty GT => cls loss > 0 and box loss == 0; one GT => both > 0. Plus get_targets bit-exact
    against the oracle and forward against the oracle on the same weights."""
    from dsl_b200 import plugin
    from oracle import fcos_oracle as O
    s = 256
    img_metas = [dict(img_shape=(s, s, 3), scale_factor=1, pad_shape=(s, s, 3))]
    head = plugin.FCOSHead(num_classes=16, in_channels=256, strides=(4, 8, 16, 32, 64), centerness_on_reg=True,
                           loss_cls=dict(type="FocalLoss", use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
                           loss_bbox=dict(type="GIoULoss", loss_weight=1.0)).cuda()
    head.train()
    g = torch.Generator().manual_seed(0)
    feat = [torch.rand(1, 256, s // f, s // f, generator=g).cuda() for f in (4, 8, 16, 32, 64)]
    cls_scores, bbox_preds, centerness = head.forward(feat)
    assert [tuple(c.shape) for c in cls_scores] == [(1, 16, s // f, s // f) for f in (4, 8, 16, 32, 64)]
    # forward vs the oracle on the same weights
    sd = {k: v.detach().cpu() for k, v in head.state_dict().items()}
    with torch.no_grad():
        rc, rb, rt = O.fcos_head_forward(sd, [f.cpu().bfloat16().float() for f in feat], strides=(4, 8, 16, 32, 64),
                                         training=True)
    for l in range(5):
        e = (cls_scores[l].cpu() - rc[l]).abs().max().item() / rc[l].abs().max().item()
        assert e < 1e-2, (l, e)
    empty = head.loss(cls_scores, bbox_preds, centerness, [torch.empty((0, 4)).cuda()], [torch.LongTensor([]).cuda()],
                      img_metas, None)
    assert empty["loss_cls"].item() > 0 and empty["loss_bbox"].item() == 0
    gt_b = [torch.Tensor([[23.6667, 23.8757, 238.6326, 151.8874]]).cuda()]
    gt_l = [torch.LongTensor([2]).cuda()]
    one = head.loss(cls_scores, bbox_preds, centerness, gt_b, gt_l, img_metas, None)
    assert one["loss_cls"].item() > 0 and one["loss_bbox"].item() > 0 and one["loss_centerness"].item() > 0
    # get_targets: bit-exact with the reference algorithm
    sizes = [tuple(c.shape[-2:]) for c in cls_scores]
    pts = head.get_points(sizes)
    labels, targets = head.get_targets(pts, gt_b, gt_l)
    rl, rt_ = O.get_targets(O.get_points(sizes, (4, 8, 16, 32, 64)), [b.cpu() for b in gt_b], [l.cpu() for l in gt_l],
                            (4, 8, 16, 32, 64), ((-1, 64), (64, 128), (128, 256), (256, 512), (512, 1e8)), 16,
                            center_sampling=False, norm_on_bbox=False)
    for l in range(5):
        assert torch.equal(labels[l].cpu(), rl[l]) and torch.equal(targets[l].cpu(), rt_[l])


@pytest.mark.gpu
def test_plugin_simple_test_and_ema_hook():
    """simple_test returns the reference's bbox2result structure; EMAOWNHook drives the fused EMA kernel bit-exactly."""
    import types
    from dsl_b200 import plugin
    from oracle import fcos_oracle as O
    student, teacher = _build().cuda(), _build().cuda()
    with torch.no_grad():
        student.store.flat.add_(torch.randn_like(student.store.flat) * 1e-3)
        teacher.store["bbox_head.conv_cls.bias"].fill_(-1.0)   # confident teacher: plenty of candidates
    teacher._dirty()
    teacher.eval()
    data = _data(2, 256, 320, 5)
    res = teacher.simple_test(data["img"], data["img_metas"], rescale=True)
    assert len(res) == 2 and len(res[0]) == 80 and all(r.shape[1] == 5 for r in res[0])
    assert sum(len(r) for r in res[0]) == 100    # max_per_img
    t0 = {k: v.detach().cpu().clone() for k, v in teacher.state_dict().items() if v.dtype.is_floating_point}
    s0 = {k: v.detach().cpu().clone() for k, v in student.state_dict().items() if v.dtype.is_floating_point}
    hook = plugin.EMAOWNHook(interval=1, mode="iteration", ratio=0.99, start_point=0)
    runner = types.SimpleNamespace(model=student, ema_model=teacher, iter=4, epoch=0, ema_flag=False)
    hook.after_train_iter(runner)
    ref = O.ema_update(t0, s0, 0.99)
    for k in ("backbone.layer2.0.conv1.weight", "bbox_head.conv_cls.bias", "backbone.bn1.running_var"):
        assert torch.equal(teacher.state_dict()[k].cpu(), ref[k]), k
    assert runner.ema_flag


@pytest.mark.gpu
def test_semi_epoch_based_runner_drives_the_fused_step(tmp_path):
    """RUNNERS['SemiEpochBasedRunner'] mirror: reference constructor / counters / hook stages / checkpoint names over the
    fused engine, with the scale-invariant extra input and the SI soft loss switched on as in the shipped DSL config."""
    import logging
    from dsl_b200 import plugin
    from dsl_b200.runner import SemiEpochBasedRunner
    cfg = {k: v for k, v in MODEL_CFG.items() if k != "type"}
    cfg["bbox_head"] = dict(cfg["bbox_head"], loss_weight=3.0, soft_weight=1.0, soft_warm_up=1)
    model, ema = plugin.FCOS(**cfg).cuda(), plugin.FCOS(**cfg).cuda()
    ema.load_state_dict(model.state_dict())
    opt = torch.optim.SGD([p for _, p in model._trainable], lr=0.01, momentum=0.9, weight_decay=1e-4)
    with pytest.raises(TypeError):
        SemiEpochBasedRunner(model, optimizer=opt, logger="not a logger", max_epochs=1)
    runner = SemiEpochBasedRunner(model, optimizer=opt, work_dir=str(tmp_path), logger=logging.getLogger("t"),
                                  meta=dict(seed=0), max_epochs=2, ema_model=ema, scale_invariant=True)
    runner.register_hook(plugin.EMAOWNHook(interval=1, mode="iteration", ratio=0.99, start_point=0))
    stages = []

    class Probe:
        def before_train_epoch(self, r): stages.append(("be", r.epoch))
        def after_train_iter(self, r): stages.append(("ai", r.iter, r.inner_iter))
        def after_train_epoch(self, r): stages.append(("ae", r.epoch))

    runner.register_hook(Probe())
    B, H, W = 2, 128, 160
    loader = []
    for i in range(2):
        loader.append(_data(B, H, W, 10 + i))
    t0 = ema.store.flat.clone()
    s0 = model.store.flat.clone()
    runner.run([loader], [("train", 1)])
    torch.cuda.synchronize()
    assert (runner.epoch, runner.iter, runner.max_iters) == (2, 4, 4)
    assert stages == [("be", 0), ("ai", 0, 0), ("ai", 1, 1), ("ae", 0), ("be", 1), ("ai", 2, 0), ("ai", 3, 1), ("ae", 1)]
    lv = runner.outputs["log_vars"]
    assert set(lv) == {"loss_cls", "loss_bbox", "loss_centerness", "loss_sisoft", "loss"} and np.isfinite(lv["loss"])
    assert runner.outputs["num_samples"] == B
    assert runner.engine.student.B == B + 1            # the scale-invariant extra image was appended on the device
    assert not torch.equal(model.store.flat, s0)       # SGD moved the student (plugin parameters are the same memory)
    assert not torch.equal(ema.store.flat, t0)         # ... and the EMA moved the teacher
    # frozen stem: student unchanged there, so the teacher's copy is unchanged too (0.01 s + 0.99 t with s == t)
    o, n = model.store.offsets["backbone.conv1.weight"]
    assert torch.allclose(ema.store.flat[o:o + n], t0[o:o + n], rtol=1e-6, atol=0)
    path = runner.save_checkpoint(str(tmp_path))
    assert os.path.basename(path) == "epoch_3.pth" and os.path.exists(path + "_ema")
    ck = torch.load(path, weights_only=False)
    assert ck["meta"]["epoch"] == 3 and ck["meta"]["iter"] == 4 and ck["meta"]["seed"] == 0
    assert "bbox_head.conv_cls.weight" in ck["state_dict"] and "backbone.layer4.2.bn3.running_var" in ck["state_dict"]
    assert int(runner.engine.post.stat_cnt.sum()) == 0 and runner.engine.post.have_prev   # adathres ran at epoch end


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.gpu
def test_standalone_resnet_and_fpn_modules_forward_backward():
    """BACKBONES['ResNet'] / NECKS['FPN'] as modules of their own (NCHW fp32 in / out, autograd-connected), chained like
    SingleStageDetector.extract_feat (single_stage.py:136-141): forward vs the fp32 oracle within the bf16 bar,
    parameter / input gradients vs the oracle's autograd (direction and norm; bf16 activations through <= 16 layers)."""
    from dsl_b200 import plugin
    from oracle import fcos_oracle as O
    torch.manual_seed(0)
    bb = plugin.ResNet(depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                       norm_cfg=dict(type="BN", requires_grad=False), norm_eval=True, style="caffe").cuda()
    neck = plugin.FPN(in_channels=[256, 512, 1024, 2048], out_channels=256, start_level=1, add_extra_convs="on_output",
                      num_outs=5, relu_before_extra_convs=True).cuda()
    with pytest.raises(NotImplementedError):
        plugin.ResNet(depth=50, style="pytorch", frozen_stages=1, norm_cfg=dict(type="BN", requires_grad=False))
    with pytest.raises(NotImplementedError):
        plugin.FPN(in_channels=[256, 512, 1024, 2048], out_channels=256, num_outs=5)     # start_level=0 layout
    B, H, W = 2, 128, 160
    rng = np.random.RandomState(4)
    x = GI.make_tensor(rng, B, 3, H, W, scale=50.0)
    cs = bb(x.cuda())
    ps = neck(cs)
    assert [tuple(c.shape) for c in cs] == [(B, 256, 32, 40), (B, 512, 16, 20), (B, 1024, 8, 10), (B, 2048, 4, 5)]
    assert [tuple(p.shape[1:]) for p in ps] == [(256, 16, 20), (256, 8, 10), (256, 4, 5), (256, 2, 3), (256, 1, 2)]
    # oracle (fp32, CPU) on the same weights
    sd_b = {k: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k)
            for k, v in bb.state_dict().items() if "num_batches" not in k}
    sd_n = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in neck.state_dict().items()}
    rc = O.resnet_forward(sd_b, x, 50)
    rp = O.fpn_forward(sd_n, rc)
    for got, ref in zip(list(cs) + list(ps), rc + rp):
        err = (got.detach().cpu() - ref.detach()).abs().max().item() / (ref.abs().max().item() + 1e-12)
        assert err < 3e-2, err
    ws = [torch.from_numpy(rng.randn(*p.shape).astype(np.float32)) for p in ps]
    sum((p * w.cuda()).sum() for p, w in zip(ps, ws)).backward()
    sum((p * w).sum() for p, w in zip(rp, ws)).backward()
    checked = 0
    for name, p in list(neck.named_parameters()):
        ref = sd_n[name].grad
        assert _cos(p.grad.cpu(), ref) > 0.999 and abs(p.grad.norm().item() / ref.norm().item() - 1) < 2e-2, name
        checked += 1
    for name, p in bb.named_parameters():
        if not p.requires_grad:
            assert p.grad is None and name.startswith(("conv1", "bn", "layer1")) or ".bn" in name or "downsample.1" in name
            continue
        ref = sd_b[name].grad
        c = _cos(p.grad.cpu(), ref)
        floor = 0.99 if name.startswith("layer4") else (0.97 if name.startswith("layer3") else 0.93)
        assert c > floor, (name, c)
        checked += 1
    assert checked > 50
    # SGD on the plugin parameters moves the flat store the kernels read (same memory)
    before = bb.store.flat.clone()
    torch.optim.SGD([p for p in bb.parameters() if p.requires_grad], lr=0.1).step()
    assert not torch.equal(bb.store.flat, before)


def test_ctypes_prototypes_match_the_header():
    """Every function bound in dsl_b200/_lib.py carries as many ctypes argtypes as include/dslb.h declares parameters
    (a mismatch would corrupt the call frame silently)."""
    import re
    from dsl_b200 import _lib as L
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "dslb.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    decl = {}
    for m in re.finditer(r"\b(dslb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = m.group(2).strip()
        decl[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    checked = 0
    for name, nargs in decl.items():
        fn = getattr(L.lib, name)
        if fn.argtypes is None:
            continue
        assert len(fn.argtypes) == nargs, (name, len(fn.argtypes), nargs)
        checked += 1
    assert checked > 45, checked


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of the ABI structs have the size and field offsets a C compiler gives include/dslb.h."""
    import ctypes
    import re
    import shutil
    import subprocess
    from dsl_b200 import _lib as L
    from dsl_b200.geometry import ImageView, View
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = {"dslb_conv_seg_t": L.ConvSeg, "dslb_wgrad_seg_t": L.WgradSeg, "dslb_gn_seg_t": L.GnSeg,
             "dslb_pack_desc_t": L.PackDesc, "dslb_unpack_desc_t": L.UnpackDesc, "dslb_fcos_level_t": L.FcosLevel,
             "dslb_view_t": View, "dslb_bn_grad_desc_t": L.BnGradDesc, "dslb_image_view_t": ImageView}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "dslb.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([cc, "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for cname, field, val in re.findall(r"(\w+) (\w+) (\d+)", out):
        cls = pairs[cname]
        if field == "size":
            assert ctypes.sizeof(cls) == int(val), (cname, ctypes.sizeof(cls), val)
        else:
            assert getattr(cls, field).offset == int(val), (cname, field)
        seen += 1
    assert seen > 80


# the model dict of the SHIPPED DSL config, configs/fcos_semi/RLA_r50_caffe_mslonger_tricks_0.Xdata_unlabel_dynamic_lw_
# nofuse_iterlabel_lowfilter_singlestage.py:1-62 — RLA_ResNet backbone (:3-13)
RLA_MODEL_CFG = dict(MODEL_CFG, backbone=dict(type="RLA_ResNet", layers=[3, 4, 6, 3], frozen_stages=1, norm_eval=True,
                                               style="pytorch", pretrained="/nonexistent/resnet50_rla_2283.pth.tar"))


def test_plugin_builds_the_shipped_rla_config():
    m = _build(RLA_MODEL_CFG)
    named = dict(m.named_parameters())
    n_bb = sum(p.numel() for n, p in named.items() if n.startswith("backbone."))
    assert n_bb == 23789632     # parameter count of the reference's RLA_ResNet([3, 4, 6, 3]) (tests/golden/rla_backbone.npz)
    sd = m.state_dict()
    for k in ("backbone.conv_outs.3.weight", "backbone.recurrent_convs.0.weight", "backbone.stage_bns.2.5.running_var",
              "backbone.stages.1.0.downsample.1.weight", "backbone.stages.0.0.conv1.weight",
              "backbone.stage_bns.3.2.num_batches_tracked"):
        assert k in sd, k
    assert sd["backbone.stages.2.1.conv1.weight"].shape == (256, 1024 + 32, 1, 1)   # conv1 acts on cat(x, h)
    # resnet_rla.py:344-377: stem + stage 1 frozen, stage_bns[3][2] frozen, every other BatchNorm affine TRAINABLE
    assert not named["backbone.stages.0.1.conv2.weight"].requires_grad
    assert not named["backbone.stage_bns.0.0.weight"].requires_grad and not named["backbone.conv_outs.0.weight"].requires_grad
    assert not named["backbone.stage_bns.3.2.weight"].requires_grad
    assert named["backbone.stages.1.0.bn1.weight"].requires_grad and named["backbone.stage_bns.3.1.bias"].requires_grad
    assert named["backbone.recurrent_convs.1.weight"].requires_grad
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "rla_backbone.npz"))
    assert sorted(n[len("backbone."):] for n, p in named.items() if n.startswith("backbone.") and p.requires_grad) == \
        list(g["trainable"])
    with pytest.raises(NotImplementedError):
        _build(dict(RLA_MODEL_CFG, backbone=dict(RLA_MODEL_CFG["backbone"], SE=True)))
    with pytest.raises(NotImplementedError):
        _build(dict(RLA_MODEL_CFG, backbone=dict(RLA_MODEL_CFG["backbone"], layers=[2, 2, 2, 2])))


def test_plugin_rla_registry_and_state_dict_names():
    """build_backbone / build_detector of the reference's registry return the B200 classes for `RLA_ResNet`, with the
    reference's state_dict names and shapes; checkpoints load both ways."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present (GPU box)")
    R = ref_loader.load()
    RefRLA = ref_loader.load_rla()
    ref_bb = RefRLA(layers=[3, 4, 6, 3], frozen_stages=1, norm_eval=True, style="pytorch")
    ref_keys = {k: tuple(v.shape) for k, v in ref_bb.state_dict().items()}
    from dsl_b200 import plugin
    reg = R.builder.DETECTORS
    names = ("FCOS", "FCOSHead", "ResNet", "RLA_ResNet", "FPN", "FocalLoss", "GIoULoss", "CrossEntropyLoss")
    originals = {n: reg.get(n) for n in names}
    keys = plugin.register(force=True)
    try:
        assert "BACKBONES.RLA_ResNet" in keys
        bb = R.builder.build_backbone(dict(RLA_MODEL_CFG["backbone"]))
        assert isinstance(bb, plugin.RLA_ResNet)
        assert {k: tuple(v.shape) for k, v in bb.state_dict().items()} == ref_keys
        bb.load_state_dict(ref_bb.state_dict())
        assert torch.equal(bb.store["stages.3.0.conv1.weight"], ref_bb.state_dict()["stages.3.0.conv1.weight"])
        ref_bb.load_state_dict(bb.state_dict())
        m = R.builder.build_detector(dict(RLA_MODEL_CFG))
        assert isinstance(m, plugin.FCOS) and m.backbone_kind == "rla"
        assert {k[len("backbone."):] for k in m.state_dict() if k.startswith("backbone.")} == set(ref_keys)
    finally:
        for n, cls in originals.items():
            if cls is not None:
                reg.register_module(name=n, force=True, module=cls)


def test_plugin_rla_pretrained_checkpoint_loads_into_the_backbone(tmp_path):
    """`pretrained` of the shipped config names an ImageNet RLA checkpoint: init_weights loads it non-strictly into
    backbone.* (classifier keys of the checkpoint are ignored, missing keys keep their initialisation)."""
    ck = {k: v for k, v in GI.rla_state_dict(5).items() if not k.startswith("stage_bns.3")}
    ck["fc.weight"] = torch.zeros(1000, 2048 + 32)          # ImageNet head of the checkpoint: not part of the detector
    path = tmp_path / "resnet50_rla.pth.tar"
    torch.save({"state_dict": ck}, path)
    m = _build(dict(RLA_MODEL_CFG, backbone=dict(RLA_MODEL_CFG["backbone"], pretrained=str(path))))
    m.init_weights()
    assert torch.equal(m.store["backbone.stages.2.3.conv1.weight"], ck["stages.2.3.conv1.weight"])
    assert torch.equal(m.store["backbone.conv_outs.1.weight"], ck["conv_outs.1.weight"])
    assert float(m.store["backbone.stage_bns.3.0.running_var"].min()) == 1.0      # missing in the checkpoint: init kept


def test_runner_keeps_one_engine_per_batch_shape_with_shared_optimizer_state(monkeypatch):
    """Multi-scale training (BASELINE configs[4]) changes the padded batch shape between iterations: the runner keeps one
    engine per (B, H, W) (LRU) and all of them share the momentum buffer, the LR scalar, the SI warm-up counter and the
    adaptive-threshold state. Host logic only: the engine is replaced by a stub, nothing touches the GPU."""
    import logging
    from types import SimpleNamespace
    from dsl_b200 import plugin, runner as R

    class FakeEngine:
        made = []

        def __init__(self, B, H, W, **kw):
            self.B, self.H, self.W, self.kw = B, H, W, kw
            self.mom, self.lr_scale = torch.zeros(4), torch.ones(1)
            self.cur_iter, self.graphs, self.lr, self.ema_keep = 0, "captured", kw["lr"], kw["ema_keep"]
            self.ema_in_step, self.max_grad_norm = True, kw["max_grad_norm"]
            self.teacher = SimpleNamespace(B=kw.get("teacher_B") or B)
            z = lambda dt: torch.zeros(80, dtype=dt)  # noqa: E731
            self.post = SimpleNamespace(stat_cnt=z(torch.int64), stat_cum=z(torch.float64), stat_prev=z(torch.float64),
                                        thr_class=z(torch.float64), class_weight=z(torch.float64), have_prev=False,
                                        cand_overflow=torch.zeros(1, dtype=torch.int32))
            for n in ("pl_gt_boxes", "pl_gt_labels", "pl_gt_off", "pl_ig_boxes", "pl_ig_off"):
                setattr(self, n, torch.zeros(4))
            FakeEngine.made.append(self)

        def set_inputs(self, *a, **k):
            self.cur_iter += 1

        def step(self):
            self.mom += 1
            self.post.stat_cnt += 1
            return dict(loss_cls=torch.tensor(1.0), loss_bbox=torch.tensor(0.5), loss_centerness=torch.tensor(0.25))

        def end_epoch(self):
            self.post.have_prev = True
            self.graphs = None

    monkeypatch.setattr(R, "DSLEngine", FakeEngine)
    m = _build()
    m.store.device = torch.device("cuda")          # the stub never dereferences it
    t = _build()
    t.store.device = torch.device("cuda")
    run = R.SemiEpochBasedRunner(m, logger=logging.getLogger("t"), max_epochs=1, ema_model=t)
    run.register_hook(plugin.EMAOWNHook(interval=1, mode="iteration", ratio=0.999, start_point=0))
    shapes = [(2, 128, 160), (2, 160, 128), (2, 128, 160), (2, 192, 160)]
    loader = [dict(img=torch.zeros(B, 3, H, W), img_metas=[dict(filename=f"{i}_{b}.jpg") for b in range(B)],
                   gt_bboxes=[torch.zeros(0, 4)] * B, gt_labels=[torch.zeros(0, dtype=torch.long)] * B,
                   gt_bboxes_ignore=[torch.zeros(0, 4)] * B) for i, (B, H, W) in enumerate(shapes)]
    run.train(loader)
    e = FakeEngine.made
    assert [(x.B, x.H, x.W) for x in e] == [(2, 128, 160), (2, 160, 128), (2, 192, 160)]      # the third batch re-used #0
    assert e[1].mom is e[0].mom and e[2].mom is e[0].mom and float(e[0].mom[0]) == 4.0           # ONE momentum buffer
    assert e[2].lr_scale is e[0].lr_scale and e[1].post.stat_cnt is e[0].post.stat_cnt
    assert int(e[0].post.stat_cnt[0]) == 4 and run._si_iter == 4 and all(x.ema_keep == 0.999 for x in e)
    assert run.engine is e[2] and all(x.post.have_prev and x.graphs is None for x in e)       # epoch end reached every shape
    assert e[0].kw["backbone"] == "resnet" and e[0].kw["student_store"] is m.store
    run.max_cached_shapes = 2
    run._engine_for(2, 224, 160)
    assert list(run._engines) == [(2, 192, 160), (2, 224, 160)] and FakeEngine.made[3].mom is e[0].mom


@pytest.mark.gpu
def test_runner_multi_shape_training_on_the_gpu():
    """Two padded shapes alternating (multi-scale training): two engines, one momentum buffer, weights keep moving, losses
    finite, the engine of the first shape is re-used (its CUDA graph replays against the shared state)."""
    import logging
    from dsl_b200 import plugin
    from dsl_b200.runner import SemiEpochBasedRunner
    cfg = {k: v for k, v in MODEL_CFG.items() if k != "type"}
    model, ema = plugin.FCOS(**cfg).cuda(), plugin.FCOS(**cfg).cuda()
    ema.load_state_dict(model.state_dict())
    runner = SemiEpochBasedRunner(model, logger=logging.getLogger("t"), max_epochs=1, ema_model=ema)
    loader = [_data(2, 128, 160, 20), _data(2, 160, 128, 21), _data(2, 128, 160, 22), _data(2, 160, 128, 23)]
    flats = []

    class Probe:
        def after_train_iter(self, r):
            torch.cuda.synchronize()
            flats.append(model.store.flat.clone())
            assert np.isfinite(r.outputs["log_vars"]["loss"])

    runner.register_hook(Probe())
    runner.run([loader], [("train", 1)])
    assert list(runner._engines) == [(2, 128, 160), (2, 160, 128)]
    e0, e1 = runner._engines.values()
    assert e0.mom is e1.mom and float(e0.mom.abs().sum()) > 0
    assert all(not torch.equal(a, b) for a, b in zip(flats, flats[1:]))      # every iteration stepped the same weights
    assert e0.post.have_prev and e1.post.have_prev and e0.graphs is None and e1.graphs is None


@pytest.mark.gpu
def test_plugin_simple_test_with_the_accurate_head_mode():
    """FCOS.set_eval_head_precision('bf16x3'): the registry module's inference path (simple_test -> bbox2result lists) on
    the split-bf16 head; detections stay those of the bf16 head up to score / box noise well inside the NMS rules."""
    m = _build().cuda()
    with torch.no_grad():
        m.store["bbox_head.conv_cls.bias"][:2] = 1.0        # two confident classes
    m._dirty()
    m.eval()
    data = _data(2, 256, 320, 17)
    base = m.simple_test(data["img"], data["img_metas"], rescale=False)
    m.set_eval_head_precision("bf16x3")
    acc = m.simple_test(data["img"], data["img_metas"], rescale=False)
    assert len(acc) == 2 and len(acc[0]) == 80
    for b in range(2):
        nb, na = sum(len(r) for r in base[b]), sum(len(r) for r in acc[b])
        assert na == 100 and nb == 100                       # max_per_img survivors in both modes
        top_b = max((r[:, 4].max() for r in base[b] if len(r)), default=0.0)
        top_a = max((r[:, 4].max() for r in acc[b] if len(r)), default=0.0)
        assert abs(float(top_a) - float(top_b)) < 2e-2 * float(top_b)
