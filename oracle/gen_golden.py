"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE in place
(oracle/ref_loader.py) on the deterministic inputs of tests/golden/inputs.py. Run in the build container only
(`python -m oracle.gen_golden`); the GPU box never sees /root/reference, it reads the committed .npz files.

Files written (all small; inputs are regenerated from seeds, only reference OUTPUTS are stored):
  head_fwd.npz   FCOSHead.forward train+eval on a 64-channel, 2-conv, GN(8) head (weights from seed)
  loss_*.npz     FCOSHead.loss: labels / bbox_targets (bit-exact contract), losses, input-gradient samples
  backbone.npz   ResNet-50 (caffe, frozen BN) + FPN forward on a 1x3x64x96 input (weights from seed)
  rla_backbone.npz  RLA_ResNet (the shipped configs' backbone) forward + sampled parameter gradients, 1x3x64x96
  rla_detector.npz  build_detector(shipped RLA model dict).forward_train: losses + sampled parameter gradients
  decode.npz     FCOSHead.get_bboxes (teacher decode + score gate + NMS) on random head outputs
  view_image.npz  pixel side of the view pipelines (Resize / PatchShuffle / RandomFlip / Normalize / Pad) on small images
  view_draws.npz  the random draws of the reference's Resize / PatchShuffle / RandomFlip under fixed seeds
  saved_files.npz  the per-image JSON files save_results2file wrote for the hook_chain.npz detections, verbatim
  misc.npz       parse_det_results / adathres / _parse_ann_info filter rule / EMA body
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from tests.golden import inputs as GI  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

HEAD_CFG = dict(
    num_classes=80, in_channels=256, stacked_convs=4, feat_channels=256, strides=[8, 16, 32, 64, 128],
    norm_on_bbox=True, centerness_on_reg=True, dcn_on_last_conv=False, center_sampling=True, conv_bias=True,
    loss_cls=dict(type="FocalLoss", use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0),
    loss_bbox=dict(type="GIoULoss", loss_weight=1.0),
    loss_centerness=dict(type="CrossEntropyLoss", use_sigmoid=True, loss_weight=1.0))  # configs/fcos_semi/*.py:22-46


def small_head(R, **over):
    cfg = dict(HEAD_CFG, in_channels=64, feat_channels=64, stacked_convs=2,
               norm_cfg=dict(type="GN", num_groups=8, requires_grad=True))
    cfg.update(over)
    return R.FCOSHead(**cfg)


class _Cfg(dict):
    __getattr__ = dict.get


def gen_head_fwd(R):
    torch.manual_seed(0)
    head = small_head(R)
    GI.fill_state_dict_(head.state_dict(), seed=11)
    rng = np.random.RandomState(12)
    H, W, B = 128, 160, 2
    feats = [GI.make_tensor(rng, B, 64, h, w) for (h, w) in GI.level_sizes(H, W)]
    out = {}
    for mode in ("train", "eval"):
        head.train(mode == "train")
        with torch.no_grad():
            cls, box, ctr = head(feats)
        for i in range(5):
            out[f"{mode}_cls{i}"] = cls[i].numpy()
            out[f"{mode}_box{i}"] = box[i].numpy()
            out[f"{mode}_ctr{i}"] = ctr[i].numpy()
    np.savez_compressed(os.path.join(OUT, "head_fwd.npz"), **out)


LOSS_CASES = {
    # name: (seed, B, H, W, head kwargs, gt kwargs)
    "base_b2": (21, 2, 256, 320, dict(), dict(with_ignore=False)),
    "dsl_b2": (22, 2, 256, 320, dict(loss_weight=3.0), dict(with_ignore=True)),
    "dsl_b3_si": (23, 3, 256, 320, dict(loss_weight=3.0, soft_weight=1.0, soft_warm_up=5000), dict(with_ignore=True)),
    "empty_gt": (24, 2, 256, 320, dict(loss_weight=3.0), dict(with_ignore=True, empty_first=True)),
    "tie_break": (25, 2, 256, 320, dict(), dict(with_ignore=False, duplicate_boxes=True)),
    "many_gt": (26, 2, 384, 512, dict(loss_weight=3.0), dict(with_ignore=True, max_gt=40, max_ignore=8)),
    "ragged_hw": (27, 4, 200 // 8 * 8, 264, dict(loss_weight=3.0), dict(with_ignore=True)),
}


def gen_loss(R):
    for name, (seed, B, H, W, hk, gk) in LOSS_CASES.items():
        head = small_head(R, **hk)
        head.train()
        cls, box, ctr = GI.make_head_outputs(seed, B, H, W, train=True)
        for t in cls + box + ctr:
            t.requires_grad_(True)
        gts, labels, ignores = GI.make_gt(seed + 1000, B, H, W, **gk)
        metas = [dict(img_shape=(H, W, 3), pad_shape=(H, W, 3), scale_factor=1.0) for _ in range(B)]
        losses = head.loss(cls, box, ctr, gts, labels, metas, gt_bboxes_ignore=ignores)
        total = sum(losses.values())
        total.backward()
        # targets, straight from the reference's get_targets
        pts = head.get_points([c.shape[-2:] for c in cls], torch.float32, "cpu")
        lab, tgt = head.get_targets(pts, gts, labels)
        out = {k: np.float64(v.item()) for k, v in losses.items()}
        out["labels"] = torch.cat(lab).numpy().astype(np.int16)
        out["bbox_targets"] = torch.cat(tgt).numpy()
        if ignores is not None:
            ig_lab = [torch.zeros(b.size(0), dtype=torch.int64) + 79 for b in ignores]
            il, _ = head.get_targets(pts, ignores, ig_lab)
            out["ig_labels"] = torch.cat(il).numpy().astype(np.int16)
        # gradient fingerprints: full grads of bbox/ctr, strided sample + sum of the (large) cls grad
        g_cls = torch.cat([c.grad.permute(0, 2, 3, 1).reshape(-1, 80) for c in cls])
        out["dcls_sample"] = g_cls.reshape(-1)[::17].numpy()
        out["dcls_abs_sum"] = np.float64(g_cls.abs().double().sum().item())
        out["dbox"] = torch.cat([b.grad.permute(0, 2, 3, 1).reshape(-1, 4) for b in box]).numpy()
        out["dctr"] = torch.cat([c.grad.permute(0, 2, 3, 1).reshape(-1) for c in ctr]).numpy()
        np.savez_compressed(os.path.join(OUT, f"loss_{name}.npz"), **out)


def gen_backbone(R):
    torch.manual_seed(0)
    bb = R.ResNet(depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                  norm_cfg=dict(type="BN", requires_grad=False), norm_eval=True, style="caffe")
    neck = R.FPN(in_channels=[256, 512, 1024, 2048], out_channels=256, start_level=1,
                 add_extra_convs="on_output", num_outs=5, relu_before_extra_convs=True)
    GI.fill_state_dict_(bb.state_dict(), seed=31)
    GI.fill_state_dict_(neck.state_dict(), seed=32)
    bb.eval()
    neck.eval()
    x = GI.make_tensor(np.random.RandomState(33), 1, 3, 64, 96)
    with torch.no_grad():
        cs = bb(x)
        ps = neck(cs)
    out = {f"c{i + 2}": c.numpy() for i, c in enumerate(cs)}
    out.update({f"p{i + 3}": p.numpy() for i, p in enumerate(ps)})
    np.savez_compressed(os.path.join(OUT, "backbone.npz"), **out)


RLA_GRAD_KEYS = ("conv_outs.1.weight", "recurrent_convs.2.weight", "stage_bns.1.0.weight", "stage_bns.2.3.bias",
                 "stages.1.0.conv1.weight", "stages.2.0.bn1.weight", "stages.2.0.downsample.1.weight",
                 "stages.3.2.bn3.bias", "stages.3.0.conv2.weight", "stages.1.3.bn2.weight")


def gen_rla_backbone(R):
    """RLA_ResNet (mmdet/models/backbones/resnet_rla.py), the reference's own class, train() mode with norm_eval=True and
    frozen_stages=1 as in configs/fcos_semi/RLA_*.py:3-13: stage outputs on a 1x3x64x96 input, which parameters are
    trainable, and gradients of a weighted sum of the outputs for a sample of parameters. `flops=True` only moves the
    initial state tensor to the CPU (:296-300)."""
    RLA = ref_loader.load_rla()
    m = RLA(layers=[3, 4, 6, 3], frozen_stages=1, norm_eval=True, style="pytorch")
    m.flops = True
    sd = GI.rla_state_dict(51)
    missing = m.load_state_dict(sd, strict=False)
    assert all(k.endswith("num_batches_tracked") for k in missing.missing_keys) and not missing.unexpected_keys
    m.train()
    x = GI.make_tensor(np.random.RandomState(52), 1, 3, 64, 96)
    cs = m(x)
    out = {f"c{i + 2}": c.detach().numpy() for i, c in enumerate(cs)}
    rng = np.random.RandomState(53)
    ws = [torch.from_numpy(rng.randn(*c.shape).astype(np.float32)) for c in cs]
    sum((c * w).sum() for c, w in zip(cs, ws)).backward()
    params = dict(m.named_parameters())
    out["trainable"] = np.array(sorted(n for n, p in params.items() if p.requires_grad))
    out["no_grad"] = np.array(sorted(n for n, p in params.items() if p.requires_grad and p.grad is None))
    for k in RLA_GRAD_KEYS:
        g = params[k].grad.reshape(-1)
        out["grad:" + k] = (g[::97] if g.numel() > 4096 else g).numpy()   # large tensors: every 97th element
    np.savez_compressed(os.path.join(OUT, "rla_backbone.npz"), **out)


RLA_DET_GRAD_KEYS = GI.RLA_DET_GRAD_KEYS


def rla_detector_cfg():
    """Model dict of the shipped config (configs/fcos_semi/RLA_r50_caffe_mslonger_tricks_0.Xdata_unlabel_dynamic_lw_
    nofuse_iterlabel_lowfilter_singlestage.py:1-62), `pretrained` dropped (no checkpoint offline)."""
    return dict(
        type="FCOS",
        backbone=dict(type="RLA_ResNet", layers=[3, 4, 6, 3], frozen_stages=1, norm_eval=True, style="pytorch"),
        neck=dict(type="FPN", in_channels=[256, 512, 1024, 2048], out_channels=256, start_level=1,
                  add_extra_convs="on_output", num_outs=5, relu_before_extra_convs=True),
        bbox_head=dict(type="FCOSHead", loss_weight=3.0, soft_weight=1.0, soft_warm_up=5000, **HEAD_CFG),
        train_cfg=dict(assigner=dict(type="MaxIoUAssigner", pos_iou_thr=0.5, neg_iou_thr=0.4, min_pos_iou=0,
                                     ignore_iof_thr=-1), allowed_border=-1, pos_weight=-1, debug=False),
        test_cfg=dict(nms_pre=1000, min_bbox_size=0, score_thr=0.05, nms=dict(type="nms", iou_threshold=0.5),
                      max_per_img=100))


def gen_rla_detector(R):
    """The reference's FCOS detector built by its own build_detector from the shipped RLA config's model dict:
    forward_train (single_stage.py:152-180 -> FCOSHead.loss) on 2x3x256x320 with GT + ignore boxes; losses and sampled
    parameter gradients of the summed loss."""
    ref_loader.load_rla()
    m = R.builder.build_detector(rla_detector_cfg())
    m.backbone.flops = True                       # initial state tensor on the CPU (resnet_rla.py:296-300)
    sd = m.state_dict()
    mine = GI.rla_detector_state(61, 62)
    assert set(mine) == {k for k in sd if not k.endswith("num_batches_tracked")}
    with torch.no_grad():
        for k, v in mine.items():
            sd[k].copy_(v)
    m.train()
    B, H, W = 2, 256, 320
    img = GI.make_tensor(np.random.RandomState(63), B, 3, H, W)
    gts, labels, ignores = GI.make_gt(64, B, H, W, max_gt=9, max_ignore=3, with_ignore=True)
    metas = [dict(img_shape=(H, W, 3), pad_shape=(H, W, 3), scale_factor=1.0, filename=f"{i}.jpg") for i in range(B)]
    losses = m.forward_train(img, metas, gts, labels, gt_bboxes_ignore=ignores)
    out = {k: np.float64(v.item()) for k, v in losses.items()}
    sum(losses.values()).backward()
    params = dict(m.named_parameters())
    for k in RLA_DET_GRAD_KEYS:
        g = params[k].grad.reshape(-1)
        out["grad:" + k] = (g[::97] if g.numel() > 4096 else g).numpy()
    np.savez_compressed(os.path.join(OUT, "rla_detector.npz"), **out)


def gen_decode(R):
    head = small_head(R, test_cfg=_Cfg(nms_pre=1000, min_bbox_size=0, score_thr=0.05,
                                       nms=dict(type="nms", iou_threshold=0.6), max_per_img=100))
    head.eval()
    B, H, W = 2, 512, 640
    cls, box, ctr = GI.make_head_outputs(41, B, H, W, train=False, cls_mean=-6.5)
    metas = [dict(img_shape=(500, 630, 3), scale_factor=np.array([1.25, 1.25, 1.25, 1.25], dtype=np.float32)),
             dict(img_shape=(512, 600, 3), scale_factor=np.array([0.8, 0.8, 0.8, 0.8], dtype=np.float32))]
    with torch.no_grad():
        res = head.get_bboxes(cls, box, ctr, metas, rescale=True)
        raw = head.get_bboxes(cls, box, ctr, metas, rescale=True, with_nms=False)
    out = {}
    for b, (dets, labels) in enumerate(res):
        out[f"dets{b}"] = dets.numpy()
        out[f"labels{b}"] = labels.numpy()
    for b, (bx, sc, cn) in enumerate(raw):
        out[f"raw_boxes{b}"] = bx.numpy()
        out[f"raw_scores_max{b}"] = sc.max(-1)[0].numpy()
        out[f"raw_ctr{b}"] = cn.numpy()
    np.savez_compressed(os.path.join(OUT, "decode.npz"), **out)


def _extract_method(path, cls_name, fn_name, glb):
    import ast
    tree = ast.parse(open(path).read())
    for n in tree.body:
        if isinstance(n, ast.ClassDef) and n.name == cls_name:
            for m in n.body:
                if isinstance(m, ast.FunctionDef) and m.name == fn_name:
                    mod = ast.Module(body=[m], type_ignores=[])
                    exec(compile(mod, path, "exec"), glb)
                    return glb[fn_name]
    raise KeyError(fn_name)


def gen_misc(R):
    out = {}
    parse_det_results, adathres = ref_loader.load_hook_functions()
    rng = np.random.RandomState(51)
    # (1) hook gate: bbox2result-style per-class arrays -> kept boxes (score >= 0.1, int() truncation)
    per_class = []
    for c in range(5):
        n = int(rng.randint(0, 6))
        b = GI.demo_boxes(rng, n, 480, 640) + rng.rand(n, 4).astype(np.float32)
        s = rng.rand(n, 1).astype(np.float32) * 0.5
        per_class.append(np.concatenate([b, s], 1))
    kept = parse_det_results(per_class, 0.1)
    out["gate_in"] = np.concatenate([np.concatenate([p, np.full((len(p), 1), c, np.float32)], 1)
                                     for c, p in enumerate(per_class)])
    out["gate_bbox"] = np.array([k["bbox"] for k in kept], dtype=np.int64).reshape(-1, 4)
    out["gate_score"] = np.array([k["score"] for k in kept], dtype=np.float64)
    out["gate_cls"] = np.array([k["category_index"] for k in kept], dtype=np.int64)

    # (2) adathres: write per-image JSONs, run the reference twice (first pass, then with history)
    with tempfile.TemporaryDirectory() as td:
        cats = [f"cat{i}" for i in range(6)]
        cat2id = {c: i for i, c in enumerate(cats)}
        id2cat = {str(i): c for i, c in enumerate(cats)}
        id2cat[str(len(cats))] = "背景"   # the converter appends a background entry (tools/coco_convert2_semicoco_json.py:47-48)
        files = []
        allsc = {c: [] for c in cats}
        for i in range(12):
            n = int(rng.randint(0, 7))
            tags = [cats[int(rng.randint(0, 6))] for _ in range(n)]
            scores = [float(np.round(rng.rand() * 0.9 + 0.1, 6)) for _ in range(n)]
            for t, s in zip(tags, scores):
                allsc[t].append(s)
            json.dump(dict(targetNum=n, tags=tags, scores=scores), open(os.path.join(td, f"im{i}.jpg.json"), "w"))
            files.append(f"x/im{i}.jpg\n")
        fn = os.path.join(td, "adathres.json")
        adathres(0, True, fn, id2cat, cat2id, files, td, {})
        first = json.load(open(fn))
        adathres(0, True, fn, id2cat, cat2id, files, td, {})
        second = json.load(open(fn))
        out["ada_scores_json"] = np.frombuffer(json.dumps(allsc).encode(), dtype=np.uint8)
        out["ada_first_json"] = np.frombuffer(json.dumps(first).encode(), dtype=np.uint8)
        out["ada_second_json"] = np.frombuffer(json.dumps(second).encode(), dtype=np.uint8)

        # (3) dataset filter rule: SemiCOCODataset._parse_ann_info on a synthetic per-image JSON
        import types
        glb = {"os": os, "json": json, "np": np}
        parse_ann = _extract_method(os.path.join(ref_loader.REF_ROOT, "mmdet/datasets/semicoco.py"),
                                    "SemiCOCODataset", "_parse_ann_info", glb)
        n = 40
        rects = (GI.demo_boxes(rng, n, 480, 640) + np.array([-30, -30, 30, 30], np.float32)).tolist()
        rects[3] = [10.0, 10.0, 10.5, 50.0]      # w < 1 -> dropped
        rects[4] = [700.0, 10.0, 720.0, 50.0]    # no overlap with the image -> dropped
        scores = [float(np.round(rng.rand() * 0.5, 4)) for _ in range(n)]
        tags = [cats[int(rng.randint(0, 6))] for _ in range(n)]
        json.dump(dict(targetNum=n, rects=rects, scores=scores, tags=tags),
                  open(os.path.join(td, "a.jpg.json"), "w"))
        thr_file = os.path.join(td, "thr.json")
        json.dump(dict(thres={"cat0": 0.33, "cat1": 0.31, "cat2": 0.35}), open(thr_file, "w"))
        for tag, thres in (("nofile", os.path.join(td, "missing.json")), ("file", thr_file), ("fixed", [0.1, 0.4]),
                           ("none", None)):
            slf = types.SimpleNamespace(ann_path=td, thres=thres, default_thres=[0.1, 0.3], thres_list_by_class={},
                                        labelmapper=dict(cat2id=cat2id))
            ann = parse_ann(slf, dict(filename="a.jpg", width=640, height=480), None)
            out[f"filt_{tag}_gt"] = ann["bboxes"]
            out[f"filt_{tag}_labels"] = ann["labels"]
            out[f"filt_{tag}_ignore"] = ann["bboxes_ignore"]
        out["filt_rects"] = np.array(rects, dtype=np.float64)
        out["filt_scores"] = np.array(scores, dtype=np.float64)
        out["filt_cls"] = np.array([cat2id[t] for t in tags], dtype=np.int64)

    # (4) EMA body (semi_epoch_based_runner.py:392-406), executed on two tiny state dicts via the same expression
    import ast
    src = open(os.path.join(ref_loader.REF_ROOT, "mmdet/runner/hooks/semi_epoch_based_runner.py")).read()
    assert "student_model_dict[key] * (1 - keep_rate) + value * keep_rate" in src  # the line we restate
    s = {"w": GI.make_tensor(rng, 7, 3), "bn.running_mean": GI.make_tensor(rng, 7),
         "bn.num_batches_tracked": torch.tensor(5)}
    t = {"w": GI.make_tensor(rng, 7, 3), "bn.running_mean": GI.make_tensor(rng, 7),
         "bn.num_batches_tracked": torch.tensor(9)}
    keep_rate = 0.99
    new = {k: s[k] * (1 - keep_rate) + v * keep_rate for k, v in t.items()}
    for k in s:
        out["ema_s_" + k] = s[k].numpy()
        out["ema_t_" + k] = t[k].numpy()
        out["ema_new_" + k] = new[k].numpy()
    _ = ast
    np.savez_compressed(os.path.join(OUT, "misc.npz"), **out)


def gen_hook_chain(R):
    np.savez_compressed(os.path.join(OUT, "hook_chain.npz"), **hook_chain_cases(R, 77, 6))


def hook_chain_cases(R, seed, ncase, keep_json=False):
    """Detections (multiclass_nms output order) -> bbox2result -> UnlabelPredHook.save_results2file (JSON on disk) ->
    SemiCOCODataset._parse_ann_info: the reference's whole pseudo-label rule chain, executed from its own source.
    Returns the dict the golden file stores (seed 77, 6 cases); the tests also run other seeds live."""
    import types
    save_results2file = ref_loader.load_hook_chain()
    C = 6
    cats = [f"cat{i}" for i in range(C)]
    cat2id = {c: i for i, c in enumerate(cats)}
    id2cat = {str(i): c for i, c in enumerate(cats)}
    id2cat[str(len(cats))] = "背景"   # the converter appends a background entry (tools/coco_convert2_semicoco_json.py:47-48)
    glb = {"os": os, "json": json, "np": np}
    parse_ann = _extract_method(os.path.join(ref_loader.REF_ROOT, "mmdet/datasets/semicoco.py"),
                                "SemiCOCODataset", "_parse_ann_info", glb)
    out = {}
    rng = np.random.RandomState(seed)
    Wi, Hi = 640, 480
    for k in range(ncase):
        n = int(rng.randint(0, 60)) if k else 100
        boxes = GI.demo_boxes(rng, n, Hi, Wi) + rng.rand(n, 4).astype(np.float32)
        if n > 10:  # near-duplicates of the same class (second NMS must fire), a thin box, an outside box
            boxes[5] = boxes[4] + np.array([0.3, 0.2, 0.4, 0.1], np.float32)
            boxes[7] = np.array([10.2, 10.7, 10.9, 80.3], np.float32)
            boxes[8] = np.array([700.5, 10.0, 720.0, 50.0], np.float32)
        scores = np.sort(rng.rand(n).astype(np.float32) * 0.6)[::-1].copy()
        labels = rng.randint(0, C, size=n).astype(np.int64)
        if n > 10:
            labels[5] = labels[4]
            scores[9] = scores[8]  # tie
        dets = torch.from_numpy(np.concatenate([boxes, scores[:, None]], 1))
        result = R.bbox2result(dets, torch.from_numpy(labels), C)
        with tempfile.TemporaryDirectory() as td:
            root = os.path.join(td, "images")
            anno = os.path.join(td, "anno")
            save = os.path.join(td, "save")
            os.makedirs(os.path.join(root, "sub"))
            os.makedirs(os.path.join(anno, "sub"))
            json.dump(dict(imageName="sub/a.jpg", targetNum=0, rects=[], tags=[], masks=[], scores=[]),
                      open(os.path.join(anno, "sub", "a.jpg.json"), "w"))
            save_results2file(result, os.path.join(root, "sub", "a.jpg"), Hi, Wi, "json", "ckpt", 0.1, id2cat, cat2id,
                              root, save, "Det", anno_root_path=anno, iou=0.6, fuse=False, first_ignore=False)
            thr_file = os.path.join(td, "thr.json")
            thr = {"cat0": 0.33, "cat1": 0.31, "cat2": 0.35, "cat3": 0.3}
            json.dump(dict(thres=thr), open(thr_file, "w"))
            slf = types.SimpleNamespace(ann_path=os.path.join(save, "sub"), thres=thr_file, default_thres=[0.1, 0.3],
                                        thres_list_by_class={}, labelmapper=dict(cat2id=cat2id))
            ann = parse_ann(slf, dict(filename="a.jpg", width=Wi, height=Hi), None)
            if keep_json:   # the per-image file exactly as the reference's hook wrote it
                out[f"c{k}_saved_json"] = np.frombuffer(open(os.path.join(save, "sub", "a.jpg.json"), "rb").read(),
                                                        dtype=np.uint8)
        out[f"c{k}_dets"] = dets.numpy()
        out[f"c{k}_labels"] = labels
        out[f"c{k}_gt"] = ann["bboxes"]
        out[f"c{k}_gt_labels"] = ann["labels"]
        out[f"c{k}_ignore"] = ann["bboxes_ignore"]
    out["thr"] = np.array([0.33, 0.31, 0.35, 0.3, 0.3, 0.3], dtype=np.float64)  # missing classes -> default 0.3
    out["meta"] = np.array([ncase, C, Wi, Hi], dtype=np.int64)
    return out


def gen_saved_files(R):
    """The per-image JSON files UnlabelPredHook.save_results2file wrote for the hook_chain.npz detections (same seed),
    kept verbatim: pins dsl_b200/formats.py's writer and the device-side export of the saved list."""
    full = hook_chain_cases(R, 77, 6, keep_json=True)
    keep = {k: v for k, v in full.items()
            if (k.endswith(("_saved_json", "_dets", "_labels")) and not k.endswith("_gt_labels")) or k == "meta"}
    np.savez_compressed(os.path.join(OUT, "saved_files.npz"), **keep)


def gen_adathres_chain(R):
    """Detections -> UnlabelPredHook.save_results2file (one JSON per image) -> adathres() twice (first pass without a
    history file, second pass gated by the first pass's thresholds): the reference's per-epoch adaptive-threshold
    statistics on the very files its own hook wrote, for the on-device accumulation (dslb_pseudo_labels_stats +
    dslb_adathres_finalize)."""
    save_results2file = ref_loader.load_hook_chain()
    _, adathres = ref_loader.load_hook_functions()
    C = 6
    cats = [f"cat{i}" for i in range(C)]
    cat2id = {c: i for i, c in enumerate(cats)}
    id2cat = {str(i): c for i, c in enumerate(cats)}
    id2cat[str(len(cats))] = "背景"   # the converter appends a background entry (tools/coco_convert2_semicoco_json.py:47-48)
    out = {}
    rng = np.random.RandomState(91)
    Wi, Hi = 640, 480
    ncase = 8
    with tempfile.TemporaryDirectory() as td:
        root = os.path.join(td, "images")
        anno = os.path.join(td, "anno")
        save = os.path.join(td, "save")
        os.makedirs(os.path.join(root, "sub"))
        os.makedirs(os.path.join(anno, "sub"))
        files = []
        for k in range(ncase):
            n = int(rng.randint(5, 80))
            boxes = GI.demo_boxes(rng, n, Hi, Wi) + rng.rand(n, 4).astype(np.float32)
            boxes[3] = boxes[2] + np.array([0.3, 0.2, 0.4, 0.1], np.float32)   # duplicate: the hook's NMS must fire
            scores = np.sort(rng.rand(n).astype(np.float32) * 0.9)[::-1].copy()
            # skewed class frequencies: thresholds of frequent classes leave the lower clip (cum_c > mean count)
            labels = rng.choice(C, size=n, p=[0.45, 0.25, 0.12, 0.08, 0.05, 0.05]).astype(np.int64)
            labels[3] = labels[2]
            dets = torch.from_numpy(np.concatenate([boxes, scores[:, None]], 1))
            result = R.bbox2result(dets, torch.from_numpy(labels), C)
            name = f"im{k}.jpg"
            json.dump(dict(imageName="sub/" + name, targetNum=0, rects=[], tags=[], masks=[], scores=[]),
                      open(os.path.join(anno, "sub", name + ".json"), "w"))
            save_results2file(result, os.path.join(root, "sub", name), Hi, Wi, "json", "ckpt", 0.1, id2cat, cat2id,
                              root, save, "Det", anno_root_path=anno, iou=0.6, fuse=False, first_ignore=False)
            files.append("sub/" + name + "\n")
            out[f"c{k}_dets"] = dets.numpy()
            out[f"c{k}_labels"] = labels
        fn = os.path.join(td, "adathres.json")
        for tag in ("first", "second"):
            adathres(0, True, fn, id2cat, cat2id, files, os.path.join(save, "sub"), {})
            res = json.load(open(fn))
            thr = np.full(C, np.nan)
            wgt = np.full(C, np.nan)
            for c, v in res["thres"].items():
                thr[cat2id[c]] = v
            for c, v in res["id"].items():
                wgt[int(c)] = v
            out[f"{tag}_thr"] = thr
            out[f"{tag}_weight"] = wgt
    out["meta"] = np.array([ncase, C, Wi, Hi], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "adathres_chain.npz"), **out)


def gen_loss_modules(R):
    """The reference's own LOSSES modules (FocalLoss -> py_sigmoid_focal_loss on CPU, GIoULoss,
    CrossEntropyLoss(use_sigmoid=True)) on random inputs: value + gradient w.r.t. the prediction for every
    reduction / avg_factor / weight combination the modules accept."""
    out = {}
    rng = np.random.RandomState(123)
    N, C, n = 257, 11, 131
    logits = torch.from_numpy((rng.randn(N, C) * 2).astype(np.float32))
    labels = torch.from_numpy(rng.randint(0, C + 1, size=N).astype(np.int64))   # C = background
    wN = torch.from_numpy((rng.rand(N) * (rng.rand(N) > 0.2)).astype(np.float32))
    b1 = torch.from_numpy(GI.demo_boxes(rng, n, 300, 400) + rng.rand(n, 4).astype(np.float32))
    b2 = torch.from_numpy(GI.demo_boxes(rng, n, 300, 400) + rng.rand(n, 4).astype(np.float32))
    b2[:7] = b1[:7]                                    # identical boxes (tie gradients)
    b2[7:12] = b1[7:12] + 500.0                        # disjoint boxes
    wn = torch.from_numpy(rng.rand(n).astype(np.float32))
    ctr = torch.from_numpy((rng.randn(n) * 2).astype(np.float32))
    ctr_t = torch.from_numpy(rng.rand(n).astype(np.float32))
    out.update(logits=logits.numpy(), labels=labels.numpy(), wN=wN.numpy(), b1=b1.numpy(), b2=b2.numpy(), wn=wn.numpy(),
               ctr=ctr.numpy(), ctr_t=ctr_t.numpy())
    cases = [("mean_w_avg", dict(weight=True, avg_factor=37.5, reduction_override=None)),
             ("mean_now", dict(weight=False, avg_factor=None, reduction_override=None)),
             ("sum_w", dict(weight=True, avg_factor=None, reduction_override="sum")),
             ("none_w", dict(weight=True, avg_factor=None, reduction_override="none"))]
    mods = [("focal", R.FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0), logits, labels, wN),
            ("giou", R.GIoULoss(loss_weight=1.0), b1, b2, wn),
            ("bce", R.CrossEntropyLoss(use_sigmoid=True, loss_weight=1.0), ctr, ctr_t, wn),
            ("focal_lw", R.FocalLoss(use_sigmoid=True, gamma=1.5, alpha=0.4, loss_weight=2.5), logits, labels, wN)]
    for mname, mod, pred, tgt, w in mods:
        for cname, kw in cases:
            x = pred.clone().requires_grad_(True)
            val = mod(x, tgt, weight=w if kw["weight"] else None, avg_factor=kw["avg_factor"],
                      reduction_override=kw["reduction_override"])
            g = torch.from_numpy(rng.rand(*val.shape).astype(np.float32)) if val.dim() else torch.tensor(1.7)
            (val * g).sum().backward()
            out[f"{mname}_{cname}_val"] = val.detach().numpy()
            out[f"{mname}_{cname}_gout"] = g.numpy()
            out[f"{mname}_{cname}_grad"] = x.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "loss_modules.npz"), **out)


def _load_pipeline_classes():
    """Resize / RandomFlip / PatchShuffle (+ get_bbox_fields) compiled from the reference's own
    mmdet/datasets/pipelines/transforms.py, without importing the module's heavy dependencies: mmcv is answered by a
    two-function stub (imcrop with inclusive corners, as mmcv.imcrop), the registry decorator by the identity."""
    import ast
    import random
    import types
    import cv2
    path = os.path.join(ref_loader.REF_ROOT, "mmdet/datasets/pipelines/transforms.py")
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name == "get_bbox_fields") or
            (isinstance(n, ast.ClassDef) and n.name in ("Resize", "RandomFlip", "PatchShuffle"))]
    mm = types.SimpleNamespace(
        imcrop=lambda img, b: img[int(b[1]):int(b[3]) + 1, int(b[0]):int(b[2]) + 1].copy(),
        is_list_of=lambda seq, t: isinstance(seq, list) and all(isinstance(x, t) for x in seq),
        is_tuple_of=lambda seq, t: isinstance(seq, tuple) and all(isinstance(x, t) for x in seq))
    reg = types.SimpleNamespace(register_module=lambda *a, **k: (lambda c: c))
    glb = {"np": np, "random": random, "mmcv": mm, "cv2": cv2, "PIPELINES": reg}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), glb)
    return glb["Resize"], glb["RandomFlip"], glb["PatchShuffle"]


def _mmcv_image_stub():
    """The mmcv image functions the pipeline classes call, restated from their published definitions (mmcv 1.3.x,
    mmcv/image/geometric.py and photometric.py — mmcv itself is not installable here): thin cv2 / numpy wrappers."""
    import types
    import cv2

    def rescale_size(old_size, scale, return_scale=False):
        w, h = old_size
        if isinstance(scale, (float, int)):
            f = scale
        else:
            f = min(max(scale) / max(h, w), min(scale) / min(h, w))
        new = (int(w * float(f) + 0.5), int(h * float(f) + 0.5))
        return (new, f) if return_scale else new

    def imresize(img, size, return_scale=False, interpolation="bilinear", out=None, backend=None):
        assert interpolation == "bilinear" and backend in (None, "cv2")
        h, w = img.shape[:2]
        r = cv2.resize(img, size, dst=out, interpolation=cv2.INTER_LINEAR)
        return (r, size[0] / w, size[1] / h) if return_scale else r

    def imrescale(img, scale, return_scale=False, interpolation="bilinear", backend=None):
        h, w = img.shape[:2]
        new, f = rescale_size((w, h), scale, return_scale=True)
        r = imresize(img, new, interpolation=interpolation, backend=backend)
        return (r, f) if return_scale else r

    def imflip(img, direction="horizontal"):
        assert direction in ("horizontal", "vertical", "diagonal")
        return np.flip(img, axis={"horizontal": 1, "vertical": 0, "diagonal": (0, 1)}[direction])

    def imnormalize(img, mean, std, to_rgb=True):
        img = img.copy().astype(np.float32)
        assert img.dtype != np.uint8
        mean = np.float64(mean.reshape(1, -1))
        stdinv = 1 / np.float64(std.reshape(1, -1))
        if to_rgb:
            cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
        cv2.subtract(img, mean, img)
        cv2.multiply(img, stdinv, img)
        return img

    def impad(img, *, shape=None, padding=None, pad_val=0, padding_mode="constant"):
        assert shape is not None and padding is None and padding_mode == "constant"
        return cv2.copyMakeBorder(img, 0, shape[0] - img.shape[0], 0, shape[1] - img.shape[1], cv2.BORDER_CONSTANT,
                                  value=pad_val)

    def impad_to_multiple(img, divisor, pad_val=0):
        ph = int(np.ceil(img.shape[0] / divisor)) * divisor
        pw = int(np.ceil(img.shape[1] / divisor)) * divisor
        return impad(img, shape=(ph, pw), pad_val=pad_val)

    return types.SimpleNamespace(
        imcrop=lambda img, b: img[int(b[1]):int(b[3]) + 1, int(b[0]):int(b[2]) + 1].copy(),
        is_list_of=lambda seq, t: isinstance(seq, list) and all(isinstance(x, t) for x in seq),
        is_tuple_of=lambda seq, t: isinstance(seq, tuple) and all(isinstance(x, t) for x in seq),
        imresize=imresize, imrescale=imrescale, imflip=imflip, imnormalize=imnormalize, impad=impad,
        impad_to_multiple=impad_to_multiple)


def _load_image_pipeline():
    """Resize / PatchShuffle / RandomFlip / Normalize / Pad compiled from the reference's own transforms.py over
    _mmcv_image_stub (registry decorator = identity)."""
    import ast
    import random
    import types
    import cv2
    path = os.path.join(ref_loader.REF_ROOT, "mmdet/datasets/pipelines/transforms.py")
    tree = ast.parse(open(path).read())
    names = ("Resize", "RandomFlip", "PatchShuffle", "Normalize", "Pad")
    keep = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name == "get_bbox_fields") or
            (isinstance(n, ast.ClassDef) and n.name in names)]
    reg = types.SimpleNamespace(register_module=lambda *a, **k: (lambda c: c))
    glb = {"np": np, "random": random, "mmcv": _mmcv_image_stub(), "cv2": cv2, "PIPELINES": reg}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), glb)
    return {n: glb[n] for n in names}


IMG_NORM = dict(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True)   # shipped config :66-67


def view_image_cases(seed, ncase):
    """uint8 BGR images through the reference's own Resize(keep_ratio) -> PatchShuffle -> RandomFlip -> Normalize ->
    Pad(32) (the labeled / weak pipeline of the shipped config, :68-81), with the random draws pinned: the scale tuple,
    the PatchShuffle mode / place and the flip flag are chosen per case and handed to the classes the way their own
    random draws would set them. Returns inputs (source images, view parameters) and outputs (fp32 HWC images, metas)."""
    import random
    P = _load_image_pipeline()
    rng = np.random.RandomState(seed)
    out = {}
    views = []
    for k in range(ncase):
        h, w = int(rng.randint(37, 90)), int(rng.randint(37, 120))
        src = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        scale = [(160, 77), (133, 96), (90, 90), (200, 64)][k % 4]
        ps_mode = [None, "flip", "flop"][k % 3]
        place = float(rng.uniform(0.0, 1.0)) if k not in (4, 5) else (0.0 if k == 4 else 1.0)      # degenerate cuts
        flip = bool(k % 2)
        res = dict(img=src.copy(), img_shape=src.shape, ori_shape=src.shape, img_fields=["img"], bbox_fields=[],
                   scale=scale, flip=flip, flip_direction="horizontal" if flip else None)
        res = P["Resize"](img_scale=[(1333, 640), (1333, 800)], multiscale_mode="value", keep_ratio=True)(res)
        if ps_mode is not None:
            ps = P["PatchShuffle"](ratio=1.0, ranges=[place, place], mode=[ps_mode])
            np.random.seed(k)
            random.seed(k)
            res = ps(res)
            assert res["PS"] and res["PS_mode"] == ps_mode
        res = P["RandomFlip"](flip_ratio=0.5)(res)
        res = P["Normalize"](**IMG_NORM)(res)
        res = P["Pad"](size_divisor=32)(res)
        out[f"c{k}_src"] = src
        out[f"c{k}_out"] = np.ascontiguousarray(res["img"], dtype=np.float32)
        out[f"c{k}_scale_factor"] = res["scale_factor"]
        out[f"c{k}_img_shape"] = np.array(res["img_shape"], dtype=np.int64)
        views.append([scale[0], scale[1], {None: 0, "flip": 1, "flop": 2}[ps_mode], place, int(flip)])
    out["views"] = np.array(views, dtype=np.float64)
    out["meta"] = np.array([ncase], dtype=np.int64)
    return out


def gen_view_image(R):
    np.savez_compressed(os.path.join(OUT, "view_image.npz"), **view_image_cases(404, 6))


def view_draw_cases(seed, n):
    """The random draws of the reference's own Resize(multi-scale, 'value') -> PatchShuffle(0.5, [0, 1]) ->
    RandomFlip(0.5) (shipped config :70-77) over n passes with NumPy's and Python's global generators seeded once: rows of
    (src_h, src_w, img_h, img_w, PS, PS_mode [0 none / 1 flip / 2 flop], PS_place, flip), for geometry.draw_view."""
    import random
    P = _load_image_pipeline()
    rs = P["Resize"](img_scale=[(1333, 640), (1333, 800)], multiscale_mode="value", keep_ratio=True)
    ps = P["PatchShuffle"](ratio=0.5, ranges=[0.0, 1.0], mode=["flip", "flop"])
    fl = P["RandomFlip"](flip_ratio=0.5)
    shapes = np.random.RandomState(seed).randint(40, 90, size=(n, 2))
    np.random.seed(seed)
    random.seed(seed)
    rows = []
    for h, w in shapes:
        res = dict(img=np.zeros((int(h), int(w), 3), np.uint8), img_fields=["img"], bbox_fields=[])
        res = fl(ps(rs(res)))
        rows.append([h, w, res["img_shape"][0], res["img_shape"][1], float(res["PS"]),
                     {None: 0, "flip": 1, "flop": 2}[res["PS_mode"]], -1.0 if res["PS_place"] is None else res["PS_place"],
                     float(res["flip"])])
    return np.array(rows, dtype=np.float64)


def gen_view_draws(R):
    np.savez_compressed(os.path.join(OUT, "view_draws.npz"), rows=view_draw_cases(505, 40), seed=np.array([505]))


def gen_view_geometry(R):
    np.savez_compressed(os.path.join(OUT, "view_geometry.npz"), **view_cases(202, 14))


def view_cases(seed, ncase):
    """Boxes through the reference's Resize._resize_bboxes -> PatchShuffle.__call__ -> RandomFlip.bbox_flip (the order of
    the train pipelines, configs/fcos_semi/*.py:70-92) for a set of views: flip / flop cuts with boxes on either side of,
    straddling and touching the cut, degenerate cuts (no-op), flips, clipping, empty lists. Returns the dict the golden
    file stores (seed 202, 14 cases); tests/test_geometry.py also runs other seeds live where the reference is present."""
    import random
    Resize, RandomFlip, PatchShuffle = _load_pipeline_classes()
    out = {}
    rng = np.random.RandomState(seed)
    views = []
    for k in range(ncase):
        oh, ow = int(rng.randint(300, 700)), int(rng.randint(300, 900))
        sx, sy = np.float32(rng.uniform(0.6, 1.9)), np.float32(rng.uniform(0.6, 1.9))
        h, w = int(round(oh * float(sy))), int(round(ow * float(sx)))
        n = 0 if k == 5 else int(rng.randint(1, 40))
        boxes = GI.demo_boxes(rng, n, oh, ow).astype(np.float32) + rng.rand(n, 4).astype(np.float32) if n else \
            np.zeros((0, 4), np.float32)
        labels = rng.randint(0, 80, size=n).astype(np.int64)
        ign = GI.demo_boxes(rng, int(rng.randint(0, 6)), oh, ow).astype(np.float32)
        ps_mode = [None, "flip", "flop"][k % 3]
        place = float(rng.uniform(0.0, 1.0)) if k not in (7, 8) else (0.0 if k == 7 else 1.0)   # degenerate cuts
        flip = bool(k % 2)
        clip = k != 3
        res = dict(img=np.zeros((h, w, 3), np.uint8), img_shape=(h, w, 3), gt_bboxes=boxes.copy(), gt_labels=labels.copy(),
                   gt_bboxes_ignore=ign.copy(), bbox_fields=["gt_bboxes_ignore", "gt_bboxes"],
                   scale_factor=np.array([sx, sy, sx, sy], dtype=np.float32))
        rz = Resize.__new__(Resize)
        rz.bbox_clip_border = clip
        rz._resize_bboxes(res)
        crop = 0
        if ps_mode is not None:
            ps = PatchShuffle(ratio=1.0, ranges=[place, place], mode=[ps_mode])
            np.random.seed(k)       # np.random.rand(1) > ratio never holds with ratio 1.0; `seed` only scales a 0 range
            random.seed(k)
            res = ps(res)
            assert res["PS"] and res["PS_mode"] == ps_mode
            ext = w if ps_mode == "flip" else h
            crop = min(int(round(ext * res["PS_place"])), ext)
        if flip:
            rf = RandomFlip.__new__(RandomFlip)
            for key in ("gt_bboxes_ignore", "gt_bboxes"):
                if len(res[key]):
                    res[key] = rf.bbox_flip(res[key], (h, w, 3), "horizontal")
        out[f"c{k}_boxes"], out[f"c{k}_labels"], out[f"c{k}_ignore"] = boxes, labels, ign
        out[f"c{k}_out_boxes"] = np.asarray(res["gt_bboxes"], np.float32).reshape(-1, 4)
        out[f"c{k}_out_labels"] = np.asarray(res["gt_labels"], np.int64)
        out[f"c{k}_out_ignore"] = np.asarray(res["gt_bboxes_ignore"], np.float32).reshape(-1, 4)
        views.append([float(sx), float(sy), w, h, int(clip), {None: 0, "flip": 1, "flop": 2}[ps_mode], crop, int(flip)])
    out["views"] = np.array(views, dtype=np.float64)
    out["meta"] = np.array([ncase], dtype=np.int64)
    return out


def main():
    import sys
    torch.set_num_threads(8)
    R = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:   # regenerate only the named files, e.g. `python -m oracle.gen_golden adathres_chain`
        for name in sys.argv[1:]:
            globals()["gen_" + name](R)
        return
    gen_head_fwd(R)
    gen_loss(R)
    gen_backbone(R)
    gen_rla_backbone(R)
    gen_rla_detector(R)
    gen_decode(R)
    gen_misc(R)
    gen_hook_chain(R)
    gen_adathres_chain(R)
    gen_loss_modules(R)
    gen_view_geometry(R)
    gen_view_image(R)
    gen_saved_files(R)
    gen_view_draws(R)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
