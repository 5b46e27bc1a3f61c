"""dslb_view_images (the pixel side of the view pipelines) through the C ABI: bit-exact against
oracle/image_oracle.py and the reference golden (its own Resize / PatchShuffle / RandomFlip / Normalize / Pad classes).
The per-pixel arithmetic of the kernel is also pinned on CPU (tests/test_image_oracle.py compiles the same header for the
host); the host-side view construction is checked here without a GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import image_oracle as IO

G = os.path.join(os.path.dirname(__file__), "golden")
MEAN, STD = (123.675, 116.28, 103.53), (58.395, 57.12, 57.375)          # shipped config :66-67


def test_image_view_host_side_matches_oracle_meta():
    from dsl_b200 import geometry as GEO
    rng = np.random.RandomState(3)
    for k in range(50):
        h, w = int(rng.randint(8, 700)), int(rng.randint(8, 700))
        scale = [(1333, 800), (1333, 640), (int(rng.randint(40, 400)), int(rng.randint(20, 300)))][k % 3]
        mode, place, flip = int(rng.randint(0, 3)), float(rng.uniform()), bool(k % 2)
        v, meta = GEO.image_view((h, w), scale, ps_mode=mode, ps_place=place, flip=flip)
        nw, nh = IO.rescale_size(w, h, scale)
        assert (v.src_h, v.src_w, v.img_h, v.img_w) == (h, w, nh, nw) and meta["img_shape"] == (nh, nw, 3)
        ext = nw if mode == 1 else nh
        assert v.ps_crop == (min(int(round(ext * place)), ext) if mode else 0) and v.flip == int(flip)
        assert np.array_equal(meta["scale_factor"], np.array([nw / w, nh / h, nw / w, nh / h], dtype=np.float32))
        bv = GEO.view_from_meta(meta)                      # the box side reads the same meta
        assert (bv.img_w, bv.img_h, bv.ps_mode, bv.ps_crop, bv.flip) == (nw, nh, v.ps_mode, v.ps_crop, v.flip)
    assert GEO.image_view((480, 640), (1333, 800), "flop", 0.5)[0].ps_mode == 2


def _run(srcs, draws, to_rgb=True, mean=MEAN, std=STD, **kw):
    from dsl_b200 import geometry as GEO
    views = [GEO.image_view(s.shape[:2], sc, ps_mode=m, ps_place=p, flip=f)[0] for s, (sc, m, p, f) in zip(srcs, draws)]
    out = GEO.view_images([torch.from_numpy(np.ascontiguousarray(s)).cuda() for s in srcs], views, mean, std,
                          to_rgb=to_rgb, **kw)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _oracle_batch(srcs, draws, H, W, to_rgb=True, mean=MEAN, std=STD):
    ref = np.zeros((len(srcs), 3, H, W), dtype=np.float32)
    for b, (s, (sc, m, p, f)) in enumerate(zip(srcs, draws)):
        o, _ = IO.view_image(s, sc, m, p, f, mean=mean, std=std, to_rgb=to_rgb)
        ref[b, :, :o.shape[1], :o.shape[2]] = o
    return ref


@pytest.mark.gpu
def test_view_images_match_reference_golden():
    g = np.load(os.path.join(G, "view_image.npz"))
    n = int(g["meta"][0])
    srcs = [g[f"c{k}_src"] for k in range(n)]
    draws = [((int(v[0]), int(v[1])), int(v[2]), float(v[3]), bool(v[4])) for v in g["views"][:n]]
    for k in range(n):                                    # one image per launch: the reference's own padded shape
        out = _run(srcs[k:k + 1], draws[k:k + 1])
        assert np.array_equal(out[0], g[f"c{k}_out"].transpose(2, 0, 1)), k
    out = _run(srcs, draws)                               # the whole set as one ragged batch
    H, W = out.shape[2:]
    assert H % 32 == 0 and W % 32 == 0
    for k in range(n):
        r = g[f"c{k}_out"].transpose(2, 0, 1)
        assert np.array_equal(out[k, :, :r.shape[1], :r.shape[2]], r), k
        assert not out[k, :, r.shape[1]:].any() and not out[k, :, :, r.shape[2]:].any()


@pytest.mark.gpu
def test_view_images_random_views_vs_oracle():
    rng = np.random.RandomState(11)
    for it in range(6):
        B = int(rng.randint(1, 6))
        srcs = [rng.randint(0, 256, size=(int(rng.randint(8, 260)), int(rng.randint(8, 260)), 3)).astype(np.uint8)
                for _ in range(B)]
        draws = [((int(rng.randint(40, 400)), int(rng.randint(20, 300))), int(rng.randint(0, 3)),
                  float(rng.choice([0.0, 1.0, rng.uniform()])), bool(rng.randint(0, 2))) for _ in range(B)]
        to_rgb = bool(it % 2)
        mean, std = rng.uniform(90, 130, 3), rng.uniform(40, 70, 3)
        out = _run(srcs, draws, to_rgb=to_rgb, mean=mean, std=std)
        assert np.array_equal(out, _oracle_batch(srcs, draws, out.shape[2], out.shape[3], to_rgb, mean, std)), it


@pytest.mark.gpu
def test_view_images_coco_sized_batch_and_properties():
    """BASELINE batch shape: four COCO-sized sources -> (4, 3, 800, 1344); bit-exact vs the oracle, plus the size-
    independent properties: flip of a flip-free view == the view mirrored inside img_w, PatchShuffle == a cyclic roll."""
    rng = np.random.RandomState(5)
    shapes = [(480, 640), (427, 640), (480, 600), (375, 500)]      # landscape: every view fits 800 x 1344
    srcs = [rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8) for h, w in shapes]
    draws = [((1333, 800), 0, 0.0, False), ((1333, 800), 1, 0.37, True), ((1333, 640), 2, 0.81, False),
             ((1333, 800), 0, 0.0, True)]
    out = _run(srcs, draws, H=800, W=1344)
    assert out.shape == (4, 3, 800, 1344)
    assert np.array_equal(out, _oracle_batch(srcs, draws, 800, 1344))
    plain = _run(srcs, [((sc, 0, 0.0, False)) for sc, _, _, _ in draws], H=800, W=1344)
    from dsl_b200 import geometry as GEO
    for b, (sc, m, p, f) in enumerate(draws):
        v = GEO.image_view(srcs[b].shape[:2], sc, m, p, f)[0]
        want = plain[b, :, :v.img_h, :v.img_w]
        if v.ps_mode == 1:
            want = np.roll(want, -v.ps_crop, axis=2)
        elif v.ps_mode == 2:
            want = np.roll(want, -v.ps_crop, axis=1)
        if f:
            want = want[:, :, ::-1]
        assert np.array_equal(out[b, :, :v.img_h, :v.img_w], want), b


@pytest.mark.gpu
def test_view_images_reject_bad_arguments():
    from dsl_b200 import geometry as GEO
    src = torch.zeros(20, 30, 3, dtype=torch.uint8, device="cuda")
    v = GEO.image_view((20, 30), (64, 48))[0]
    with pytest.raises(ValueError):
        GEO.view_images([src.float()], [v], MEAN, STD)
    with pytest.raises(ValueError):
        GEO.view_images([src], [v], MEAN, STD, H=8, W=8)
    with pytest.raises(Exception):
        GEO.view_images([src], [v], MEAN, (1.0, 0.0, 1.0))


@pytest.mark.gpu
def test_engine_inputs_rendered_from_uint8_sources():
    """DSLEngine.set_images_from_sources writes both networks' static input buffers (incl. the scale-invariant layout,
    where the batch occupies the first B slots) with the pipeline's exact pixels, and the step runs on them."""
    from dsl_b200 import geometry as GEO
    from dsl_b200.trainer import DSLEngine
    from tests.golden import inputs as GI
    B, H, W = 2, 128, 160
    eng = DSLEngine(B, H, W, depth=50, seed=0, use_graphs=False, scale_invariant=True, soft_weight=1.0, soft_warm_up=5000,
                    teacher_B=1)           # the reference's literal mix, as bench.py --mix literal builds it
    rng = np.random.RandomState(2)
    srcs = [rng.randint(0, 256, size=(int(rng.randint(60, 140)), int(rng.randint(60, 140)), 3)).astype(np.uint8)
            for _ in range(B + 1)]
    draws = [((128, 128), 0, 0.0, False), ((120, 100), 1, 0.4, True), ((128, 96), 2, 0.7, False)]
    views = [GEO.image_view(s.shape[:2], sc, m, p, f)[0] for s, (sc, m, p, f) in zip(srcs, draws)]
    dsrcs = [torch.from_numpy(s).cuda() for s in srcs]
    eng.set_images_from_sources(dsrcs[:B], views[:B], dsrcs[B:], views[B:])
    gts, labels, ignores = GI.make_gt(1, B, H, W, with_ignore=True)
    eng.set_inputs(None, gts, labels, ignores)
    torch.cuda.synchronize()
    ref = _oracle_batch(srcs, draws, H, W)
    assert np.array_equal(eng.student.img[:B].cpu().numpy(), ref[:B])
    assert np.array_equal(eng.teacher.img.cpu().numpy(), ref[B:])
    losses = eng.step()
    torch.cuda.synchronize()
    assert all(np.isfinite(float(v)) for v in losses.values())


@pytest.mark.gpu
def test_adathres_state_round_trips_through_the_reference_file_format(tmp_path):
    """(kept in this file so that it runs at the end of the GPU suite) Device statistics -> dslb_adathres_finalize ->
    TeacherPost.adathres_state -> formats.save_adathres == the JSON the reference's adathres() writes for the same
    scores (oracle restatement, pinned on the reference's own file in tests/test_formats.py); load_adathres installs
    thresholds + history, with -inf for the classes the file does not list."""
    import json
    from dsl_b200 import formats as FM
    from dsl_b200.postprocess import TeacherPost
    from oracle import fcos_oracle as O
    C = 6
    cats = [f"cat{i}" for i in range(C)]
    post = TeacherPost(1, [(4, 4)], (8,), C, "cuda")
    rng = np.random.RandomState(9)
    by_class = {c: [float(s) for s in np.round(rng.rand(int(rng.randint(1, 40))) * 0.6 + 0.3, 6)] for c in (0, 1, 3, 4)}
    cnt, cum = np.zeros(C, np.int64), np.zeros(C, np.float64)
    for c, ss in by_class.items():
        cnt[c] = len(ss)
        for s in ss:                                       # the reference's left-to-right fp64 accumulation
            cum[c] += s
    post.stat_cnt.copy_(torch.from_numpy(cnt))
    post.stat_cum.copy_(torch.from_numpy(cum))
    post.adathres_update()
    torch.cuda.synchronize()
    thr, wgt = post.adathres_state()
    thres, weights = O.adathres(by_class)
    p = str(tmp_path / "adathres.json")
    FM.save_adathres(p, thr, wgt, cats)
    d = json.load(open(p))
    assert set(d["thres"]) == {cats[c] for c in by_class}
    for c in by_class:
        assert abs(d["thres"][cats[c]] - thres[c]) <= 1e-12 and abs(d["id"][str(c)] - weights[c]) <= 1e-12 * weights[c]
    post2 = TeacherPost(1, [(4, 4)], (8,), C, "cuda")
    post2.load_adathres(*FM.adathres_from_json(p, cats, absent_thr=0.3))
    torch.cuda.synchronize()
    assert post2.have_prev
    prev = post2.stat_prev.cpu().numpy()
    assert np.isneginf(prev[[2, 5]]).all() and np.array_equal(prev[[0, 1, 3, 4]], np.array(thr)[[0, 1, 3, 4]])
    assert np.array_equal(post2.thr_class.cpu().numpy(), np.where(np.isin(np.arange(C), [2, 5]), 0.3, np.array(thr)))


@pytest.mark.gpu
def test_saved_records_match_the_files_the_reference_hook_wrote():
    """dslb_pseudo_labels_saved -> TeacherPost.saved_records: the per-image JSON dicts are exactly the files the
    reference's UnlabelPredHook.save_results2file wrote for the same detections (golden saved_files.npz, verbatim), and
    the GT / ignore lists of the same launch still equal hook_chain.npz (the export does not disturb the rule chain)."""
    import json
    from dsl_b200.postprocess import TeacherPost
    g = np.load(os.path.join(G, "saved_files.npz"))
    h = np.load(os.path.join(G, "hook_chain.npz"))
    ncase, C, Wi, Hi = (int(v) for v in g["meta"])
    cats = [f"cat{i}" for i in range(C)]
    post = TeacherPost(ncase, [(8, 8)], (8,), C, "cuda", max_per_img=100)
    post.set_meta([(Hi, Wi, 3)] * ncase, None)
    post.set_class_thresholds(h["thr"])
    for k in range(ncase):
        d = torch.from_numpy(g[f"c{k}_dets"]).float()
        post.dets[k, :len(d)] = d.cuda()
        post.det_labels[k, :len(d)] = torch.from_numpy(g[f"c{k}_labels"]).int().cuda()
        post.det_count[k] = len(d)
    recs = post.saved_records(["sub/a.jpg"] * ncase, cats)
    for k in range(ncase):
        assert json.loads(json.dumps(recs[k])) == json.loads(bytes(g[f"c{k}_saved_json"]).decode()), k
    v = post._sv
    go, io = v["go"].cpu().tolist(), v["io"].cpu().tolist()
    for k in range(ncase):
        assert np.array_equal(v["gt"][go[k]:go[k + 1]].cpu().numpy(), h[f"c{k}_gt"].reshape(-1, 4)), k
        assert np.array_equal(v["gl"][go[k]:go[k + 1]].cpu().numpy(), h[f"c{k}_gt_labels"]), k
        assert np.array_equal(v["ig"][io[k]:io[k + 1]].cpu().numpy(), h[f"c{k}_ignore"].reshape(-1, 4)), k


def _check_draws(rows, seed):
    import random
    from dsl_b200 import geometry as GEO
    np.random.seed(seed)
    random.seed(seed)
    seen = set()
    for h, w, ih, iw, ps, mode, place, flip in rows:
        v, meta = GEO.draw_view((int(h), int(w)), [(1333, 640), (1333, 800)], "value", ps_ratio=0.5, ps_ranges=[0.0, 1.0],
                                ps_modes=["flip", "flop"], flip_ratio=0.5)
        assert (v.img_h, v.img_w, v.ps_mode, v.flip) == (int(ih), int(iw), int(mode), int(flip)), (h, w)
        assert meta["PS"] == bool(ps) and (not ps or meta["PS_place"] == place)
        ext = v.img_w if v.ps_mode == 1 else v.img_h
        assert v.ps_crop == (min(int(round(ext * place)), ext) if ps else 0)
        seen.add((int(mode), int(flip)))
    return seen


def test_draw_view_follows_the_reference_random_stream():
    """geometry.draw_view consumes NumPy's / Python's generators like the reference's Resize('value') -> PatchShuffle ->
    RandomFlip: under the same seeds it lands on the same scale, cut mode / place and flip flag, pass after pass (golden
    view_draws.npz from the reference's own classes; other seeds live where the reference tree is present)."""
    g = np.load(os.path.join(G, "view_draws.npz"))
    seen = _check_draws(g["rows"], int(g["seed"][0]))
    assert {(0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (2, 1)} <= seen
    from oracle import ref_loader
    if ref_loader.available():
        from oracle.gen_golden import view_draw_cases
        for seed in (1, 2, 3):
            _check_draws(view_draw_cases(seed, 25), seed)
