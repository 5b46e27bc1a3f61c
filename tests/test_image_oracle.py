"""The pixel-side view oracle (oracle/image_oracle.py; SURVEY section 8(f) row 3, the next widening step) pinned on CPU:
the uint8 bilinear resize bit-exactly against cv2.resize — the arithmetic the reference reaches through
mmcv.imrescale(backend='cv2') — and the whole Resize -> PatchShuffle -> RandomFlip -> Normalize -> Pad chain bit-exactly
against the reference's own pipeline classes (golden view_image.npz, plus live sweeps where the reference is present)."""
import os

import numpy as np
import pytest

from oracle import image_oracle as IO

G = os.path.join(os.path.dirname(__file__), "golden")


def test_bilinear_u8_resize_is_bit_exact_with_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(0)
    sizes = [(rng.randint(20, 300), rng.randint(20, 300), rng.randint(20, 400), rng.randint(20, 400)) for _ in range(40)]
    sizes += [(480, 640, 800, 1067), (427, 640, 800, 1199), (64, 64, 64, 64), (33, 57, 1, 1), (2, 2, 31, 17),
              (1, 1, 5, 5), (1, 7, 3, 40), (3, 2, 31, 17), (480, 640, 240, 320), (100, 100, 50, 200)]
    for sh, sw, dh, dw in sizes:
        img = rng.randint(0, 256, size=(sh, sw, 3)).astype(np.uint8)
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(IO.imresize_bilinear_u8(img, dw, dh), ref), (sh, sw, dh, dw)


def test_rescale_size_matches_the_coco_shapes_of_the_config():
    # img_scale (1333, 800) / (1333, 640), keep_ratio: the usual COCO sizes
    assert IO.rescale_size(640, 480, (1333, 800)) == (1067, 800)
    assert IO.rescale_size(640, 427, (1333, 800)) == (1199, 800)
    assert IO.rescale_size(500, 375, (1333, 640)) == (853, 640)
    assert IO.rescale_size(333, 500, (1333, 800)) == (800, 1201)


def _check_cases(g):
    n = int(g["meta"][0])
    seen = set()
    for k in range(n):
        sl, ss, mode, place, flip = g["views"][k]
        out, meta = IO.view_image(g[f"c{k}_src"], (int(sl), int(ss)), ps_mode=int(mode), ps_place=float(place),
                                  flip=bool(flip))
        ref = g[f"c{k}_out"]                                      # HWC fp32, padded
        assert out.shape == (3,) + ref.shape[:2], (k, out.shape, ref.shape)
        assert np.array_equal(out, ref.transpose(2, 0, 1)), k     # bit-exact float32
        assert np.array_equal(meta["scale_factor"], g[f"c{k}_scale_factor"])
        assert tuple(g[f"c{k}_img_shape"]) == meta["img_shape"] and meta["pad_shape"][0] % 32 == 0
        seen.add((int(mode), bool(flip)))
    return seen


def test_view_image_matches_reference_golden():
    seen = _check_cases(np.load(os.path.join(G, "view_image.npz")))
    assert {(0, False), (1, True), (2, False)} <= seen


def test_view_image_vs_live_reference_other_seeds():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present (GPU box): covered by view_image.npz")
    from oracle.gen_golden import view_image_cases
    for seed in range(410, 414):
        _check_cases(view_image_cases(seed, 12))


# ---- the kernel's per-pixel arithmetic (dsl_b200/csrc/view_image.cuh) compiled for the host ----------------------------

@pytest.fixture(scope="module")
def host_math(tmp_path_factory):
    import ctypes
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    so = str(tmp_path_factory.mktemp("vi") / "view_image_host.so")
    src = os.path.join(os.path.dirname(__file__), "view_image_host.cpp")
    subprocess.run([cxx, "-O2", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", so], check=True)
    lib = ctypes.CDLL(so)
    lib.view_image_host.restype = None
    lib.view_images_grid_host.restype = None
    return lib


def _host_view_image(lib, src, scale, ps_mode, ps_place, flip, mean=(123.675, 116.28, 103.53),
                     std=(58.395, 57.12, 57.375), to_rgb=True):
    import ctypes
    h, w = src.shape[:2]
    nw, nh = IO.rescale_size(w, h, scale)
    ext = nw if ps_mode == 1 else nh
    crop = min(int(round(ext * ps_place)), ext) if ps_mode else 0
    H, W = -(-nh // 32) * 32, -(-nw // 32) * 32
    view = np.array([h, w, nh, nw, ps_mode, crop, int(flip), 0], dtype=np.int32)
    m, s = np.asarray(mean, np.float32), np.asarray(std, np.float32)
    out = np.full((3, H, W), np.nan, dtype=np.float32)
    src = np.ascontiguousarray(src)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    lib.view_image_host(p(src), p(view), p(m), p(s), int(to_rgb), p(out), H, W)
    return out


def test_kernel_pixel_math_on_host_matches_reference_golden(host_math):
    g = np.load(os.path.join(G, "view_image.npz"))
    for k in range(int(g["meta"][0])):
        sl, ss, mode, place, flip = g["views"][k]
        out = _host_view_image(host_math, g[f"c{k}_src"], (int(sl), int(ss)), int(mode), float(place), bool(flip))
        assert np.array_equal(out, g[f"c{k}_out"].transpose(2, 0, 1)), k


def test_kernel_pixel_math_on_host_matches_oracle_random_views(host_math):
    rng = np.random.RandomState(7)
    for k in range(40):
        h, w = int(rng.randint(8, 200)), int(rng.randint(8, 200))
        src = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        scale = (int(rng.randint(40, 400)), int(rng.randint(20, 300)))
        mode = int(rng.randint(0, 3))
        place = float(rng.choice([0.0, 1.0, rng.uniform()]))
        flip = bool(rng.randint(0, 2))
        to_rgb = bool(k % 3)
        mean, std = rng.uniform(90, 130, 3), rng.uniform(40, 70, 3)
        ref, _ = IO.view_image(src, scale, mode, place, flip, mean=mean, std=std, to_rgb=to_rgb)
        out = _host_view_image(host_math, src, scale, mode, place, flip, mean=mean, std=std, to_rgb=to_rgb)
        assert np.array_equal(out, ref), (k, h, w, scale, mode, place, flip)
    for h, w, scale in [(1, 1, (9, 5)), (1, 7, (40, 3)), (2, 3, (31, 17)), (3, 2, (17, 31)), (64, 48, (64, 48)),
                        (480, 640, (1333, 800)), (640, 427, (1333, 640))]:     # degenerate, identity and COCO-sized
        src = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        for mode, flip in ((0, False), (1, True), (2, True)):
            ref, _ = IO.view_image(src, scale, mode, 0.3, flip)
            assert np.array_equal(_host_view_image(host_math, src, scale, mode, 0.3, flip), ref), (h, w, scale, mode)


def test_gpu_test_bodies_dry_run_with_host_math(host_math, monkeypatch):
    """The bodies of the view_images GPU tests (view construction from the draws, ragged batching, padding, the roll /
    mirror properties, comparison with the oracle and the golden) executed here with the host-compiled kernel arithmetic
    standing in for the launch — the host harness walks the SAME grid through the SAME per-thread function (vi_thread)
    the __global__ kernel calls, so block / thread indexing, tile edges and the zero padding are covered too."""
    import ctypes
    import torch
    from dsl_b200 import geometry as GEO
    from tests import test_view_image_gpu as T

    def fake_view_images(srcs, views, mean, std, to_rgb=True, H=None, W=None, size_divisor=32, out=None):
        up = lambda n: (n + size_divisor - 1) // size_divisor * size_divisor  # noqa: E731
        H = up(max(v.img_h for v in views)) if H is None else H
        W = up(max(v.img_w for v in views)) if W is None else W
        assert all(v.img_h <= H and v.img_w <= W for v in views)
        m, sd = np.asarray(mean, np.float32), np.asarray(std, np.float32)
        p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
        arrs = [np.ascontiguousarray(s.numpy()) for s in srcs]
        for a, v in zip(arrs, views):
            assert a.dtype == np.uint8 and a.shape == (v.src_h, v.src_w, 3)
        ptrs = (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        vv = np.array([[v.src_h, v.src_w, v.img_h, v.img_w, v.ps_mode, v.ps_crop, v.flip, 0] for v in views], np.int32)
        res = np.full((len(srcs), 3, H, W), np.nan, np.float32)          # NaN: every element must be written
        # the launch exactly as dslb_view_images issues it: same grid, same per-thread function as the __global__ kernel
        host_math.view_images_grid_host(ptrs, p(vv), len(arrs), p(m), p(sd), int(to_rgb), p(res), H, W)
        assert not np.isnan(res).any()
        return torch.from_numpy(res)

    monkeypatch.setattr(GEO, "view_images", fake_view_images)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    T.test_view_images_match_reference_golden()
    T.test_view_images_random_views_vs_oracle()
    T.test_view_images_coco_sized_batch_and_properties()
    # output sizes that are not multiples of the 32 x 32 tile: the edge guards of the grid walk
    rng = np.random.RandomState(3)
    srcs = [rng.randint(0, 256, size=(30, 45, 3)).astype(np.uint8), rng.randint(0, 256, size=(21, 19, 3)).astype(np.uint8)]
    draws = [((60, 40), 1, 0.3, True), ((33, 33), 2, 0.6, False)]
    views = [GEO.image_view(s.shape[:2], sc, m, p, f)[0] for s, (sc, m, p, f) in zip(srcs, draws)]
    out = fake_view_images([torch.from_numpy(s) for s in srcs], views, T.MEAN, T.STD, H=50, W=70).numpy()
    ref = np.zeros((2, 3, 50, 70), np.float32)
    for b, (s_, (sc, m, p, f)) in enumerate(zip(srcs, draws)):
        o, meta = IO.view_image(s_, sc, m, p, f)
        nh, nw = meta["img_shape"][:2]
        ref[b, :, :nh, :nw] = o[:, :nh, :nw]
    assert np.array_equal(out, ref)
