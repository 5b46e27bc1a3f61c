"""Device-side view geometry: carries packed box lists through the box part of the reference's
Resize -> PatchShuffle -> RandomFlip pipeline steps (mmdet/datasets/pipelines/transforms.py:249-257, 2168-2248, 397-429),
renders the pixel side of the same steps + Normalize + Pad from uint8 source images (`view_images`, bit-exact with the
cv2 / mmcv CPU pipeline) and assembles zero-padded batches, on libdslb.so. With it the EMA teacher's boxes
(original-image coordinates) reach the student's strong view without the host data pipeline. CUDA only."""
import ctypes as C

import torch

from . import _lib as L


class View(C.Structure):
    """dslb_view_t"""
    _fields_ = [("sx", C.c_float), ("sy", C.c_float), ("img_w", C.c_int32), ("img_h", C.c_int32), ("clip", C.c_int32),
                ("ps_mode", C.c_int32), ("ps_crop", C.c_int32), ("flip", C.c_int32)]


PS_MODES = {None: 0, False: 0, "flip": 1, "flop": 2}


def view_from_meta(meta, clip=True):
    """img_metas entry of the reference's Collect (`scale_factor`, `img_shape`, `flip`, `PS`, `PS_mode`, `PS_place`,
    configs/fcos_semi/*.py:80) -> View. ps_crop = min(int(round(extent * PS_place)), extent) as PatchShuffle computes it."""
    h, w = int(meta["img_shape"][0]), int(meta["img_shape"][1])
    sf = meta.get("scale_factor", (1.0, 1.0, 1.0, 1.0))
    mode = PS_MODES[meta.get("PS_mode")] if meta.get("PS") else 0
    crop = 0
    if mode:
        ext = w if mode == 1 else h
        crop = min(int(round(ext * float(meta["PS_place"]))), ext)
    if meta.get("flip") and meta.get("flip_direction", "horizontal") != "horizontal":
        raise NotImplementedError("dsl_b200.geometry: only horizontal flips (the DSL configs' RandomFlip default)")
    return View(float(sf[0]), float(sf[1]), w, h, int(bool(clip)), mode, crop, int(bool(meta.get("flip"))))


class ViewGeometry:
    """Buffers + launch for B images and up to `max_boxes` input boxes."""

    def __init__(self, B, max_boxes=1024, device="cuda"):
        self.B, self.max_in, self.max_out = B, max_boxes, 2 * max_boxes
        self.dev = torch.device(device)
        self.views = torch.zeros(B * C.sizeof(View), dtype=torch.uint8, device=self.dev)
        self._h_views = torch.zeros(B * C.sizeof(View), dtype=torch.uint8).pin_memory()
        self.ws_bytes = L.lib.dslb_view_boxes_workspace_bytes(max_boxes)
        self.ws = torch.zeros(self.ws_bytes, dtype=torch.uint8, device=self.dev)
        self.out_boxes = torch.zeros(self.max_out, 4, dtype=torch.float32, device=self.dev)
        self.out_labels = torch.zeros(self.max_out, dtype=torch.int64, device=self.dev)
        self.out_off = torch.zeros(B + 1, dtype=torch.int32, device=self.dev)

    def set_views(self, views):
        assert len(views) == self.B
        arr = (View * self.B)(*views)
        self._h_views.copy_(torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8))
        self.views.copy_(self._h_views, non_blocking=True)

    def run(self, boxes, labels, off, out_boxes=None, out_labels=None, out_off=None):
        """boxes (n,4) fp32 / labels (n,) int64 or None / off (B+1,) int32, all on the device, packed over images."""
        ob = self.out_boxes if out_boxes is None else out_boxes
        ol = (self.out_labels if out_labels is None else out_labels) if labels is not None else None
        oo = self.out_off if out_off is None else out_off
        assert boxes.shape[0] <= self.max_in
        L.check(L.lib.dslb_view_boxes(L.ptr(boxes), L.ptr(labels) if labels is not None else None, L.ptr(off),
                                      L.ptr(self.views), self.B, self.max_in, int(ob.shape[0]), L.ptr(self.ws),
                                      self.ws_bytes, L.ptr(ob), L.ptr(ol) if ol is not None else None, L.ptr(oo),
                                      L.cur_stream()), "view_boxes")
        return ob, ol, oo


class ImageView(C.Structure):
    """dslb_image_view_t"""
    _fields_ = [("src_h", C.c_int32), ("src_w", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32),
                ("ps_mode", C.c_int32), ("ps_crop", C.c_int32), ("flip", C.c_int32), ("reserved", C.c_int32)]


def rescale_size(w, h, scale):
    """mmcv.rescale_size for an img_scale tuple (Resize keep_ratio=True, transforms.py:218-230): the largest size that
    keeps the long edge <= max(scale) and the short edge <= min(scale), each edge rounded half up."""
    f = min(max(scale) / max(h, w), min(scale) / min(h, w))
    return int(w * float(f) + 0.5), int(h * float(f) + 0.5)


def image_view(src_hw, scale, ps_mode=None, ps_place=0.0, flip=False):
    """The draws of one pipeline pass (Resize's img_scale tuple, PatchShuffle's mode / place, RandomFlip's flag) ->
    (ImageView, img_meta dict with the keys the reference's Collect hands on: img_shape, scale_factor, flip, PS, ...)."""
    import numpy as np
    h, w = int(src_hw[0]), int(src_hw[1])
    nw, nh = rescale_size(w, h, scale)
    import numbers
    if isinstance(ps_mode, numbers.Integral) and not isinstance(ps_mode, bool):
        mode = int(ps_mode)
        if mode not in (0, 1, 2):
            raise ValueError(f"dsl_b200.geometry.image_view: ps_mode {mode} (0 off, 1 'flip', 2 'flop')")
    else:
        mode = PS_MODES[ps_mode]
    crop = 0
    if mode:
        ext = nw if mode == 1 else nh
        crop = min(int(round(ext * float(ps_place))), ext)
    ws, hs = nw / w, nh / h
    meta = dict(ori_shape=(h, w, 3), img_shape=(nh, nw, 3), scale_factor=np.array([ws, hs, ws, hs], dtype=np.float32),
                flip=bool(flip), flip_direction="horizontal" if flip else None, PS=bool(mode),
                PS_mode={0: None, 1: "flip", 2: "flop"}[mode], PS_place=float(ps_place))
    return ImageView(h, w, nh, nw, mode, crop, int(bool(flip)), 0), meta


def draw_view(src_hw, img_scales, multiscale_mode="value", ps_ratio=None, ps_ranges=(0.2, 0.8), ps_modes=("flip", "flop"),
              flip_ratio=None):
    """The random draws of one pass through Resize -> PatchShuffle -> RandomFlip, consuming NumPy's and Python's global
    generators in the reference's order and amounts (transforms.py:127-129 random_select / :145-152 random_sample,
    :2169-2180 PatchShuffle, :443-462 RandomFlip), so that under the same seeds the device-side views are the views the
    reference's pipeline would have produced. ps_ratio / flip_ratio None: that step is not in the pipeline (no draw).
    Returns image_view(...)'s (ImageView, img_meta)."""
    import random
    import numpy as np
    img_scales = list(img_scales)
    if len(img_scales) == 1:
        scale = img_scales[0]
    elif multiscale_mode == "value":
        scale = img_scales[np.random.randint(len(img_scales))]
    elif multiscale_mode == "range":
        assert len(img_scales) == 2
        longs, shorts = [max(s) for s in img_scales], [min(s) for s in img_scales]
        long_edge = np.random.randint(min(longs), max(longs) + 1)
        short_edge = np.random.randint(min(shorts), max(shorts) + 1)
        scale = (long_edge, short_edge)
    else:
        raise NotImplementedError(f"dsl_b200.geometry.draw_view: multiscale_mode {multiscale_mode!r}")
    mode, place = None, 0.0
    if ps_ratio is not None and not (np.random.rand(1) > ps_ratio):
        seed = np.random.rand(1)[0]
        place = seed * abs(ps_ranges[1] - ps_ranges[0]) + ps_ranges[0]
        mode = random.choice(list(ps_modes))
    flip = False
    if flip_ratio is not None:
        flip = np.random.choice(["horizontal", None], p=[flip_ratio, 1 - flip_ratio]) is not None
    return image_view(src_hw, scale, ps_mode=mode, ps_place=float(place), flip=flip)


def view_images(srcs, views, mean, std, to_rgb=True, H=None, W=None, size_divisor=32, out=None):
    """Resize -> PatchShuffle -> RandomFlip -> Normalize -> Pad -> collate of the reference's train pipelines
    (configs/fcos_semi/*.py:70-92) for a batch: `srcs` = uint8 HWC 3-channel CUDA tensors (cv2.imread order), `views` =
    ImageView per image -> (B, 3, H, W) fp32 CUDA batch, zero-padded (H, W = per-batch maxima rounded up to the divisor
    unless given), one kernel launch."""
    B = len(srcs)
    assert B == len(views) and B >= 1
    srcs = [s.contiguous() for s in srcs]
    for s, v in zip(srcs, views):
        if s.dtype != torch.uint8 or s.dim() != 3 or s.shape[2] != 3 or not s.is_cuda:
            raise ValueError("dsl_b200.geometry.view_images: sources must be uint8 HWC 3-channel CUDA tensors")
        if (int(s.shape[0]), int(s.shape[1])) != (v.src_h, v.src_w):
            raise ValueError("dsl_b200.geometry.view_images: view.src_h / src_w do not match the source image")
        if min(v.src_h, v.src_w, v.img_h, v.img_w) < 1:
            raise ValueError("dsl_b200.geometry.view_images: empty source image or view")
    up = lambda n: (n + size_divisor - 1) // size_divisor * size_divisor  # noqa: E731
    H = up(max(v.img_h for v in views)) if H is None else H
    W = up(max(v.img_w for v in views)) if W is None else W
    if any(v.img_h > H or v.img_w > W for v in views):
        raise ValueError("dsl_b200.geometry.view_images: a view is larger than the output batch")
    dev = srcs[0].device
    ptrs = torch.tensor([s.data_ptr() for s in srcs], dtype=torch.int64, device=dev)
    varr = (ImageView * B)(*views)
    vdev = torch.frombuffer(bytearray(bytes(varr)), dtype=torch.uint8).to(dev)
    if out is None:
        out = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev)
    assert out.shape == (B, 3, H, W) and out.dtype == torch.float32 and out.is_contiguous()
    m = (C.c_float * 3)(*[float(x) for x in mean])
    sd = (C.c_float * 3)(*[float(x) for x in std])
    L.check(L.lib.dslb_view_images(L.ptr(ptrs), L.ptr(vdev), B, C.cast(m, C.c_void_p), C.cast(sd, C.c_void_p),
                                   int(bool(to_rgb)), L.ptr(out), H, W, L.cur_stream()), "view_images")
    return out


def pad_batch(imgs, H=None, W=None, size_divisor=32):
    """Pad(size_divisor) + collate of the reference's loader: list of (C, h_i, w_i) fp32 CUDA tensors -> (B, C, H, W)
    zero-padded batch (H, W = per-batch maxima rounded up to the divisor unless given)."""
    B, Cc = len(imgs), int(imgs[0].shape[0])
    imgs = [i.contiguous().float() for i in imgs]
    hs, ws = [int(i.shape[1]) for i in imgs], [int(i.shape[2]) for i in imgs]
    up = lambda v: (v + size_divisor - 1) // size_divisor * size_divisor  # noqa: E731
    H = up(max(hs)) if H is None else H
    W = up(max(ws)) if W is None else W
    dev = imgs[0].device
    ptrs = torch.tensor([i.data_ptr() for i in imgs], dtype=torch.int64, device=dev)
    hw = torch.tensor([[h, w] for h, w in zip(hs, ws)], dtype=torch.int32, device=dev)
    out = torch.empty(B, Cc, H, W, dtype=torch.float32, device=dev)
    L.check(L.lib.dslb_pad_batch(L.ptr(ptrs), L.ptr(hw), L.ptr(out), B, Cc, H, W, L.cur_stream()), "pad_batch")
    return out
