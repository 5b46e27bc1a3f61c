"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy, integer / fp32 arithmetic spelled out) of the PIXEL side of the
reference's view pipelines (configs/fcos_semi/*.py:70-92,108-121): Resize(keep_ratio) -> PatchShuffle -> RandomFlip ->
Normalize(to_rgb) -> Pad(size_divisor=32) -> CHW, i.e. what turns one uint8 HWC BGR image into the student's / teacher's
fp32 network input. SURVEY section 8(f) row 3 lists the device-side version of this as the next widening step; this file
is its oracle (the box side is oracle/fcos_oracle.py::view_boxes, already a kernel).

Pinned by tests/test_image_oracle.py: `imresize_bilinear_u8` bit-exactly against cv2.resize(INTER_LINEAR) — the arithmetic
the reference reaches through mmcv.imrescale(backend='cv2') (mmdet/datasets/pipelines/transforms.py:218-247) — and the whole
chain bit-exactly against the reference's own Resize / PatchShuffle / RandomFlip / Normalize / Pad classes executed in
place (golden tests/golden/view_image.npz + live sweeps where the reference tree is present).

The un-vendored dependency here is mmcv (pinned >=1.3.8,<=1.4.0, mmdet/__init__.py:19-27): imrescale / imflip /
imnormalize / impad_to_multiple are thin cv2 / numpy wrappers whose published definitions (mmcv/image/geometric.py,
photometric.py of 1.3.x) are restated where used; the only non-trivial arithmetic, the bilinear resize, is OpenCV's.
"""
import numpy as np

COEF_BITS = 11                      # OpenCV INTER_RESIZE_COEF_BITS
COEF_SCALE = np.float32(1 << COEF_BITS)


def rescale_size(w, h, scale):
    """mmcv.rescale_size for a (long edge, short edge) tuple: the largest size that fits both edges, rounded half up
    (mmcv/image/geometric.py: `int(w * float(scale) + 0.5)`)."""
    max_long, max_short = max(scale), min(scale)
    f = min(max_long / max(h, w), max_short / min(h, w))
    return int(w * float(f) + 0.5), int(h * float(f) + 0.5)


def _linear_coeffs(dn, sn, vertical):
    """Source indices and 11-bit fixed-point weights of OpenCV's INTER_LINEAR (modules/imgproc/src/resize.cpp, cv::resize
    -> resizeGeneric_): fx = (float)((dx + 0.5) * scale - 0.5) with scale = 1 / (dsize / ssize) in double; sx = floor(fx);
    weights cvRound((1 - fx) * 2048), cvRound(fx * 2048). Horizontally the weight is reset at the borders (sx < 0 or
    sx >= w - 1); vertically only the ROWS are clamped and the weights are kept."""
    scale = 1.0 / (dn / sn)
    f = ((np.arange(dn) + 0.5) * scale - 0.5).astype(np.float32)
    i0 = np.floor(f).astype(np.int64)
    fr = (f - i0.astype(np.float32)).astype(np.float32)
    if not vertical:
        lo, hi = i0 < 0, i0 >= sn - 1
        fr = np.where(lo | hi, np.float32(0), fr).astype(np.float32)
        i0 = np.where(lo, 0, np.where(hi, sn - 1, i0))
    a1 = np.rint(fr * COEF_SCALE).astype(np.int32)
    a0 = np.rint((np.float32(1) - fr) * COEF_SCALE).astype(np.int32)
    return np.clip(i0, 0, sn - 1), np.clip(i0 + 1, 0, sn - 1), a0, a1


def imresize_bilinear_u8(img, dw, dh):
    """cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR) for uint8 HWC images, bit-exact: horizontal pass in
    int32 (pixel * 11-bit weight), vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2."""
    assert img.dtype == np.uint8 and img.ndim == 3
    sh, sw, _ = img.shape
    s = img.astype(np.int32)
    x0, x1, ax0, ax1 = _linear_coeffs(dw, sw, vertical=False)
    y0, y1, ay0, ay1 = _linear_coeffs(dh, sh, vertical=True)
    rows = s[:, x0, :] * ax0[None, :, None] + s[:, x1, :] * ax1[None, :, None]          # [sh][dw][c], scaled by 2^11
    s0, s1 = rows[y0], rows[y1]
    b0, b1 = ay0[:, None, None], ay1[:, None, None]
    out = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def patch_shuffle_pixels(img, ps_mode, ps_crop):
    """PatchShuffle.__call__, pixel part (transforms.py:2180-2199): 'flip' (1) moves the left `crop` columns to the right
    end, 'flop' (2) the top `crop` rows to the bottom; crop == 0 or == extent is the no-op the reference returns early on."""
    h, w = img.shape[:2]
    if ps_mode == 1 and 0 < ps_crop < w:
        return np.concatenate([img[:, ps_crop:], img[:, :ps_crop]], axis=1)
    if ps_mode == 2 and 0 < ps_crop < h:
        return np.concatenate([img[ps_crop:], img[:ps_crop]], axis=0)
    return img


def normalize_bgr_u8(img, mean, std, to_rgb=True):
    """mmcv.imnormalize on a uint8 BGR image: float32 copy, BGR -> RGB, cv2.subtract(img, float64 mean), cv2.multiply(img,
    float64 1 / std), both in place on the float32 image. What OpenCV does with the float64 scalar operands (measured
    against cv2 in tests/test_image_oracle.py): the subtraction runs in float32 with the mean rounded to float32, the
    multiplication takes the float32 value times the DOUBLE 1 / std and rounds the product to float32 once. The
    reference's Normalize stores mean and std as float32 arrays first (transforms.py:665-666), so 1 / std is the double
    reciprocal of the float32-rounded std."""
    x = img.astype(np.float32)
    if to_rgb:
        x = x[..., ::-1]
    mean32 = np.asarray(mean, dtype=np.float64).astype(np.float32)
    inv64 = 1.0 / np.asarray(std, dtype=np.float32).astype(np.float64)
    return ((x - mean32).astype(np.float64) * inv64).astype(np.float32)


def view_image(src, scale, ps_mode=0, ps_place=0.0, flip=False, mean=(123.675, 116.28, 103.53),
               std=(58.395, 57.12, 57.375), to_rgb=True, divisor=32):
    """One uint8 HWC BGR image -> (fp32 CHW network input zero-padded to a multiple of `divisor`, meta dict with
    img_shape, pad_shape, scale_factor, ps_crop): Resize(img_scale=scale, keep_ratio=True) -> PatchShuffle(mode, place)
    -> RandomFlip(horizontal) -> Normalize -> Pad -> DefaultFormatBundle's HWC -> CHW."""
    h, w = src.shape[:2]
    nw, nh = rescale_size(w, h, scale)
    img = imresize_bilinear_u8(src, nw, nh)
    scale_factor = np.array([nw / w, nh / h, nw / w, nh / h], dtype=np.float32)            # transforms.py:229-242
    ext = nw if ps_mode == 1 else nh
    crop = min(int(round(ext * ps_place)), ext) if ps_mode else 0                           # transforms.py:2181,2190
    img = patch_shuffle_pixels(img, ps_mode, crop)
    if flip:
        img = img[:, ::-1]                                                                  # mmcv.imflip, horizontal
    x = normalize_bgr_u8(img, mean, std, to_rgb)
    ph, pw = -(-nh // divisor) * divisor, -(-nw // divisor) * divisor                       # mmcv.impad_to_multiple
    out = np.zeros((3, ph, pw), dtype=np.float32)
    out[:, :nh, :nw] = x.transpose(2, 0, 1)
    return out, dict(img_shape=(nh, nw, 3), pad_shape=(ph, pw, 3), scale_factor=scale_factor, ps_crop=crop)
