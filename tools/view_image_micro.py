#!/usr/bin/env python
"""Micro-benchmark of dslb_view_images (pixel side of the view pipelines) at the BASELINE batch shape: B COCO-sized uint8
sources -> (B, 3, 800, 1344) fp32. CUDA events on the launching stream, an L2 flush between launches, achieved GB/s of
the ALGORITHMIC bytes (12 B written per padded output pixel + the source bytes once) against MEASURED_PEAKS.json's HBM
copy bandwidth. Also prints the pinned-host -> device cost of the two input formats (uint8 sources vs the fp32 batch).
  python tools/view_image_micro.py [--batch 4] [--iters 50]
  ncu --set full --clock-control none -k regex:view_images -c 1 -o gpurun_out/view_images python tools/view_image_micro.py --iters 1
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--iters", type=int, default=50)
    args = ap.parse_args()
    from dsl_b200 import geometry as GEO
    H, W = 800, 1344
    rng = np.random.RandomState(0)
    shapes = [(480, 640), (427, 640), (480, 600), (375, 500)]      # landscape: every view fits 800 x 1344
    srcs = [rng.randint(0, 256, size=shapes[b % 4] + (3,)).astype(np.uint8) for b in range(args.batch)]
    draws = [((1333, 800), b % 3, 0.37, bool(b % 2)) for b in range(args.batch)]
    views = [GEO.image_view(s.shape[:2], sc, m, p, f)[0] for s, (sc, m, p, f) in zip(srcs, draws)]
    hsrc = [torch.from_numpy(s).pin_memory() for s in srcs]
    dsrc = [h.cuda() for h in hsrc]
    out = torch.empty(args.batch, 3, H, W, dtype=torch.float32, device="cuda")
    mean, std = (123.675, 116.28, 103.53), (58.395, 57.12, 57.375)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # the launch alone (pointer / view tables uploaded once), so that the events bracket the kernel and not host work
    import ctypes as C
    from dsl_b200 import _lib as L
    check = GEO.view_images(dsrc, views, mean, std, H=H, W=W)                   # the public path once, for comparison
    ptrs = torch.tensor([s.data_ptr() for s in dsrc], dtype=torch.int64, device="cuda")
    varr = (GEO.ImageView * len(views))(*views)
    vdev = torch.frombuffer(bytearray(bytes(varr)), dtype=torch.uint8).cuda()
    m, sd = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)

    def run():
        L.check(L.lib.dslb_view_images(L.ptr(ptrs), L.ptr(vdev), len(views), C.cast(m, C.c_void_p), C.cast(sd, C.c_void_p),
                                       1, L.ptr(out), H, W, L.cur_stream()), "view_images")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    assert torch.equal(out, check), "direct launch and geometry.view_images disagree"
    ms = []
    for _ in range(args.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms)) * 1e-3
    bytes_alg = out.numel() * 4 + sum(s.size for s in srcs)
    peak = None
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs")
    # host -> device cost of the two input formats
    hf = torch.empty(args.batch, 3, H, W, dtype=torch.float32).pin_memory()

    def h2d(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 10
    t_u8 = h2d(lambda: [d.copy_(h, non_blocking=True) for d, h in zip(dsrc, hsrc)])
    t_f32 = h2d(lambda: out.copy_(hf, non_blocking=True))
    # the reference's CPU path for the same batch: the cv2 calls mmcv makes (resize / hconcat / flip / normalize / pad)
    cpu = None
    try:
        import time
        import cv2

        def cpu_batch():
            res = np.zeros((args.batch, 3, H, W), np.float32)
            mean64 = np.float64(np.array(mean, np.float32).reshape(1, -1))
            inv64 = 1 / np.float64(np.array(std, np.float32).reshape(1, -1))
            for b, (s, v) in enumerate(zip(srcs, views)):
                img = cv2.resize(s, (v.img_w, v.img_h), interpolation=cv2.INTER_LINEAR)
                if v.ps_mode == 1 and 0 < v.ps_crop < v.img_w:
                    img = cv2.hconcat([img[:, v.ps_crop:], img[:, :v.ps_crop]])
                elif v.ps_mode == 2 and 0 < v.ps_crop < v.img_h:
                    img = cv2.vconcat([img[v.ps_crop:], img[:v.ps_crop]])
                if v.flip:
                    img = np.flip(img, axis=1)
                img = img.copy().astype(np.float32)
                cv2.cvtColor(img, cv2.COLOR_BGR2RGB, img)
                cv2.subtract(img, mean64, img)
                cv2.multiply(img, inv64, img)
                res[b, :, :v.img_h, :v.img_w] = img.transpose(2, 0, 1)
            return res
        ref = cpu_batch()
        t0 = time.perf_counter()
        for _ in range(3):
            cpu_batch()
        cpu = dict(ms_per_batch=round((time.perf_counter() - t0) / 3 * 1e3, 2), threads=cv2.getNumThreads(),
                   bit_exact_with_kernel=bool(np.array_equal(ref, check.cpu().numpy())))
    except Exception as e:  # cv2 missing: the GPU figures stand on their own
        cpu = dict(error=repr(e))
    print(json.dumps(dict(kernel="view_images_kernel", batch=args.batch, out_shape=[args.batch, 3, H, W],
                          us_median=round(t * 1e6, 2), us_min=round(min(ms) * 1e3, 2), algorithmic_bytes=bytes_alg,
                          achieved_gbps=round(bytes_alg / t / 1e9, 1), peak_gbps=peak,
                          frac=None if not peak else round(bytes_alg / t / 1e9 / peak, 4),
                          note="CUDA events around the C-ABI launch alone; L2 flushed before every launch",
                          h2d_ms=dict(uint8_sources=round(t_u8, 3), fp32_batch=round(t_f32, 3)),
                          cpu_reference=cpu)))


if __name__ == "__main__":
    main()
