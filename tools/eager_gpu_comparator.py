#!/usr/bin/env python
"""The honest GPU comparator of SURVEY section 8(d): the reference's arithmetic for one teacher+student step (the oracle
restatement, oracle/cpu_step.py — plain torch modules' functional calls + autograd, NCHW) executed by STOCK torch eager +
cuDNN on the same B200, at the benchmark shape. Baseline only (test / bench infrastructure): nothing of the product runs
here. Prints one JSON line per precision: fp32 (cuDNN TF32 convs, torch's default) and bf16 autocast.
  python tools/eager_gpu_comparator.py [--batch 4] [--hw 800x1344] [--steps 5] [--backbone resnet|rla]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--hw", default="800x1344")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--depth", type=int, default=50)
    ap.add_argument("--backbone", default="resnet", choices=["resnet", "rla"])
    ap.add_argument("--device", default="cuda")
    args = ap.parse_args()
    from oracle.cpu_step import CpuStep
    H, W = (int(v) for v in args.hw.split("x"))
    cuda = args.device.startswith("cuda")
    for name, amp in (("fp32 (cuDNN, TF32 convs allowed)", None), ("bf16 autocast", torch.bfloat16)):
        cs = CpuStep(args.batch, H, W, depth=args.depth, backbone=args.backbone, device=args.device)
        cs.autocast = amp
        for _ in range(args.warmup):
            losses = cs.step()
        if cuda:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        else:
            import time
            t0 = time.perf_counter()
        for _ in range(args.steps):
            losses = cs.step()
        if cuda:
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
        else:
            ms = (time.perf_counter() - t0) * 1e3 / args.steps
        print(json.dumps(dict(impl="torch eager + cuDNN, reference arithmetic (oracle restatement)", precision=name,
                              device=args.device, batch=args.batch, hw=[H, W], backbone=args.backbone, steps=args.steps,
                              ms_per_step=round(ms, 2), img_per_s=round(args.batch / (ms * 1e-3), 2),
                              losses={k: round(v, 4) for k, v in losses.items()},
                              note="includes the host syncs of the reference's loss (nonzero, float()) like the "
                                   "reference's own training loop")), flush=True)
        del cs


if __name__ == "__main__":
    main()
