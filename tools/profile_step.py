#!/usr/bin/env python
"""ncu driver: one EAGER teacher+student step of the bench workload between cudaProfilerStart/Stop.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [--batch 4] [--hw 800x1344]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from dsl_b200.trainer import DSLEngine
from bench import confident_heads, make_gt

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--hw", default="800x1344")
ap.add_argument("--warm", type=int, default=1)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--backbone", default="resnet", choices=["resnet", "rla"])
ap.add_argument("--depth", type=int, default=50)
a = ap.parse_args()
H, W = (int(v) for v in a.hw.split("x"))
B = a.batch
eng = DSLEngine(B, H, W, depth=a.depth, seed=0, use_graphs=False, backbone=a.backbone)
rng = np.random.RandomState(100)
img_s = torch.from_numpy((rng.rand(B, 3, H, W) * 255 - 115).astype(np.float32))
img_t = torch.from_numpy((rng.rand(B, 3, H, W) * 255 - 115).astype(np.float32))
gts, labels, ignores = make_gt(200, B, H, W, max_gt=20, max_ignore=5)
confident_heads(eng)
eng.set_inputs(img_s, gts, labels, ignores, teacher_img=img_t)
for _ in range(a.warm):
    eng.step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    eng.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("losses", {k: float(v) for k, v in eng.student.losses().items()})
