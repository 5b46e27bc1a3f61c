#!/usr/bin/env python
"""ncu driver for single conv plans: each named case of tools/conv_micro.py is launched once (after a warm launch on a
different buffer set) between cudaProfilerStart/Stop.

    ncu --profile-from-start off --set full --import-source on --clock-control none -o gpurun_out/cases \
        python tools/ncu_cases.py l1_conv3_64_256_res tower_3x3_256
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dsl_b200.engine import ConvPlan
from tools.conv_micro import BF, CASES, dev


def build(name, N, H, W, Ci, Co, R, stride, pad, res, mask, affine, relu):
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    w = (torch.randn(R * R, Co, Ci, device=dev) * 0.05).to(BF)
    scale = torch.rand(Co, device=dev) + 0.5
    shift = torch.randn(Co, device=dev)
    plans = []
    for _ in range(2):
        x = torch.randn(N, H, W, Ci, device=dev).to(BF)
        y = torch.zeros(N, Ho, Wo, Co, dtype=BF, device=dev)
        seg = dict(x=x, w=w, y=y, N=N, H=H, W=W, Cin=Ci, Cout=Co, cout_pad=Co, R=R, S=R, stride=stride, pad=pad, ldc=Co,
                   relu_nch=Co if relu else 0)
        if affine:
            seg.update(shift=shift)  # BN scale is folded into the packed weights, as in the engine
        if res:
            seg["residual"] = torch.randn(N, Ho, Wo, Co, device=dev).to(BF)
        if mask:
            seg["relu_mask"] = torch.randn(N, Ho, Wo, Co, device=dev).to(BF)
        plans.append(ConvPlan([seg], name))
    return plans


if __name__ == "__main__":
    for n in sys.argv[1:]:
        plans = build(n, *CASES[n])
        plans[0].run()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        plans[1].run()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print("captured", n, flush=True)
