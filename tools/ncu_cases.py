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


def build_tower(B=4, H=800, W=1344):
    """One FCOSHead tower layer exactly as the engine launches it (engine._build_head): 2 branches x 5 FPN levels = 10
    segments of 3x3 256 -> 256 with conv bias and GroupNorm statistics in the epilogue, one persistent launch."""
    from dsl_b200 import _lib as L
    sizes, h, w = [], H // 8, W // 8
    for _ in range(3):
        sizes.append((h, w))
        h, w = (h + 1) // 2, (w + 1) // 2
    h5, w5 = sizes[2]
    h6, w6 = (h5 + 1) // 2, (w5 + 1) // 2
    sizes = sizes[:3] + [(h6, w6), ((h6 + 1) // 2, (w6 + 1) // 2)]
    plans = []
    for _ in range(2):
        segs = []
        for br in range(2):
            wt = (torch.randn(9, 256, 256, device=dev) * 0.02).to(BF)
            bias = torch.randn(256, device=dev) * 0.1
            for (hh, ww) in sizes:
                x = torch.randn(B, hh, ww, 256, device=dev).to(BF)
                y = torch.zeros(B, hh, ww, 256, dtype=BF, device=dev)
                st = torch.zeros(B, 32, L.GN_STAT_STRIDE, dtype=torch.float64, device=dev)
                segs.append(dict(x=x, w=wt, y=y, N=B, H=hh, W=ww, Cin=256, Cout=256, cout_pad=256, R=3, S=3, stride=1,
                                 pad=1, ldc=256, shift=bias, gn_stats=st, gn_cpg=8))
        plans.append(ConvPlan(segs, "head.tower"))
    return plans


def build_stem(B=4, H=800, W=1344):
    from dsl_b200 import _lib as L

    class Stem:
        def __init__(self):
            self.img = torch.randn(B, 3, H, W, device=dev) * 50
            self.w = torch.randn(64, 3, 7, 7, device=dev) * 0.05
            self.bn = [torch.ones(64, device=dev), torch.zeros(64, device=dev), torch.zeros(64, device=dev),
                       torch.ones(64, device=dev)]
            self.ws = torch.zeros(B, H, W, 4, dtype=BF, device=dev)
            self.out = torch.zeros(B, H // 2, W // 2, 64, dtype=BF, device=dev)

        def run(self):
            L.check(L.lib.dslb_stem_conv(L.ptr(self.img), L.ptr(self.w), *[L.ptr(t) for t in self.bn], 1e-5, L.ptr(self.ws),
                                         L.ptr(self.out), B, H, W, L.cur_stream()), "stem")

    return [Stem(), Stem()]


if __name__ == "__main__":
    for n in sys.argv[1:]:
        plans = build_tower() if n == "tower_full" else build_stem() if n == "stem_full" else build(n, *CASES[n])
        plans[0].run()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        plans[1].run()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print("captured", n, flush=True)
