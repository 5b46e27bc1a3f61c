#!/usr/bin/env python
"""Micro-benchmark of single conv plans at the benchmark's layer shapes (B=4, 800x1344): CUDA-event time per launch
over a loop that cycles through enough buffer sets to defeat the L2 (each launch sees cold operands).
Prints us, TFLOP/s and effective HBM GB/s (algorithmic bytes). DSLB_LIB selects the library build (A/B tests)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dsl_b200 import _lib as L
from dsl_b200.engine import ConvPlan

dev = "cuda"
BF = torch.bfloat16
CASES = {
    # name: (N, H, W, Cin, Cout, R, stride, pad, residual, mask, affine, relu)
    "stem_1x1_192_64": (4, 400, 672, 192, 64, 1, 1, 0, False, False, True, True),
    "l1_conv1_256_64": (4, 200, 336, 256, 64, 1, 1, 0, False, False, True, True),
    "l1_conv2_3x3_64": (4, 200, 336, 64, 64, 3, 1, 1, False, False, True, True),
    "l1_conv3_64_256_res": (4, 200, 336, 64, 256, 1, 1, 0, True, False, True, True),
    "l1_conv3_64_256_nores": (4, 200, 336, 64, 256, 1, 1, 0, False, False, True, True),
    "l2_conv3_128_512_res": (4, 100, 168, 128, 512, 1, 1, 0, True, False, True, True),
    "l3_conv3_256_1024_res": (4, 50, 84, 256, 1024, 1, 1, 0, True, False, True, True),
    "l2_conv1dgrad_128_512_res_mask": (4, 100, 168, 128, 512, 1, 1, 0, True, True, False, False),
    "tower_3x3_256": (4, 100, 168, 256, 256, 3, 1, 1, False, False, False, False),
    "l2_conv2_3x3_128": (4, 100, 168, 128, 128, 3, 1, 1, False, False, True, True),
    "l2_conv2dgrad_3x3_128_mask": (4, 100, 168, 128, 128, 3, 1, 1, False, True, False, False),
    "pred_cls_3x3_256_80_fp32": (4, 100, 168, 256, 80, 3, 1, 1, False, False, True, False),
    # RLA_ResNet state path at C2 / C3 (engine_rla.py): conv_out 4*planes -> 64-channel state rows, the h half of conv1,
    # recurrent_conv 3x3 on the state rows, conv1's x half with the h half as the residual operand
    "rla_c2_conv_out_256_64": (4, 200, 336, 256, 64, 1, 1, 0, False, False, False, False),
    "rla_c2_conv1h_64_64": (4, 200, 336, 64, 64, 1, 1, 0, False, False, False, False),
    "rla_c2_recurrent_3x3_64": (4, 200, 336, 64, 64, 3, 1, 1, False, False, False, False),
    "rla_c2_conv1x_256_64_res": (4, 200, 336, 256, 64, 1, 1, 0, True, False, True, True),
    "rla_c3_conv_out_512_64": (4, 100, 168, 512, 64, 1, 1, 0, False, False, False, False),
    "rla_c3_recurrent_3x3_64": (4, 100, 168, 64, 64, 3, 1, 1, False, False, False, False),
}


def run(name, N, H, W, Ci, Co, R, stride, pad, res, mask, affine, relu, iters=20):
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    bytes_per = 2 * (N * H * W * Ci + N * Ho * Wo * Co * (1 + int(res) + int(mask)))
    nset = max(2, int(400e6 // bytes_per) + 1)
    w = (torch.randn(R * R, Co, Ci, device=dev) * 0.05).to(BF)
    scale = torch.rand(Co, device=dev) + 0.5
    shift = torch.randn(Co, device=dev)
    plans = []
    for _ in range(nset):
        x = torch.randn(N, H, W, Ci, device=dev).to(BF)
        fp32 = name.endswith("fp32")
        y = torch.zeros(N, Ho, Wo, Co, dtype=torch.float32 if fp32 else BF, device=dev)
        seg = dict(x=x, w=w, y=y, N=N, H=H, W=W, Cin=Ci, Cout=Co, cout_pad=Co, R=R, S=R, stride=stride, pad=pad, ldc=Co,
                   relu_nch=Co if relu else 0, out_fp32=int(fp32))
        if affine:
            seg.update(shift=shift)  # BN scale is folded into the packed weights, as in the engine
        if res:
            seg["residual"] = torch.randn(N, Ho, Wo, Co, device=dev).to(BF)
        if mask:
            seg["relu_mask"] = torch.randn(N, Ho, Wo, Co, device=dev).to(BF)
        plans.append(ConvPlan([seg], name))
    for p in plans:
        p.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        plans[i % nset].run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    print(f"{name:34s} {us:8.1f} us  {plans[0].flops / us / 1e6:8.1f} TF/s  {bytes_per / us / 1e3:8.1f} GB/s "
          f"({bytes_per / 1e6:.0f} MB, {nset} buffer sets)", flush=True)


if __name__ == "__main__":
    print("lib:", L.LIB_PATH)
    names = sys.argv[1:] or list(CASES)
    for n in names:
        run(n, *CASES[n])
