"""GPU bring-up of the tcgen05 implicit-GEMM conv (run under gpurun; writes gpurun_out/conv_bringup.log).

Compares dslb_conv_plan_* with torch F.conv2d (fp32 math on the same bf16-rounded inputs).
"""
import ctypes as C
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dsl_b200 import _lib as L  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = "cuda"
LOG = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.append(s)


def pack_w(w, cout_pad):
    """OIHW fp32 -> [R*S][cout_pad][Cin] bf16"""
    O, I, R, S = w.shape
    out = torch.zeros(R * S, cout_pad, I, dtype=torch.bfloat16, device=w.device)
    out[:, :O, :] = w.permute(2, 3, 0, 1).reshape(R * S, O, I).to(torch.bfloat16)
    return out.contiguous()


def run_case(name, N, H, W, Cin, Cout, R, stride, pad, out_fp32=False, relu=False, residual=False, affine=False,
             mask=False, gn=False, cout_pad=None, ldc=None, timing=False, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(N, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(dev)
    wb = w.to(torch.bfloat16)
    cout_pad = cout_pad or ((Cout + 15) // 16 * 16)
    wp = pack_w(w, cout_pad)
    Ho = (H + 2 * pad - R) // stride + 1
    Wo = (W + 2 * pad - R) // stride + 1
    ldc = ldc or Cout
    ydt = torch.float32 if out_fp32 else torch.bfloat16
    y = torch.full((N, Ho, Wo, ldc), -7.0, dtype=ydt, device=dev)
    scale = (torch.rand(Cout, generator=g) + 0.5).to(dev) if affine else None
    shift = torch.randn(Cout, generator=g).to(dev) if affine else None
    res = torch.randn(N, Ho, Wo, ldc, generator=g).to(dev).to(torch.bfloat16) if residual else None
    msk = torch.randn(N, Ho, Wo, ldc, generator=g).to(dev).to(torch.bfloat16) if mask else None
    groups = 32 if gn else 0
    stats = torch.zeros(N, groups, 32, dtype=torch.float64, device=dev) if gn else None

    seg = L.ConvSeg()
    seg.x, seg.w, seg.y = x.data_ptr(), wp.data_ptr(), y.data_ptr()
    seg.residual = res.data_ptr() if residual else None
    seg.relu_mask = msk.data_ptr() if mask else None
    seg.scale = scale.data_ptr() if affine else None
    seg.shift = shift.data_ptr() if affine else None
    seg.gn_stats = stats.data_ptr() if gn else None
    seg.N, seg.H, seg.W, seg.Cin, seg.Cout, seg.cout_pad = N, H, W, Cin, Cout, cout_pad
    seg.R, seg.S, seg.stride, seg.pad = R, R, stride, pad
    seg.ldc, seg.out_fp32, seg.relu_nch = ldc, int(out_fp32), (Cout if relu else 0)
    seg.gn_cpg = Cout // groups if gn else 0
    plan = C.c_void_p()
    L.check(L.lib.dslb_conv_plan_create(C.byref(seg), 1, C.byref(plan)), name)
    L.check(L.lib.dslb_conv_plan_run(plan, L.cur_stream()), name)
    torch.cuda.synchronize()

    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wb.float(), stride=stride, padding=pad)
    if affine:
        ref = ref * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    ref = ref.permute(0, 2, 3, 1)
    if residual:
        ref = ref + res[..., :Cout].float()
    if relu:
        ref = ref.clamp_min(0)
    if mask:
        ref = torch.where(msk[..., :Cout].float() > 0, ref, torch.zeros_like(ref))
    got = y[..., :Cout].float()
    err = (got - ref).abs().max().item()
    den = ref.abs().max().item() + 1e-12
    ok = err / den < (2e-5 if out_fp32 else 1e-2)
    extra = ""
    if ldc > Cout:
        untouched = bool((y[..., Cout:].float() == -7.0).all().item())
        extra += f" pad_untouched={untouched}"
        ok = ok and untouched
    if gn:
        r = y[..., :Cout].float().reshape(N, Ho * Wo, groups, Cout // groups).double()
        s1 = r.sum(dim=(1, 3))
        s2 = (r * r).sum(dim=(1, 3))
        e1 = ((stats[..., 0] - s1).abs().max() / (s1.abs().max() + 1e-9)).item()
        e2 = ((stats[..., 1] - s2).abs().max() / (s2.abs().max() + 1e-9)).item()
        extra += f" gn_sum_err={e1:.2e} gn_sq_err={e2:.2e}"
        ok = ok and e1 < 5e-3 and e2 < 5e-3
    line = f"[{'OK' if ok else 'FAIL'}] {name}: max_abs_err={err:.3e} rel={err / den:.3e}{extra}"
    if timing:
        for _ in range(3):
            L.lib.dslb_conv_plan_run(plan, L.cur_stream())
        torch.cuda.synchronize()
        e0, e1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            L.lib.dslb_conv_plan_run(plan, L.cur_stream())
        e1_.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1_) / iters
        fl = L.lib.dslb_conv_plan_flops(plan)
        line += f" | {ms * 1e3:.1f} us, {fl / ms / 1e9:.1f} TFLOP/s"
    log(line)
    L.lib.dslb_conv_plan_destroy(plan)
    return ok


def main():
    log("device:", torch.cuda.get_device_name(0), "dslb version", L.lib.dslb_version())
    allok = True
    cases = [
        dict(name="1x1 64->64 tiny", N=1, H=8, W=16, Cin=64, Cout=64, R=1, stride=1, pad=0),
        dict(name="1x1 64->256 fp32out", N=2, H=25, W=42, Cin=64, Cout=256, R=1, stride=1, pad=0, out_fp32=True),
        dict(name="3x3 256->256 P5", N=2, H=25, W=42, Cin=256, Cout=256, R=3, stride=1, pad=1),
        dict(name="3x3 256->256 P5 fp32", N=2, H=25, W=42, Cin=256, Cout=256, R=3, stride=1, pad=1, out_fp32=True),
        dict(name="3x3 256->256 P7 small", N=1, H=7, W=11, Cin=256, Cout=256, R=3, stride=1, pad=1, out_fp32=True),
        dict(name="3x3 s2 256->256 P6", N=2, H=25, W=42, Cin=256, Cout=256, R=3, stride=2, pad=1, out_fp32=True),
        dict(name="1x1 s2 512->256", N=2, H=50, W=84, Cin=512, Cout=256, R=1, stride=2, pad=0, out_fp32=True),
        dict(name="1x1 256->1024 (4 n-tiles)", N=2, H=25, W=42, Cin=256, Cout=1024, R=1, stride=1, pad=0),
        dict(name="3x3 256->80 cls", N=2, H=25, W=42, Cin=256, Cout=80, R=3, stride=1, pad=1, out_fp32=True),
        dict(name="3x3 256->5 regctr ldc8", N=2, H=25, W=42, Cin=256, Cout=5, R=3, stride=1, pad=1, out_fp32=True,
             ldc=8),
        dict(name="3x3 affine+res+relu", N=2, H=25, W=42, Cin=128, Cout=128, R=3, stride=1, pad=1, affine=True,
             residual=True, relu=True),
        dict(name="1x1 mask", N=2, H=25, W=42, Cin=128, Cout=64, R=1, stride=1, pad=0, mask=True),
        dict(name="3x3 gn stats", N=3, H=13, W=21, Cin=256, Cout=256, R=3, stride=1, pad=1, affine=True, gn=True),
        dict(name="7x7 s2 64->64", N=1, H=32, W=48, Cin=64, Cout=64, R=7, stride=2, pad=3, out_fp32=True),
        dict(name="3x3 256->256 P3 bs4 (timing)", N=4, H=100, W=168, Cin=256, Cout=256, R=3, stride=1, pad=1,
             timing=True),
        dict(name="1x1 256->1024 C4 bs4 (timing)", N=4, H=50, W=84, Cin=256, Cout=1024, R=1, stride=1, pad=0,
             timing=True),
        dict(name="1x1 64->256 C2 bs4 (timing)", N=4, H=200, W=336, Cin=64, Cout=256, R=1, stride=1, pad=0,
             timing=True),
    ]
    for c in cases:
        try:
            allok = run_case(**c) and allok
        except Exception as e:  # keep going to collect as much evidence as possible per GPU call
            log(f"[EXC ] {c['name']}: {type(e).__name__}: {e}")
            allok = False
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                log("CUDA context is dead:", e2)
                break
    log("ALL OK" if allok else "SOME FAILED")
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/conv_bringup.log", "w") as f:
        f.write("\n".join(LOG) + "\n")
    return 0 if allok else 1


if __name__ == "__main__":
    t = time.time()
    rc = main()
    print("elapsed", time.time() - t)
    sys.exit(rc)
