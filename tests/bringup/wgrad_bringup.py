"""GPU bring-up of the tcgen05 wgrad kernel + multi-segment fprop (writes gpurun_out/wgrad_bringup.log)."""
import ctypes as C
import os
import sys

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


def wgrad_case(name, levels, Cin, Cout, R, stride, pad, ldy=None, timing=False, seed=0):
    """levels: list of (N,H,W) segments sharing one dw."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    ldy = ldy or ((Cout + 63) // 64 * 64)
    dw = torch.zeros(R * R, Cout, Cin, dtype=torch.float32, device=dev)
    segs = (L.WgradSeg * len(levels))()
    keep = []
    ref = torch.zeros(Cout, Cin, R, R, device=dev)
    for i, (N, H, W) in enumerate(levels):
        Ho = (H + 2 * pad - R) // stride + 1
        Wo = (W + 2 * pad - R) // stride + 1
        x = torch.randn(N, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
        dy = torch.zeros(N, Ho, Wo, ldy, dtype=torch.bfloat16, device=dev)
        dy[..., :Cout] = torch.randn(N, Ho, Wo, Cout, generator=g).to(dev).to(torch.bfloat16)
        keep += [x, dy]
        s = segs[i]
        s.x, s.dy, s.dw = x.data_ptr(), dy.data_ptr(), dw.data_ptr()
        s.N, s.H, s.W, s.Cin, s.Cout, s.ldy, s.dw_rows = N, H, W, Cin, Cout, ldy, Cout
        s.R, s.S, s.stride, s.pad = R, R, stride, pad
        ref += torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (Cout, Cin, R, R),
                                           dy[..., :Cout].float().permute(0, 3, 1, 2), stride=stride, padding=pad)
    plan = C.c_void_p()
    L.check(L.lib.dslb_wgrad_plan_create(segs, len(levels), C.byref(plan)), name)
    L.check(L.lib.dslb_wgrad_plan_run(plan, L.cur_stream()), name)
    torch.cuda.synchronize()
    got = dw.reshape(R, R, Cout, Cin).permute(2, 3, 0, 1)
    err = (got - ref).abs().max().item()
    den = ref.abs().max().item() + 1e-12
    ok = err / den < 1e-4
    line = f"[{'OK' if ok else 'FAIL'}] {name}: max_abs_err={err:.3e} rel={err / den:.3e}"
    if timing:
        for _ in range(3):
            L.lib.dslb_wgrad_plan_run(plan, L.cur_stream())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            L.lib.dslb_wgrad_plan_run(plan, L.cur_stream())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = L.lib.dslb_wgrad_plan_flops(plan)
        line += f" | {ms * 1e3:.1f} us, {fl / ms / 1e9:.1f} TFLOP/s"
    log(line)
    L.lib.dslb_wgrad_plan_destroy(plan)
    return ok


def pack_w(w, cout_pad):
    O, I, R, S = w.shape
    out = torch.zeros(R * S, cout_pad, I, dtype=torch.bfloat16, device=w.device)
    out[:, :O, :] = w.permute(2, 3, 0, 1).reshape(R * S, O, I).to(torch.bfloat16)
    return out.contiguous()


def multiseg_fprop(name, timing=False, use_shift=True, use_stats=True, nlev=5, ntow=2):
    """FCOSHead tower layer: 5 levels x 2 towers in ONE launch, with bias + GroupNorm statistics."""
    g = torch.Generator(device="cpu").manual_seed(1)
    N = 4
    levels = [(100, 168), (50, 84), (25, 42), (13, 21), (7, 11)][:nlev]
    ws = [(torch.randn(256, 256, 3, 3, generator=g) / 48.0).to(dev) for _ in range(2)]
    bs = [torch.randn(256, generator=g).to(dev) * 0.1 for _ in range(2)]
    wps = [pack_w(w, 256) for w in ws]
    segs = (L.ConvSeg * 10)()
    keep, outs = [], []
    k = 0
    for t in range(ntow):
        for (H, W) in levels:
            x = torch.randn(N, H, W, 256, generator=g).to(dev).to(torch.bfloat16)
            y = torch.empty(N, H, W, 256, dtype=torch.bfloat16, device=dev)
            st = torch.zeros(N, 32, 32, dtype=torch.float64, device=dev)
            keep += [x, y, st]
            outs.append((t, x, y, st))
            s = segs[k]
            k += 1
            s.x, s.w, s.y = x.data_ptr(), wps[t].data_ptr(), y.data_ptr()
            s.shift = bs[t].data_ptr() if use_shift else None
            s.gn_stats = st.data_ptr() if use_stats else None
            s.N, s.H, s.W, s.Cin, s.Cout, s.cout_pad = N, H, W, 256, 256, 256
            s.R, s.S, s.stride, s.pad = 3, 3, 1, 1
            s.ldc, s.out_fp32, s.relu_nch, s.gn_cpg = 256, 0, 0, 8
    plan = C.c_void_p()
    L.check(L.lib.dslb_conv_plan_create(segs, k, C.byref(plan)), name)
    L.check(L.lib.dslb_conv_plan_run(plan, L.cur_stream()), name)
    torch.cuda.synchronize()
    ok = True
    worst = 0.0
    worst_s = 0.0
    for (t, x, y, st) in outs:
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), ws[t].to(torch.bfloat16).float(), bias=bs[t] if use_shift else None, padding=1)
        ref = ref.permute(0, 2, 3, 1)
        e = ((y.float() - ref).abs().max() / ref.abs().max()).item()
        worst = max(worst, e)
        r = y.float().reshape(N, -1, 32, 8).double()
        s1, s2 = r.sum(dim=(1, 3)), (r * r).sum(dim=(1, 3))
        es = max(((st[..., 0] - s1).abs().max() / s1.abs().max()).item(),
                 ((st[..., 1] - s2).abs().max() / s2.abs().max()).item())
        worst_s = max(worst_s, es)
    ok = worst < 1e-2 and (worst_s < 5e-3 or not use_stats)
    line = f"[{'OK' if ok else 'FAIL'}] {name}: worst rel={worst:.3e} worst stats rel={worst_s:.3e}"
    if timing:
        for _ in range(3):
            L.lib.dslb_conv_plan_run(plan, L.cur_stream())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            L.lib.dslb_conv_plan_run(plan, L.cur_stream())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = L.lib.dslb_conv_plan_flops(plan)
        line += f" | {ms * 1e3:.1f} us, {fl / ms / 1e9:.1f} TFLOP/s"
    log(line)
    L.lib.dslb_conv_plan_destroy(plan)
    return ok


def main():
    log("device:", torch.cuda.get_device_name(0))
    allok = True
    tests = [
        lambda: wgrad_case("wgrad 1x1 64->128 tiny", [(1, 8, 16)], 64, 128, 1, 1, 0),
        lambda: wgrad_case("wgrad 1x1 128->64 (M half)", [(2, 13, 21)], 128, 64, 1, 1, 0),
        lambda: wgrad_case("wgrad 3x3 256->256 P5", [(2, 25, 42)], 256, 256, 3, 1, 1),
        lambda: wgrad_case("wgrad 3x3 256->80 ldy128", [(2, 25, 42)], 256, 80, 3, 1, 1, ldy=128),
        lambda: wgrad_case("wgrad 3x3 256->5 ldy64", [(2, 25, 42)], 256, 5, 3, 1, 1, ldy=64),
        lambda: wgrad_case("wgrad 1x1 s2 512->256", [(2, 50, 84)], 512, 256, 1, 2, 0),
        lambda: wgrad_case("wgrad 3x3 s2 256->256", [(2, 25, 42)], 256, 256, 3, 2, 1),
        lambda: wgrad_case("wgrad 1x1 1024->2048", [(2, 25, 42)], 1024, 2048, 1, 1, 0),
        lambda: wgrad_case("wgrad 3x3 256->256 5 levels shared", [(2, 50, 84), (2, 25, 42), (2, 13, 21), (2, 7, 11)],
                           256, 256, 3, 1, 1),
        lambda: wgrad_case("wgrad 3x3 256->256 head bs4 5 levels (timing)",
                           [(4, 100, 168), (4, 50, 84), (4, 25, 42), (4, 13, 21), (4, 7, 11)], 256, 256, 3, 1, 1,
                           timing=True),
        lambda: wgrad_case("wgrad 1x1 256->1024 C4 bs4 (timing)", [(4, 50, 84)], 256, 1024, 1, 1, 0, timing=True),
        lambda: multiseg_fprop("fprop head layer 5x2 shift+stats", timing=True),
        lambda: multiseg_fprop("fprop head layer 5x2 plain", timing=True, use_shift=False, use_stats=False),
        lambda: multiseg_fprop("fprop head layer 5x2 shift only", timing=True, use_shift=True, use_stats=False),
        lambda: multiseg_fprop("fprop head layer 5x2 stats only", timing=True, use_shift=False, use_stats=True),
        lambda: multiseg_fprop("fprop head layer 1x1 (P3 only) plain", timing=True, use_shift=False, use_stats=False, nlev=1, ntow=1),
        lambda: multiseg_fprop("fprop head layer 1x2 (P3 x 2 towers) plain", timing=True, use_shift=False, use_stats=False, nlev=1, ntow=2),
        lambda: multiseg_fprop("fprop head layer 5x1 plain", timing=True, use_shift=False, use_stats=False, nlev=5, ntow=1),
    ]
    for t in tests:
        try:
            allok = t() and allok
        except Exception as e:
            log(f"[EXC ] {type(e).__name__}: {e}")
            allok = False
            try:
                torch.cuda.synchronize()
            except Exception as e2:
                log("CUDA context is dead:", e2)
                break
    log("ALL OK" if allok else "SOME FAILED")
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/wgrad_bringup.log", "w") as f:
        f.write("\n".join(LOG) + "\n")
    return 0 if allok else 1


if __name__ == "__main__":
    sys.exit(main())
