"""Round-2 GPU parity net: every implicit-GEMM code path on IDENTICAL bf16-representable inputs against torch fp32
(`F.conv2d` + autograd on the GPU), one kernel at a time, so that a dgrad / wgrad / epilogue bug of a few percent cannot
hide behind the bf16 emulation floor of the end-to-end tests.

Bars: outputs that are rounded to bf16 once — 4e-3 of the tensor's max (one bf16 ulp is 2^-8 relative); fp32 outputs
(weight gradients, fp32 predictor maps) — 1e-4. Paths covered (file: dsl_b200/csrc/conv_igemm.cu, conv_wgrad.cu,
stem.cu, rla.cu): general 8-warp kernel, 16-warp `fast4` kernel, residual on the tensor core (`res_mma`), TMA-loaded
mask / residual tiles (`aux_kind`), stride-2 1x1 fprop, scatter2 dgrad, zero-upsample 3x3 stride-2 dgrad, split-K
tails (npix % 128 != 0), wgrad for 1x1 / 3x3 / 7x7-free shapes and strides, the fused 7x7 stem, `bn_grad_plan`.
Also: the device top-k (radix select) against torch.topk, the 20-class (VOC) teacher decode, the candidate-cap flag.
"""
import ctypes as C
import zlib

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.golden import inputs as GI

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def _bf(t):
    return t.bfloat16().float()


def _nhwc(t, dtype=torch.bfloat16):
    return t.permute(0, 2, 3, 1).contiguous().to(dtype).to(DEV)


def _pack(w, transpose, rows_pad=None, cols_pad=None):
    from dsl_b200 import _lib as L
    O, I, R, S = w.shape
    if transpose:
        rows_pad, cols_pad = rows_pad or I, cols_pad or O
    else:
        rows_pad, cols_pad = rows_pad or O, cols_pad or I
    out = torch.zeros(R * S, rows_pad, cols_pad, dtype=torch.bfloat16, device=DEV)
    L.check(L.lib.dslb_pack_weight(L.ptr(w.to(DEV)), L.ptr(out), O, I, R, S, rows_pad, cols_pad, None, int(transpose),
                                   L.cur_stream()), "pack")
    return out


# (name, N, H, W, Cin, Cout, k, stride, pad, residual, relu, mask)
FPROP_CASES = [
    ("1x1 256->64 relu (general kernel, N=64 tile)", 2, 50, 84, 256, 64, 1, 1, 0, False, True, False),
    ("1x1 64->256 residual relu (res_mma + fast4)", 2, 50, 84, 64, 256, 1, 1, 0, True, True, False),
    ("1x1 64->256 plain (fast4)", 2, 50, 84, 64, 256, 1, 1, 0, False, False, False),
    ("1x1 256->1024 residual relu, tail tile", 1, 25, 42, 256, 1024, 1, 1, 0, True, True, False),
    ("1x1 128->512 mask (TMA mask tile, aux 2)", 2, 25, 42, 128, 512, 1, 1, 0, False, False, True),
    ("1x1 128->512 residual + mask (res_mma + aux 2)", 2, 25, 42, 128, 512, 1, 1, 0, True, False, True),
    ("1x1 stride 2 256->128 relu", 2, 50, 84, 256, 128, 1, 2, 0, False, True, False),
    ("1x1 stride 2 256->512 (downsample)", 2, 50, 84, 256, 512, 1, 2, 0, False, False, False),
    ("3x3 64->64 relu (layer1 conv2)", 2, 50, 84, 64, 64, 3, 1, 1, False, True, False),
    ("3x3 128->128 relu, odd map", 2, 25, 43, 128, 128, 3, 1, 1, False, True, False),
    ("3x3 stride 2 256->256 (P6)", 2, 25, 42, 256, 256, 3, 2, 1, False, False, False),
    ("3x3 256->256 mask (dgrad-shaped)", 1, 13, 21, 256, 256, 3, 1, 1, False, False, True),
    ("3x3 512->512 relu, tiny map (layer4 conv2)", 2, 7, 11, 512, 512, 3, 1, 1, False, True, False),
    # halo-tile kernel (conv_halo.cu): several patches per CTA (ring wrap-around), resident and streamed weights,
    # patches hanging over the right / bottom edge, every Cin x Cout combination it accepts
    ("3x3 64->64 relu, 1100 patches (halo, resident weights)", 4, 100, 168, 64, 64, 3, 1, 1, False, True, False),
    ("3x3 128->128 relu, 600 patches (halo, streamed weights)", 4, 100, 170, 128, 128, 3, 1, 1, False, True, False),
    ("3x3 64->128 no relu (halo)", 3, 40, 60, 64, 128, 3, 1, 1, False, False, False),
    ("3x3 128->64 mask (halo)", 3, 33, 41, 128, 64, 3, 1, 1, False, False, True),
    ("3x3 64->64 no shift path, W = 8 (halo)", 5, 70, 8, 64, 64, 3, 1, 1, False, False, False),
]


@pytest.mark.parametrize("case", FPROP_CASES, ids=[c[0] for c in FPROP_CASES])
def test_conv_fprop_identical_inputs(case):
    from dsl_b200.engine import ConvPlan
    _, N, H, W, Ci, Co, k, stride, pad, use_res, relu, use_mask = case
    g = torch.Generator().manual_seed(zlib.crc32(case[0].encode()) % 1000)
    x = _bf(torch.randn(N, Ci, H, W, generator=g))
    w = _bf(torch.randn(Co, Ci, k, k, generator=g) * (1.0 / (Ci * k * k) ** 0.5))
    shift = torch.randn(Co, generator=g) * 0.1
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    res = _bf(torch.randn(N, Co, Ho, Wo, generator=g)) if use_res else None
    mask = _bf(torch.randn(N, Co, Ho, Wo, generator=g)) if use_mask else None
    ref = F.conv2d(x.to(DEV), w.to(DEV), stride=stride, padding=pad) + shift.to(DEV).view(1, -1, 1, 1)
    if use_res:
        ref = ref + res.to(DEV)
    if relu:
        ref = F.relu(ref)
    if use_mask:
        ref = ref * (mask.to(DEV) > 0)
    y = torch.full((N, Ho, Wo, Co), float("nan"), dtype=torch.bfloat16, device=DEV)
    seg = dict(x=_nhwc(x), w=_pack(w, False), y=y, N=N, H=H, W=W, Cin=Ci, Cout=Co, cout_pad=Co, R=k, S=k,
               stride=stride, pad=pad, ldc=Co, shift=shift.to(DEV), relu_nch=Co if relu else 0)
    if use_res:
        seg["residual"] = _nhwc(res)
    if use_mask:
        seg["relu_mask"] = _nhwc(mask)
    ConvPlan([seg], "fprop").run()
    torch.cuda.synchronize()
    got = y.permute(0, 3, 1, 2).float()
    assert torch.isfinite(got).all(), "unwritten output elements"
    e = _rel(got, ref)
    print(f"{case[0]}: rel {e:.2e}")
    assert e < 4e-3


# (name, N, Hin, Win, Cin, Cout, k, stride, pad, residual(accumulate), mask)
DGRAD_CASES = [
    ("1x1 s1 64<-256 (conv3 dgrad, K=256)", 2, 50, 84, 64, 256, 1, 1, 0, False, True),
    ("1x1 s1 256<-64 residual + mask (conv1 dgrad)", 2, 50, 84, 256, 64, 1, 1, 0, True, True),
    ("1x1 s1 512<-128 residual + mask, tail", 1, 25, 43, 512, 128, 1, 1, 0, True, True),
    ("1x1 s2 256<-128 scatter accumulate", 2, 50, 84, 256, 128, 1, 2, 0, True, False),
    ("1x1 s2 256<-512 scatter accumulate (downsample)", 2, 26, 42, 256, 512, 1, 2, 0, True, False),
    ("3x3 s1 128<-128 mask", 2, 25, 42, 128, 128, 3, 1, 1, False, True),
    ("3x3 s2 256<-256 zero-upsample", 2, 25, 42, 256, 256, 3, 2, 1, False, False),
    ("3x3 s2 256<-256 zero-upsample, odd map + residual", 2, 13, 21, 256, 256, 3, 2, 1, True, False),
    ("3x3 s1 128<-128 mask, 600 patches (halo)", 4, 100, 168, 128, 128, 3, 1, 1, False, True),
]


@pytest.mark.parametrize("case", DGRAD_CASES, ids=[c[0] for c in DGRAD_CASES])
def test_conv_dgrad_identical_inputs(case):
    from dsl_b200 import _lib as L
    from dsl_b200.engine import ConvPlan
    _, N, H, W, Ci, Co, k, stride, pad, use_res, use_mask = case
    g = torch.Generator().manual_seed(zlib.crc32(case[0].encode()) % 1000 + 1)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    w = _bf(torch.randn(Co, Ci, k, k, generator=g) * (1.0 / (Co * k * k) ** 0.5))
    dy = _bf(torch.randn(N, Co, Ho, Wo, generator=g))
    res = _bf(torch.randn(N, Ci, H, W, generator=g)) if use_res else None
    mask = _bf(torch.randn(N, Ci, H, W, generator=g)) if use_mask else None
    xr = torch.zeros(N, Ci, H, W, device=DEV, requires_grad=True)
    F.conv2d(xr, w.to(DEV), stride=stride, padding=pad).backward(dy.to(DEV))
    ref = xr.grad
    if use_res:
        ref = ref + res.to(DEV)
    if use_mask:
        ref = ref * (mask.to(DEV) > 0)
    wpT = _pack(w, True)
    dy_d = _nhwc(dy)
    dx = _nhwc(res) if use_res else torch.full((N, H, W, Ci), float("nan"), dtype=torch.bfloat16, device=DEV)
    seg = dict(w=wpT, y=dx, N=N, Cin=Co, Cout=Ci, cout_pad=Ci, R=k, S=k, stride=1, pad=k - 1 - pad, ldc=Ci)
    if stride == 2 and k == 1:
        if not use_res:
            dx.zero_()
        seg.update(x=dy_d, H=Ho, W=Wo, scatter2=1, Hs=H, Ws=W)
    elif stride == 2:
        up = torch.full((N, H, W, Co), float("nan"), dtype=torch.bfloat16, device=DEV)
        L.check(L.lib.dslb_zero_upsample2(L.ptr(dy_d), L.ptr(up), N, Ho, Wo, H, W, Co, L.cur_stream()), "zero_upsample2")
        seg.update(x=up, H=H, W=W)
    else:
        seg.update(x=dy_d, H=Ho, W=Wo)
    if use_res:
        seg["residual"] = dx
    if use_mask:
        seg["relu_mask"] = _nhwc(mask)
    ConvPlan([seg], "dgrad").run()
    torch.cuda.synchronize()
    got = dx.permute(0, 3, 1, 2).float()
    assert torch.isfinite(got).all()
    e = _rel(got, ref)
    print(f"{case[0]}: rel {e:.2e}")
    assert e < 4e-3


# (name, N, H, W, Cin, Cout, k, stride, pad)
WGRAD_CASES = [
    ("1x1 256->64", 2, 50, 84, 256, 64, 1, 1, 0),
    ("1x1 64->256, npix % 128 != 0", 1, 25, 43, 64, 256, 1, 1, 0),
    ("1x1 stride 2 256->512", 2, 50, 84, 256, 512, 1, 2, 0),
    ("3x3 128->128", 2, 25, 42, 128, 128, 3, 1, 1),
    ("3x3 stride 2 256->256", 2, 25, 43, 256, 256, 3, 2, 1),
    ("3x3 256->80 (conv_cls, ldy 128)", 2, 13, 21, 256, 80, 3, 1, 1),
    ("3x3 512->512 tiny map", 2, 7, 11, 512, 512, 3, 1, 1),
    # CTA-pair wgrad (Cout % 256 == 0): tower shape, several Cin tiles, several Cout pairs, Cin = 128 (one box per CTA)
    ("3x3 256->256 tower level (CTA pairs)", 4, 50, 84, 256, 256, 3, 1, 1),
    ("1x1 1024->256, four Cin tiles (CTA pairs)", 2, 50, 84, 1024, 256, 1, 1, 0),
    ("1x1 256->1024, four Cout pairs... two pairs (CTA pairs)", 2, 25, 43, 256, 1024, 1, 1, 0),
    ("3x3 128->256 (CTA pairs, one X box per CTA)", 2, 25, 42, 128, 256, 3, 1, 1),
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=[c[0] for c in WGRAD_CASES])
def test_conv_wgrad_identical_inputs(case):
    from dsl_b200.engine import WgradPlan
    _, N, H, W, Ci, Co, k, stride, pad = case
    g = torch.Generator().manual_seed(zlib.crc32(case[0].encode()) % 1000 + 2)
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    x = _bf(torch.randn(N, Ci, H, W, generator=g))
    dy = _bf(torch.randn(N, Co, Ho, Wo, generator=g))
    wr = torch.zeros(Co, Ci, k, k, device=DEV, requires_grad=True)
    F.conv2d(x.to(DEV), wr, stride=stride, padding=pad).backward(dy.to(DEV))
    ldy = (Co + 63) // 64 * 64
    dy_d = torch.zeros(N, Ho, Wo, ldy, dtype=torch.bfloat16, device=DEV)
    dy_d[..., :Co] = _nhwc(dy)
    dwp = torch.zeros(k * k, Co, Ci, dtype=torch.float32, device=DEV)
    WgradPlan([dict(x=_nhwc(x), dy=dy_d, dw=dwp, N=N, H=H, W=W, Cin=Ci, Cout=Co, ldy=ldy, dw_rows=Co, R=k, S=k,
                    stride=stride, pad=pad)], "wgrad").run()
    torch.cuda.synchronize()
    e = _rel(dwp.view(k, k, Co, Ci).permute(2, 3, 0, 1), wr.grad)
    print(f"{case[0]}: rel {e:.2e}")
    assert e < 1e-4


def test_predictor_fp32_output_identical_inputs():
    """The fp32 direct-store epilogue (conv_cls N=80 and the fused reg + centerness N=5 -> 16 with per-channel scale /
    shift and ReLU on the first 4 channels), fcos_head.py:139-168."""
    from dsl_b200.engine import ConvPlan
    g = torch.Generator().manual_seed(11)
    N, H, W, Ci = 2, 25, 42, 256
    x = _bf(torch.randn(N, Ci, H, W, generator=g))
    for Co, ld, relu_nch, with_scale in ((80, 80, 0, False), (5, 8, 4, True)):
        w = _bf(torch.randn(Co, Ci, 3, 3, generator=g) * 0.02)
        shift = torch.randn(Co, generator=g) * 0.1
        scale = (torch.rand(Co, generator=g) + 0.5) if with_scale else None
        ref = F.conv2d(x.to(DEV), w.to(DEV), padding=1)
        if with_scale:
            ref = ref * scale.to(DEV).view(1, -1, 1, 1)
        ref = ref + shift.to(DEV).view(1, -1, 1, 1)
        if relu_nch:
            ref = torch.cat([F.relu(ref[:, :relu_nch]), ref[:, relu_nch:]], 1)
        y = torch.zeros(N, H, W, ld, dtype=torch.float32, device=DEV)
        seg = dict(x=_nhwc(x), w=_pack(w, False, rows_pad=(Co + 15) // 16 * 16), y=y, N=N, H=H, W=W, Cin=Ci, Cout=Co,
                   cout_pad=(Co + 15) // 16 * 16, R=3, S=3, stride=1, pad=1, ldc=ld, out_fp32=1, relu_nch=relu_nch,
                   shift=shift.to(DEV))
        if with_scale:
            seg["scale"] = scale.to(DEV)
        ConvPlan([seg], "pred").run()
        torch.cuda.synchronize()
        e = _rel(y[..., :Co].permute(0, 3, 1, 2), ref)
        print(f"predictor Cout={Co}: rel {e:.2e}")
        assert e < 1e-4   # fp32 accumulate, fp32 store: no bf16 rounding on the way out


def test_gn_stats_epilogue_identical_inputs():
    """GroupNorm statistics accumulated by the conv epilogue (fp64 atomics over per-tile fp32 partial sums of the fp32
    conv + bias values, i.e. BEFORE the bf16 rounding of the stored map), several segments in one launch incl. a map
    smaller than one tile and a tile that straddles two images."""
    from dsl_b200 import _lib as L
    from dsl_b200.engine import ConvPlan
    g = torch.Generator().manual_seed(12)
    Ci = Co = 256
    w = _bf(torch.randn(Co, Ci, 3, 3, generator=g) * 0.02)
    wp = _pack(w, False)
    bias = (torch.randn(Co, generator=g) * 0.1).to(DEV)
    segs, keep = [], []
    for (N, H, W) in ((2, 25, 42), (2, 7, 11), (3, 13, 21)):
        x = _bf(torch.randn(N, Ci, H, W, generator=g))
        y = torch.zeros(N, H, W, Co, dtype=torch.bfloat16, device=DEV)
        stats = torch.zeros(N, 32, L.GN_STAT_STRIDE, dtype=torch.float64, device=DEV)
        segs.append(dict(x=_nhwc(x), w=wp, y=y, N=N, H=H, W=W, Cin=Ci, Cout=Co, cout_pad=Co, R=3, S=3, stride=1,
                         pad=1, ldc=Co, shift=bias, gn_stats=stats, gn_cpg=8))
        keep.append((x, y, stats))
    ConvPlan(segs, "tower").run()
    torch.cuda.synchronize()
    for x, y, stats in keep:
        ref = F.conv2d(x.to(DEV), w.to(DEV), padding=1) + bias.view(1, -1, 1, 1)
        assert _rel(y.permute(0, 3, 1, 2).float(), ref) < 4e-3
        N = x.shape[0]
        yg = ref.permute(0, 2, 3, 1).double().reshape(N, -1, 32, 8)
        s1, s2 = yg.sum(dim=(1, 3)), (yg * yg).sum(dim=(1, 3))
        e1, e2 = _rel(stats[:, :, 0], s1), _rel(stats[:, :, 1], s2)
        print(f"gn stats {tuple(x.shape)}: sum rel {e1:.2e} sumsq rel {e2:.2e}")
        assert e1 < 2e-5 and e2 < 2e-5


def test_stem_identical_inputs():
    """Fused 7x7/2 stem + frozen BatchNorm + ReLU (resnet.py:597-610, 630-637) vs F.conv2d on a bf16-representable
    image and weights; the BatchNorm is chosen so that its folded scale is exactly 1 (gamma 1, var 1 - eps)."""
    from dsl_b200 import _lib as L
    g = torch.Generator().manual_seed(13)
    for (N, H, W) in ((2, 64, 96), (1, 160, 224)):
        img = _bf(torch.randn(N, 3, H, W, generator=g) * 2)
        w = _bf(torch.randn(64, 3, 7, 7, generator=g) * 0.08)
        gamma = torch.ones(64)
        var = torch.full((64,), 1.0 - 1e-5)
        beta = torch.randn(64, generator=g) * 0.1
        mean = torch.randn(64, generator=g) * 0.1
        Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        scale = gamma / torch.sqrt(var + 1e-5)
        wf = _bf(w * scale.view(-1, 1, 1, 1))   # what the kernel feeds the tensor core
        ref = F.relu(F.conv2d(img.to(DEV), wf.to(DEV), stride=2, padding=3)
                     + (beta - mean * scale).to(DEV).view(1, -1, 1, 1))
        ws = torch.zeros(N, H, W, 4, dtype=torch.bfloat16, device=DEV)
        out = torch.full((N, Ho, Wo, 64), float("nan"), dtype=torch.bfloat16, device=DEV)
        t = [v.to(DEV) for v in (img, w, gamma, beta, mean, var)]
        L.check(L.lib.dslb_stem_conv(L.ptr(t[0]), L.ptr(t[1]), L.ptr(t[2]), L.ptr(t[3]), L.ptr(t[4]), L.ptr(t[5]), 1e-5,
                                     L.ptr(ws), L.ptr(out), N, H, W, L.cur_stream()), "stem")
        torch.cuda.synchronize()
        got = out.permute(0, 3, 1, 2).float()
        assert torch.isfinite(got).all()
        e = _rel(got, ref)
        print(f"stem {N}x{H}x{W}: rel {e:.2e}")
        assert e < 4e-3


def test_bn_grad_plan_identical_inputs():
    """dgamma of a trainable BatchNorm folded into its conv (RLA_ResNet under norm_eval, resnet_rla.py:344-377):
    dgamma = (<dW', W> - mean dbeta) / sqrt(var + eps) vs autograd through conv -> batch_norm(eval)."""
    from dsl_b200 import _lib as L
    from dsl_b200.engine_rla import BnGradPlan
    g = torch.Generator().manual_seed(14)
    N, H, W, Ci, Co = 2, 13, 21, 128, 64
    # bf16-representable operands: cuDNN's TF32 products of the fp32 reference are then exact
    x = _bf(torch.randn(N, Ci, H, W, generator=g)).to(DEV)
    w = _bf(torch.randn(Co, Ci, 3, 3, generator=g) * 0.05).to(DEV)
    gamma = (torch.rand(Co, generator=g) + 0.5).to(DEV).requires_grad_(True)
    beta = torch.zeros(Co, device=DEV, requires_grad=True)
    mean = (torch.randn(Co, generator=g) * 0.2).to(DEV)
    var = (torch.rand(Co, generator=g) + 0.5).to(DEV)
    dy = _bf(torch.randn(N, Co, H, W, generator=g)).to(DEV)
    y = F.batch_norm(F.conv2d(x, w, padding=1), mean, var, gamma, beta, False, 0.0, 1e-5)
    y.backward(dy)
    # the folded conv's weight gradient dW' = d loss / d (W * gamma / sigma), packed [tap][O][I]
    wf = (w * (gamma.detach() / torch.sqrt(var + 1e-5)).view(-1, 1, 1, 1)).requires_grad_(True)
    F.conv2d(x, wf, padding=1).backward(dy)
    dwp = wf.grad.permute(2, 3, 0, 1).reshape(9, Co, Ci).contiguous()
    dbeta = beta.grad.clone()
    dgamma = torch.zeros(Co, device=DEV)
    plan = BnGradPlan([dict(dw0=dwp, w0=w, mean=mean, var=var, dbeta=dbeta, dgamma=dgamma, O=Co, R=3, S=3, I0=Ci,
                            dw_ld0=Ci, w_ld0=Ci, rows0=Co, bn_eps=1e-5)])
    plan.run()
    torch.cuda.synchronize()
    e = _rel(dgamma, gamma.grad)
    print(f"bn_grad_plan dgamma rel {e:.2e}")
    assert e < 1e-4


# ------------------------------------------------------------------------------------------------ teacher decode
def test_topk_points_matches_torch_topk():
    """dslb_fcos_topk_points (radix select) vs torch.topk: same SET per (level, image); ties at the cut go to the lower
    index; levels of different sizes in one launch (fcos_head.py:452-460)."""
    from dsl_b200 import _lib as L
    g = torch.Generator().manual_seed(21)
    B, K = 3, 1000
    ns = [16800, 4200, 1050, 1001]
    scores = [torch.rand(B, n, generator=g).to(DEV) for n in ns]
    scores[1][0, :100] = 0.0          # exact zeros and duplicates below the cut
    scores[2][1] = 0.5                # every score equal: the K lowest indices must win
    scores[3][2, 5::7] = scores[3][2, 3]   # duplicates that may straddle the cut
    sel = [torch.full((B, K), -1, dtype=torch.int64, device=DEV) for _ in ns]
    n_arr = (C.c_int32 * len(ns))(*ns)
    k_arr = (C.c_int32 * len(ns))(*[K] * len(ns))
    sp = (C.c_void_p * len(ns))(*[s.data_ptr() for s in scores])
    op = (C.c_void_p * len(ns))(*[s.data_ptr() for s in sel])
    L.check(L.lib.dslb_fcos_topk_points(sp, op, n_arr, k_arr, len(ns), B, L.cur_stream()), "topk")
    torch.cuda.synchronize()
    for s, o, n in zip(scores, sel, ns):
        for b in range(B):
            got = o[b].cpu().numpy()
            assert got.min() >= 0 and got.max() < n and len(set(got.tolist())) == K, "indices must be valid and distinct"
            v = s[b].cpu().numpy()
            order = np.lexsort((np.arange(n), -v))      # score descending, index ascending among equals
            assert set(got.tolist()) == set(order[:K].tolist())


def test_decode_20_classes_matches_oracle():
    """VOC configs (configs/fcos_semi/voc/*.py: 20 classes): teacher decode + NMS vs the oracle; C % 16 != 0."""
    from dsl_b200.postprocess import TeacherPost
    from oracle import fcos_oracle as O
    B, H, W, Cn = 2, 256, 320, 20
    cls, box, ctr = GI.make_head_outputs(43, B, H, W, num_classes=Cn, train=False, cls_mean=-4.5)
    sizes = GI.level_sizes(H, W)
    shapes, sfs = [(250, 310, 3), (256, 300, 3)], [[1.25] * 4, [0.8] * 4]
    post = TeacherPost(B, sizes, GI.STRIDES, Cn, DEV, nms_pre=1000, score_thr=0.05, iou_thr=0.6, max_per_img=100)
    post.set_meta(shapes, sfs)
    cls_out, rc_out = [], []
    for l, (h, w) in enumerate(sizes):
        cls_out.append(cls[l].permute(0, 2, 3, 1).contiguous().to(DEV))
        rc = torch.zeros(B, h, w, 8, device=DEV)
        rc[..., :4] = box[l].permute(0, 2, 3, 1).to(DEV)
        rc[..., 4] = ctr[l][:, 0].to(DEV)
        rc_out.append(rc)
    post.decode(cls_out, rc_out)
    post.nms()
    torch.cuda.synchronize()
    assert not post.overflowed()
    cand = O.decode_candidates(cls, box, ctr, [s[:2] for s in shapes], sfs, nms_pre=1000, score_thr=0.05)
    for b, (dets, labels) in enumerate(post.results()):
        rb, rs, rl, _ = cand[b]
        rd, rlab = O.multiclass_nms(rb, rs, rl, 0.6, 100)
        assert int(post.cand_counts[b]) == len(rs), "same gated candidate count as the reference's gate"
        assert dets.shape == rd.shape and len(dets) > 10
        np.testing.assert_allclose(dets.numpy(), rd.numpy(), rtol=1e-5, atol=1e-5)
        assert torch.equal(labels, rlab)


def test_candidate_cap_overflow_is_reported():
    """More gated candidates than cand_cap: the sticky device flag is raised (the reference has no cap)."""
    from dsl_b200.postprocess import TeacherPost
    B, H, W = 1, 128, 160
    cls, box, ctr = GI.make_head_outputs(44, B, H, W, train=False, cls_mean=0.0)   # every score passes the gate
    sizes = GI.level_sizes(H, W)
    post = TeacherPost(B, sizes, GI.STRIDES, 80, DEV, cand_cap=1024)
    post.set_meta([(H, W, 3)], [[1.0] * 4])
    cls_out = [c.permute(0, 2, 3, 1).contiguous().to(DEV) for c in cls]
    rc_out = []
    for l, (h, w) in enumerate(sizes):
        rc = torch.zeros(B, h, w, 8, device=DEV)
        rc[..., :4] = box[l].permute(0, 2, 3, 1).to(DEV)
        rc_out.append(rc)
    post.decode(cls_out, rc_out)
    torch.cuda.synchronize()
    assert int(post.cand_counts[0]) > 1024
    assert post.overflowed() and not post.overflowed(), "flag is sticky until read, then cleared"


# ---- multi-tensor operand refresh: one table launch == per-tensor torch restatement, bit for bit ----------------------
@pytest.mark.gpu
def test_pack_table_tiled_tiles_and_merged_operands_bit_exact():
    """dslb_pack_plan_*: the tiled kernel picks its tile width from the filter (256 input channels for a 1x1 down to 32
    for a 3x3) and writes the fprop AND the dgrad operand of a conv from one read when both descriptors are in the
    table. Every operand must equal the bf16 rounding of w * bn_scale in its layout exactly."""
    from dsl_b200.engine import TablePlan
    g = torch.Generator(device="cpu").manual_seed(5)
    dev = "cuda"
    shapes = [(64, 64, 1, True, True), (256, 64, 1, True, False), (128, 512, 1, True, True), (512, 2048, 1, False, True),
              (64, 96, 1, True, True), (128, 128, 3, True, True), (256, 256, 3, False, True), (64, 64, 3, True, False),
              (32, 160, 1, False, True), (96, 32, 3, False, True)]
    descs, checks = [], []
    for O, I, k, bn, dgrad in shapes:
        w = torch.randn(O, I, k, k, generator=g).to(dev)
        d = dict(w=w, O=O, I=I, R=k, S=k, fill_padding=1)
        sc = torch.ones(O, device=dev)
        extra = {}
        if bn:
            gam, bet = torch.rand(O, generator=g).to(dev) + 0.5, torch.randn(O, generator=g).to(dev)
            mu, var = torch.randn(O, generator=g).to(dev), torch.rand(O, generator=g).to(dev) + 0.1
            sc = gam / torch.sqrt(var + 1e-5)
            extra = dict(bn_gamma=gam, bn_beta=bet, bn_mean=mu, bn_var=var, bn_eps=1e-5,
                         scale_out=torch.zeros(O, device=dev), shift_out=torch.zeros(O, device=dev))
            checks.append(("scale", extra["scale_out"], sc))
            checks.append(("shift", extra["shift_out"], bet - mu * sc))
        ws = (w * sc.view(-1, 1, 1, 1))
        wp = torch.full((k * k, O, I), 7.0, dtype=torch.bfloat16, device=dev)
        descs.append(dict(d, out=wp, rows_pad=O, cols_pad=I, mode=0, **extra))
        checks.append((f"fprop {O}x{I}x{k}", wp, ws.permute(2, 3, 0, 1).reshape(k * k, O, I).to(torch.bfloat16)))
        if dgrad:
            wpT = torch.full((k * k, I, O), 7.0, dtype=torch.bfloat16, device=dev)
            dd = dict(d, out=wpT, rows_pad=I, cols_pad=O, mode=1, **extra)
            # the dgrad descriptor of every other conv sits far from its fprop one in the table
            (descs.append if len(descs) % 4 else (lambda x: descs.insert(0, x)))(dd)
            checks.append((f"dgrad {O}x{I}x{k}", wpT,
                           ws.flip(2, 3).permute(2, 3, 1, 0).reshape(k * k, I, O).to(torch.bfloat16)))
    plan = TablePlan(descs, "pack", "test table")
    plan.run()
    torch.cuda.synchronize()
    for name, got, want in checks:
        if got.dtype == torch.bfloat16:
            assert torch.equal(got.view(torch.int16), want.contiguous().view(torch.int16)), name
        else:
            assert torch.allclose(got, want, rtol=1e-6, atol=1e-7), name
    # dgrad-only descriptor (no fprop partner in the table)
    w = torch.randn(128, 256, 3, 3, generator=g).to(dev)
    wpT = torch.zeros(9, 256, 128, dtype=torch.bfloat16, device=dev)
    TablePlan([dict(w=w, out=wpT, O=128, I=256, R=3, S=3, rows_pad=256, cols_pad=128, mode=1, fill_padding=1)], "pack",
              "dgrad only").run()
    torch.cuda.synchronize()
    assert torch.equal(wpT.view(torch.int16),
                       w.flip(2, 3).permute(2, 3, 1, 0).reshape(9, 256, 128).to(torch.bfloat16).contiguous().view(torch.int16))


# ---- GroupNorm backward with the group sums taken in the producing dgrad's epilogue ------------------------------------
def _gn_bwd_case(g, N, H, W, C=256):
    """Synthetic tower layer: pre-norm map x (bf16), its GroupNorm(32) statistics / (mean, rstd), gamma, beta."""
    from dsl_b200 import _lib as L
    x = _bf(torch.randn(N, H, W, C, generator=g) * 1.5 + 0.3).to(DEV)
    xg = x.double().view(N, H * W, 32, 8)
    stats = torch.zeros(N, 32, L.GN_STAT_STRIDE, dtype=torch.float64, device=DEV)
    stats[:, :, 0] = xg.sum(dim=(1, 3))
    stats[:, :, 1] = (xg * xg).sum(dim=(1, 3))
    m = float(H * W * 8)
    mean = stats[:, :, 0] / m
    rstd = 1.0 / torch.sqrt((stats[:, :, 1] / m - mean * mean).clamp_min(0) + 1e-5)
    mr = torch.zeros(N, 32, 4, dtype=torch.float32, device=DEV)
    mr[:, :, 0], mr[:, :, 1] = mean.float(), rstd.float()
    gamma = (torch.rand(C, generator=g) + 0.5).to(DEV)
    beta = (torch.randn(C, generator=g) * 0.2).to(DEV)
    return x.to(torch.bfloat16), stats, mr, gamma, beta


@pytest.mark.parametrize("pairs", ["1", "0"], ids=["cta-pairs", "single-cta"])
def test_gn_bwd_group_sums_in_dgrad_epilogue(pairs, monkeypatch):
    """dslb_conv_seg_t::gnb_*: the dgrad whose output is dz leaves sum(gamma*dy) and sum(gamma*dy*xhat) per (image, group)
    in gnb_sums, dy = bf16(dz) * [xhat*gamma + beta > 0]. Against a torch restatement on the stored (bf16) dz; maps that are
    smaller than a tile, tiles that straddle two images, three segments in one launch, both kernel flavours."""
    from dsl_b200 import _lib as L
    from dsl_b200.engine import ConvPlan
    monkeypatch.setenv("DSLB_CTA2", pairs)
    g = torch.Generator().manual_seed(21)
    C = 256
    w = _bf(torch.randn(C, C, 3, 3, generator=g) * 0.02)
    wpT = _pack(w, True)
    segs, keep = [], []
    for (N, H, W) in ((2, 25, 42), (2, 7, 11), (3, 13, 21)):
        dy_in = _bf(torch.randn(N, C, H, W, generator=g))
        x, stats, mr, gamma, beta = _gn_bwd_case(g, N, H, W)
        dz = torch.zeros(N, H, W, C, dtype=torch.bfloat16, device=DEV)
        sums = torch.zeros(N, 32, L.GN_STAT_STRIDE, dtype=torch.float64, device=DEV)
        segs.append(dict(x=_nhwc(dy_in), w=wpT, y=dz, N=N, H=H, W=W, Cin=C, Cout=C, cout_pad=C, R=3, S=3, stride=1, pad=1,
                         ldc=C, gnb_x=x, gnb_mr=mr, gnb_gamma=gamma, gnb_beta=beta, gnb_sums=sums, gn_cpg=8))
        keep.append((dy_in, x, mr, gamma, beta, dz, sums))
    ConvPlan(segs, "tower dgrad").run()
    torch.cuda.synchronize()
    for dy_in, x, mr, gamma, beta, dz, sums in keep:
        N = x.shape[0]
        ref = F.conv_transpose2d(dy_in.to(DEV), w.to(DEV), padding=1)   # == the dgrad the packed transposed operand computes
        assert _rel(dz.permute(0, 3, 1, 2).float(), ref) < 4e-3
        xg = x.float().view(N, -1, 32, 8)
        xh = (xg - mr[:, :, 0].view(N, 1, 32, 1)) * mr[:, :, 1].view(N, 1, 32, 1)
        ga, be = gamma.view(1, 1, 32, 8), beta.view(1, 1, 32, 8)
        dzg = dz.float().view(N, -1, 32, 8)
        gdy = torch.where(torch.addcmul(be, xh, ga) > 0, dzg, torch.zeros_like(dzg)) * ga
        s1, s2 = gdy.double().sum(dim=(1, 3)), (gdy.double() * xh.double()).sum(dim=(1, 3))
        # the sums cancel heavily (zero-mean dz): measure against the sum of magnitudes
        e1 = ((sums[:, :, 0] - s1).abs().max() / gdy.double().abs().sum(dim=(1, 3)).max()).item()
        e2 = ((sums[:, :, 1] - s2).abs().max() / (gdy.double() * xh.double()).abs().sum(dim=(1, 3)).max()).item()
        print(f"gnb sums {tuple(x.shape)}: S1 {e1:.2e} S2 {e2:.2e}")
        # 2e-4 of the magnitude sum = about one element of the smallest map: leaves room for a ReLU gate that flips on a
        # pre-activation within one fp32 ulp of zero, nothing systematic passes
        assert e1 < 2e-4 and e2 < 2e-4
        assert float(sums[:, :, 2:].abs().max()) == 0.0


def test_gn_bwd_one_pass_equals_reduce_plus_apply():
    """dslb_gn_bwd with gsums (one apply launch) against the same call without them (reduce + finalize + apply): same dx
    up to the rounding of the group constants, same per-channel sums for dgamma / dbeta and the same conv-bias gradient."""
    from dsl_b200 import _lib as L
    g = torch.Generator().manual_seed(22)
    C = 256
    cases = []
    for (N, H, W) in ((2, 25, 42), (1, 7, 11), (3, 13, 21)):
        x, stats, mr, gamma, beta = _gn_bwd_case(g, N, H, W)
        dz = _bf(torch.randn(N, H * W, C, generator=g)).to(DEV).to(torch.bfloat16)
        xg = x.float().view(N, -1, 32, 8)
        xh = (xg - mr[:, :, 0].view(N, 1, 32, 1)) * mr[:, :, 1].view(N, 1, 32, 1)
        ga, be = gamma.view(1, 1, 32, 8), beta.view(1, 1, 32, 8)
        dzg = dz.float().view(N, -1, 32, 8)
        gdy = torch.where(torch.addcmul(be, xh, ga) > 0, dzg, torch.zeros_like(dzg)) * ga
        gs = torch.zeros(N, 32, L.GN_STAT_STRIDE, dtype=torch.float64, device=DEV)
        gs[:, :, 0] = gdy.double().sum(dim=(1, 3))
        gs[:, :, 1] = (gdy.double() * xh.double()).sum(dim=(1, 3))
        cases.append((N, H * W, x, stats, mr, gamma, beta, dz, gs))

    def run(fused):
        arr = (L.GnSeg * len(cases))()
        outs = []
        for a, (N, HW, x, stats, mr, gamma, beta, dz, gs) in zip(arr, cases):
            dx = torch.zeros(N, HW, C, dtype=torch.bfloat16, device=DEV)
            red = torch.zeros(N, C, 2, dtype=torch.float64, device=DEV)
            dbias = torch.zeros(C, device=DEV)
            mr2 = mr.clone()
            for k, v in dict(x=x, y=dx, dz=dz, stats=stats, gamma=gamma, beta=beta, red=red, dbias=dbias, mr=mr2).items():
                setattr(a, k, v.data_ptr())
            a.N, a.HW = N, HW
            if fused:
                a.gsums = gs.data_ptr()
            outs.append((dx, red, dbias, mr2))
        nb = L.lib.dslb_gn_bwd_blocks(arr, len(cases))
        host = (C_int * (2 * nb))()
        L.check(L.lib.dslb_gn_bwd_plan(arr, len(cases), host), "plan")
        tab = torch.tensor(list(host), dtype=torch.int32, device=DEV)
        L.check(L.lib.dslb_gn_bwd(arr, len(cases), C, 32, 1e-5, L.ptr(tab), nb, L.cur_stream()), "gn_bwd")
        torch.cuda.synchronize()
        return outs

    import ctypes
    C_int = ctypes.c_int
    for (dx0, red0, db0, _), (dx1, red1, db1, _) in zip(run(False), run(True)):
        assert _rel(dx1.float(), dx0.float()) < 1e-2          # one bf16 ulp of the largest entry
        assert (dx1.float() - dx0.float()).abs().mean().item() < 1e-4 * dx0.float().abs().mean().item() + 1e-7
        scale = red0.abs().max().item()
        assert (red1 - red0).abs().max().item() < 2e-5 * scale + 1e-9
        assert (db1 - db0).abs().max().item() < 2e-3 * db0.abs().max().item() + 1e-4


@pytest.mark.parametrize("shape", [(2, 64, 50, 84), (1, 64, 37, 53), (3, 128, 8, 6), (1, 64, 5, 1), (2, 64, 400, 672)])
def test_maxpool_bit_exact(shape):
    """dslb_maxpool3x3s2 (two output pixels per thread, odd widths end with a single one) == nn.MaxPool2d(3, 2, 1) on the
    same bf16 values, bit for bit (max is exact)."""
    from dsl_b200 import _lib as L
    N, C, H, W = shape
    g = torch.Generator().manual_seed(H * 1000 + W)
    x = _bf(torch.randn(N, C, H, W, generator=g))
    xd = _nhwc(x)
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    y = torch.full((N, Ho, Wo, C), 7.0, dtype=torch.bfloat16, device=DEV)
    L.check(L.lib.dslb_maxpool3x3s2(L.ptr(xd), L.ptr(y), N, H, W, C, L.cur_stream()), "maxpool")
    ref = F.max_pool2d(x.to(DEV), 3, 2, 1)
    assert torch.equal(y.permute(0, 3, 1, 2).float(), ref)


def test_sgd_with_ema_equals_sgd_then_ema():
    """dslb_sgd_ema_step == dslb_sgd_step followed by dslb_ema_update on the same range, bit for bit (weights, momentum
    buffer and teacher), including a length that is not a multiple of four."""
    from dsl_b200 import _lib as L
    n = 1_000_003
    g = torch.Generator().manual_seed(3)
    p, gr, buf, t = (torch.randn(n, generator=g).to(DEV) for _ in range(4))
    coef = torch.tensor([0.37, 0.0], device=DEV)
    lrs = torch.tensor([0.5], device=DEV)
    k = 0.9996
    c_s, c_t = float(torch.tensor(1 - k, dtype=torch.float32)), float(torch.tensor(k, dtype=torch.float32))
    p1, b1, t1 = p.clone(), buf.clone(), t.clone()
    L.check(L.lib.dslb_sgd_step(L.ptr(p1), L.ptr(gr), L.ptr(b1), n, L.ptr(coef), L.ptr(lrs), 0.01, 0.9, 1e-4, 0,
                                L.cur_stream()), "sgd")
    L.check(L.lib.dslb_ema_update(L.ptr(t1), L.ptr(p1), n, c_s, c_t, L.cur_stream()), "ema")
    p2, b2, t2 = p.clone(), buf.clone(), t.clone()
    L.check(L.lib.dslb_sgd_ema_step(L.ptr(p2), L.ptr(gr), L.ptr(b2), n, L.ptr(coef), L.ptr(lrs), 0.01, 0.9, 1e-4, 0,
                                    L.ptr(t2), c_s, c_t, L.cur_stream()), "sgd+ema")
    torch.cuda.synchronize()
    assert torch.equal(p1, p2) and torch.equal(b1, b2) and torch.equal(t1, t2)
