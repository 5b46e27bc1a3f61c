"""CTA-pair (tcgen05 cta_group::2) variant of the implicit-GEMM conv kernel, opt-in with DSLB_CTA2=1: two CTAs of a cluster
compute two consecutive 128-pixel tiles as ONE M=256 MMA, each staging its own A tile and half of the weight tile.
Parity against torch fp32 on identical inputs and against the single-CTA kernel on the ten-segment tower layer."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


@pytest.fixture()
def cta2_env():
    old = os.environ.get("DSLB_CTA2")
    os.environ["DSLB_CTA2"] = "1"
    yield
    if old is None:
        os.environ.pop("DSLB_CTA2", None)
    else:
        os.environ["DSLB_CTA2"] = old


@pytest.mark.parametrize("shape", [(2, 25, 42), (1, 7, 11), (4, 100, 168), (3, 13, 21)])
def test_cta2_conv_matches_fp32_reference(cta2_env, shape):
    from tests.test_gpu_kernels_r2 import _bf, _nhwc, _pack
    from dsl_b200.engine import ConvPlan
    N, H, W = shape
    g = torch.Generator().manual_seed(N * 1000 + H)
    x = _bf(torch.randn(N, 256, H, W, generator=g))
    w = _bf(torch.randn(256, 256, 3, 3, generator=g) * 0.02)
    shift = torch.randn(256, generator=g) * 0.1
    ref = F.relu(F.conv2d(x.to(DEV), w.to(DEV), padding=1) + shift.to(DEV).view(1, -1, 1, 1))
    y = torch.full((N, H, W, 256), float("nan"), dtype=torch.bfloat16, device=DEV)
    ConvPlan([dict(x=_nhwc(x), w=_pack(w, False), y=y, N=N, H=H, W=W, Cin=256, Cout=256, cout_pad=256, R=3, S=3, stride=1,
                   pad=1, ldc=256, shift=shift.to(DEV), relu_nch=256)], "cta2").run()
    torch.cuda.synchronize()
    got = y.permute(0, 3, 1, 2).float()
    assert torch.isfinite(got).all()
    e = _rel(got, ref)
    print(f"cta2 {shape}: rel {e:.2e}")
    assert e < 4e-3


# (name, N, H, W, Cin, Cout, k, stride, pad, relu, mask)
CASES = [
    ("3x3 512->512, two n-tiles (layer4 conv2)", 2, 25, 42, 512, 512, 3, 1, 1, True, False),
    ("1x1 1024->256 K=16 (layer3 conv1)", 2, 50, 84, 1024, 256, 1, 1, 0, True, False),
    ("1x1 2048->256 (lateral)", 2, 25, 42, 2048, 256, 1, 1, 0, False, False),
    ("1x1 stride 2 1024->512 (layer4 conv1, caffe style)", 2, 50, 84, 1024, 512, 1, 2, 0, True, False),
    ("3x3 256->256 mask (layer3 conv2 dgrad)", 3, 50, 84, 256, 256, 3, 1, 1, False, True),
    ("3x3 128->256 (predictor dgrad shape, K=18)", 2, 25, 43, 128, 256, 3, 1, 1, False, False),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cta2_other_shapes(cta2_env, case):
    from tests.test_gpu_kernels_r2 import _bf, _nhwc, _pack
    from dsl_b200.engine import ConvPlan
    _, N, H, W, Ci, Co, k, stride, pad, relu, use_mask = case
    g = torch.Generator().manual_seed(Ci + Co + H)
    x = _bf(torch.randn(N, Ci, H, W, generator=g))
    w = _bf(torch.randn(Co, Ci, k, k, generator=g) * (1.0 / (Ci * k * k) ** 0.5))
    shift = torch.randn(Co, generator=g) * 0.1
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    mask = _bf(torch.randn(N, Co, Ho, Wo, generator=g)) if use_mask else None
    ref = F.conv2d(x.to(DEV), w.to(DEV), stride=stride, padding=pad) + shift.to(DEV).view(1, -1, 1, 1)
    if relu:
        ref = F.relu(ref)
    if use_mask:
        ref = ref * (mask.to(DEV) > 0)
    y = torch.full((N, Ho, Wo, Co), float("nan"), dtype=torch.bfloat16, device=DEV)
    seg = dict(x=_nhwc(x), w=_pack(w, False), y=y, N=N, H=H, W=W, Cin=Ci, Cout=Co, cout_pad=Co, R=k, S=k, stride=stride,
               pad=pad, ldc=Co, shift=shift.to(DEV), relu_nch=Co if relu else 0)
    if use_mask:
        seg["relu_mask"] = _nhwc(mask)
    ConvPlan([seg], "cta2").run()
    torch.cuda.synchronize()
    got = y.permute(0, 3, 1, 2).float()
    assert torch.isfinite(got).all()
    e = _rel(got, ref)
    print(f"cta2 {case[0]}: rel {e:.2e}")
    assert e < 4e-3


def test_cta2_tower_layer_matches_single_cta_kernel(cta2_env):
    """The ten-segment FCOSHead tower layer (GroupNorm statistics in the epilogue) under both kernels: same bf16 maps,
    same statistics (summation order inside a tile is identical; across tiles it is fp64 atomics)."""
    from tools.ncu_cases import build_tower
    torch.manual_seed(0)
    plans = build_tower(B=2, H=256, W=320)
    pair = plans[0]
    os.environ["DSLB_CTA2"] = "0"
    from dsl_b200.engine import ConvPlan
    segs = []
    outs = []
    for s in pair.segs:
        s2 = dict(s)
        s2["y"] = torch.zeros_like(s["y"])
        s2["gn_stats"] = torch.zeros_like(s["gn_stats"])
        outs.append((s["y"], s["gn_stats"], s2["y"], s2["gn_stats"]))
        segs.append(s2)
    single = ConvPlan(segs, "single")
    pair.run()
    single.run()
    torch.cuda.synchronize()
    for y2, st2, y1, st1 in outs:
        assert torch.equal(y2, y1), "the pair kernel must produce the same bf16 map as the single-CTA kernel"
        assert _rel(st2[..., :2], st1[..., :2]) < 1e-9
