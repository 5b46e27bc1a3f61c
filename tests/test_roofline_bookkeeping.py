"""Host-side roofline bookkeeping (no GPU): algorithmic HBM bytes of a conv segment, the tensor-bound / HBM-bound split
bench.py reports under roofline.by_bound, and that every plan of a network built on the emulator carries its bytes."""
import torch

from dsl_b200 import engine as E
from dsl_b200.trainer import DSLEngine


def test_seg_bytes_matches_the_design_formula():
    # layer1 conv3 at B=4, 800x1344: 1x1 64 -> 256 with the identity as residual (2*npix*(Cin + 2*Cout))
    npix = 4 * 200 * 336
    s = dict(N=4, H=200, W=336, Cin=64, Cout=256, R=1, S=1, stride=1, pad=0, residual=torch.zeros(1))
    assert E.seg_bytes(s) == 2 * npix * (64 + 2 * 256) + 2 * 64 * 256
    # head tower layer at P3: 3x3 256 -> 256
    s = dict(N=4, H=100, W=168, Cin=256, Cout=256, R=3, S=3, stride=1, pad=1)
    assert E.seg_bytes(s) == 2 * 67200 * 512 + 2 * 9 * 256 * 256
    # stride-2 1x1 (downsample): reads only the pixels it keeps; fp32 output; mask + residual reads
    s = dict(N=1, H=8, W=8, Cin=64, Cout=128, R=1, S=1, stride=2, pad=0, out_fp32=1, relu_mask=torch.zeros(1),
             residual=torch.zeros(1))
    assert E.seg_bytes(s) == 2 * 16 * 64 + 2 * 64 * 128 + 16 * 128 * (4 + 4)
    # stride-2 3x3 (FPN P6): whole input
    s = dict(N=1, H=9, W=9, Cin=64, Cout=64, R=3, S=3, stride=2, pad=1)
    assert E.seg_bytes(s) == 2 * 81 * 64 + 2 * 9 * 64 * 64 + 25 * 64 * 2


def test_split_by_bound():
    rows = [(1.0, 79.3e9, 70e6),        # tower conv: 1133 FLOP/B -> tensor
            (0.5, 8.8e9, 310e6),        # layer1 conv3: 28 FLOP/B -> hbm
            (0.25, 8.8e9, 172e6),       # layer1 conv1 -> hbm
            (0.1, 1.0e9, 0.0)]          # no byte figure: stays with the tensor class
    out = DSLEngine.split_by_bound(rows, ridge=210.0)
    assert out["tensor"]["n"] == 2 and out["hbm"]["n"] == 2
    assert out["hbm"]["bytes"] == 482e6 and out["hbm"]["ms"] == 0.75 and out["tensor"]["flops"] == 80.3e9


def test_every_conv_plan_of_a_network_carries_its_bytes():
    from tests import emu_lib
    with emu_lib.installed():
        net = E.FCOSNet(1, 64, 96, 50, 80, train=True, device="cpu", seed=0, parts="backbone")
        plans = [getattr(op, "__self__", None) for op in net.fwd_ops + net.bwd_ops]
        plans = [p for p in plans if isinstance(p, E.ConvPlan)]
        assert len(plans) > 60
        for p in plans:
            assert p.bytes > 0, p.what
        c3 = [p for p in plans if p.what.endswith("layer1.0.conv3")][0]
        npix = 16 * 24
        assert c3.bytes == 2 * npix * (64 + 2 * 256) + 2 * 64 * 256
        del net, plans, c3
