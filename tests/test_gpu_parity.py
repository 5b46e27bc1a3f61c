"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle restatement on the same seeded
inputs, and against the committed golden fixtures produced by the reference's own source.

Bars (BASELINE.json north_star): pseudo-label index tensors bit-exact; head outputs within 1e-2 (bf16 compute) of the
fp32 reference; losses within 1e-3 relative.
"""
import os

import numpy as np
import pytest
import torch

from tests.golden import inputs as GI

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


@pytest.fixture(scope="module")
def small_net():
    from dsl_b200.engine import FCOSNet
    torch.manual_seed(0)
    net = FCOSNet(2, 256, 320, depth=50, train=True, seed=3, loss_weight=3.0, parity_outputs=True)
    return net


def _oracle_state(net, requires_grad=False):
    sd = {}
    for name, v in net.store.state_dict().items():
        t = v.detach().cpu().clone()
        sd[name] = t
    bb = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    neck = {k[len("neck."):]: v for k, v in sd.items() if k.startswith("neck.")}
    head = {k[len("bbox_head."):]: v for k, v in sd.items() if k.startswith("bbox_head.")}
    if requires_grad:
        for d in (bb, neck, head):
            for k, v in d.items():
                if v.dtype.is_floating_point and "running" not in k:
                    v.requires_grad_(True)
    return bb, neck, head


def _nchw(t, C):
    return t[..., :C].permute(0, 3, 1, 2).contiguous().float().cpu()


def test_forward_matches_oracle(small_net):
    """Backbone + FPN + head forward (train mode) vs the fp32 oracle on the same weights and image."""
    from oracle import fcos_oracle as O
    net = small_net
    rng = np.random.RandomState(5)
    img = GI.make_tensor(rng, 2, 3, 256, 320, scale=50.0)
    net.img.copy_(img)
    net.forward()
    torch.cuda.synchronize()
    bb, neck, head = _oracle_state(net)
    with torch.no_grad():
        cs = O.resnet_forward(bb, img, 50)
        ps = O.fpn_forward(neck, cs)
        cls, box, ctr = O.fcos_head_forward(head, ps, training=True)
    for i, (x, h, w, c) in enumerate(net.stage_out):
        e = _rel(_nchw(x, c), cs[i])
        print(f"C{i + 2} rel err {e:.3e}")
        assert e < 3e-2
    for l in range(5):
        e = _rel(_nchw(net.p[l], 256), ps[l])
        print(f"P{l + 3} rel err {e:.3e}")
        assert e < 3e-2
    for l in range(5):
        e1 = _rel(_nchw(net.cls_out[l], 80), cls[l])
        e2 = _rel(_nchw(net.rc_out[l], 4), box[l])
        e3 = (_nchw(net.rc_out[l][..., 4:5], 1) - ctr[l]).abs().max().item()
        print(f"level {l}: cls rel {e1:.3e} bbox rel {e2:.3e} ctr abs {e3:.3e}")
        # 60+ stacked bf16 convs: a looser bar than the head-only test below
        assert e1 < 2e-2 and e2 < 6e-2 and e3 < 6e-2


def test_head_forward_identical_inputs(small_net):
    """FCOSHead.forward on IDENTICAL (bf16-representable) input features, train and eval mode: cls logits within
    1e-2 of the fp32 reference (north_star bar for bf16 compute)."""
    from dsl_b200.engine import FCOSNet
    from oracle import fcos_oracle as O
    rng = np.random.RandomState(9)
    for net in (small_net, FCOSNet(2, 256, 320, depth=50, train=False, store=small_net.store)):
        feats = []
        for l, (h, w) in enumerate(net.psize):
            f = GI.make_tensor(rng, 2, 256, h, w).to(torch.bfloat16)
            net.p[l].copy_(f.permute(0, 2, 3, 1))
            feats.append(f.float())
        net.forward_head()
        torch.cuda.synchronize()
        _, _, head = _oracle_state(net)
        with torch.no_grad():
            cls, box, ctr = O.fcos_head_forward(head, feats, training=net.train)
        for l in range(5):
            e1 = _rel(_nchw(net.cls_out[l], 80), cls[l])
            e2 = _rel(_nchw(net.rc_out[l], 4), box[l])
            e3 = (_nchw(net.rc_out[l][..., 4:5], 1) - ctr[l]).abs().max().item() / (ctr[l].abs().max().item() + 1e-12)
            print(f"train={net.train} level {l}: cls rel {e1:.3e} bbox rel {e2:.3e} ctr rel {e3:.3e}")
            assert e1 < 1e-2 and e2 < 1e-2 and e3 < 1e-2


def _run_loss(net, gts, labels, ignores):
    net.set_targets([g.cuda() for g in gts], [l.cuda() for l in labels],
                    None if ignores is None else [i.cuda() for i in ignores])
    net.run_targets()
    net.run_loss()
    torch.cuda.synchronize()


def test_loss_and_targets_match_oracle_on_cuda_outputs(small_net):
    """Targets bit-exact and losses / input gradients within 1e-3 of the oracle evaluated on the SAME head outputs."""
    from oracle import fcos_oracle as O
    net = small_net
    B, H, W = 2, 256, 320
    gts, labels, ignores = GI.make_gt(77, B, H, W, with_ignore=True)
    _run_loss(net, gts, labels, ignores)
    cls = [_nchw(net.cls_out[l], 80).requires_grad_(True) for l in range(5)]
    box = [_nchw(net.rc_out[l], 4).requires_grad_(True) for l in range(5)]
    ctr = [_nchw(net.rc_out[l][..., 4:5], 1).requires_grad_(True) for l in range(5)]
    out = O.fcos_loss(cls, box, ctr, gts, labels, ignores, loss_weight=3.0, return_aux=True)
    aux = out.pop("_aux")
    assert torch.equal(net.labels.cpu(), aux["labels"]), "labels must be bit-exact"
    assert torch.equal(net.bbox_targets.cpu(), aux["bbox_targets"]), "bbox_targets must be bit-exact"
    assert torch.equal(net.weights.cpu(), aux["weight"])
    got = net.losses()
    for k in ("loss_cls", "loss_bbox", "loss_centerness"):
        r = abs(got[k].item() - out[k].item()) / (abs(out[k].item()) + 1e-12)
        print(k, got[k].item(), out[k].item(), f"rel {r:.2e}")
        assert r < 1e-3
    sum(out.values()).backward()
    for l in range(5):
        gc = cls[l].grad.permute(0, 2, 3, 1)
        e = _rel(net.dcls_f32[l].cpu(), gc)
        gb = box[l].grad.permute(0, 2, 3, 1)
        eb = (net.drc_f32[l][..., :4].cpu() - gb).abs().max().item() / (gb.abs().max().item() + 1e-12)
        gt_ = ctr[l].grad.permute(0, 2, 3, 1)
        et = (net.drc_f32[l][..., 4:5].cpu() - gt_).abs().max().item() / (gt_.abs().max().item() + 1e-12)
        print(f"level {l}: dcls rel {e:.2e} dbox rel {eb:.2e} dctr rel {et:.2e}")
        assert e < 1e-3 and eb < 1e-3 and et < 1e-3


LOSS_CASES = {
    "base_b2": (21, 2, 256, 320, dict(), dict(with_ignore=False)),
    "dsl_b2": (22, 2, 256, 320, dict(loss_weight=3.0), dict(with_ignore=True)),
    "dsl_b3_si": (23, 3, 256, 320, dict(loss_weight=3.0, soft_weight=1.0, soft_warm_up=5000), dict(with_ignore=True)),
    "empty_gt": (24, 2, 256, 320, dict(loss_weight=3.0), dict(with_ignore=True, empty_first=True)),
    "tie_break": (25, 2, 256, 320, dict(), dict(with_ignore=False, duplicate_boxes=True)),
    "many_gt": (26, 2, 384, 512, dict(loss_weight=3.0), dict(with_ignore=True, max_gt=40, max_ignore=8)),
    "ragged_hw": (27, 4, 200, 264, dict(loss_weight=3.0), dict(with_ignore=True)),
}


@pytest.mark.parametrize("name", sorted(LOSS_CASES))
def test_loss_kernels_match_reference_golden(name):
    """The two loss kernels driven directly through the C ABI on the golden inputs: labels / bbox_targets bit-exact
    with the reference's own get_targets, losses and gradients within 1e-3 of the reference's values."""
    import ctypes as C
    from dsl_b200 import _lib as L
    seed, B, H, W, hk, gk = LOSS_CASES[name]
    g = np.load(os.path.join(G, f"loss_{name}.npz"))
    cls, box, ctr = GI.make_head_outputs(seed, B, H, W, train=True)
    gts, labels, ignores = GI.make_gt(seed + 1000, B, H, W, **gk)
    sizes = GI.level_sizes(H, W)
    dev = "cuda"
    nl = 5
    lv = (L.FcosLevel * nl)()
    keep = []
    P = B * sum(h * w for h, w in sizes)
    for l, (h, w) in enumerate(sizes):
        c = cls[l].permute(0, 2, 3, 1).contiguous().to(dev)
        rc = torch.zeros(B, h, w, 8, device=dev)
        rc[..., :4] = box[l].permute(0, 2, 3, 1).to(dev)
        rc[..., 4] = ctr[l][:, 0].to(dev)
        dc = torch.zeros(B, h, w, 80, device=dev)
        dr = torch.zeros(B, h, w, 8, device=dev)
        keep += [c, rc, dc, dr]
        a = lv[l]
        a.cls, a.regctr, a.dcls_f32, a.dregctr_f32 = c.data_ptr(), rc.data_ptr(), dc.data_ptr(), dr.data_ptr()
        a.h, a.w, a.stride = h, w, GI.STRIDES[l]
        a.ld_cls, a.ld_dcls, a.ld_dreg = 80, 0, 0
        a.rr_lo, a.rr_hi = float(GI.REGRESS_RANGES[l][0]), float(GI.REGRESS_RANGES[l][1])
        a.scale, a.cs_radius = 1.0, float(GI.STRIDES[l] * 1.5)
    offs = np.cumsum([0] + [len(b) for b in gts]).astype(np.int32)
    gt_b = torch.cat(gts).to(dev) if offs[-1] else torch.zeros(1, 4, device=dev)
    gt_l = torch.cat(labels).to(dev) if offs[-1] else torch.zeros(1, dtype=torch.int64, device=dev)
    gt_o = torch.from_numpy(offs).to(dev)
    if ignores is not None:
        ioffs = np.cumsum([0] + [len(b) for b in ignores]).astype(np.int32)
        ig_b = torch.cat(ignores).to(dev) if ioffs[-1] else torch.zeros(1, 4, device=dev)
        ig_o = torch.from_numpy(ioffs).to(dev)
    lab = torch.zeros(P, dtype=torch.int64, device=dev)
    tgt = torch.zeros(P, 4, device=dev)
    wts = torch.zeros(P, device=dev)
    ctt = torch.zeros(P, device=dev)
    counts = torch.zeros(2, dtype=torch.float64, device=dev)
    norm = torch.zeros(2, device=dev)
    sums = torch.zeros(4, dtype=torch.float64, device=dev)
    lw = hk.get("loss_weight", 1.0)
    n_lab = B // 2 if B % 2 == 0 else (B - 1) // 2
    s = L.cur_stream()
    L.check(L.lib.dslb_fcos_targets(lv, nl, B, 80, L.ptr(gt_b), L.ptr(gt_l), L.ptr(gt_o),
                                    L.ptr(ig_b) if ignores is not None else None,
                                    L.ptr(ig_o) if ignores is not None else None, 1, 1, lw, n_lab, L.ptr(lab),
                                    L.ptr(tgt), L.ptr(wts), L.ptr(ctt), L.ptr(counts), s), "targets")
    L.check(L.lib.dslb_fcos_norm(L.ptr(counts), 1.0, L.ptr(norm), s), "norm")
    si = 0.0
    if B % 2 == 1 and hk.get("soft_weight", 0.0) != 0.0:
        si = hk["soft_weight"] / 1000.0  # warm-up branch (soft_warm_up >= cur_iter), fcos_head.py:325-327
    L.check(L.lib.dslb_fcos_loss(lv, nl, B, 80, L.ptr(lab), L.ptr(tgt), L.ptr(wts), L.ptr(ctt), L.ptr(norm), 0.25, 2.0,
                                 lw, n_lab, si, None, L.ptr(sums), None, s), "loss")
    torch.cuda.synchronize()
    assert np.array_equal(lab.cpu().numpy().astype(np.int16), g["labels"]), "labels vs reference get_targets"
    assert np.array_equal(tgt.cpu().numpy(), g["bbox_targets"]), "bbox_targets vs reference get_targets"
    names = ["loss_cls", "loss_bbox", "loss_centerness", "loss_sisoft"]
    for i, k in enumerate(names):
        if k in g.files:
            r = abs(sums[i].item() - float(g[k])) / (abs(float(g[k])) + 1e-12)
            print(name, k, sums[i].item(), float(g[k]), f"rel {r:.2e}")
            assert r < 1e-3
    dcls = torch.cat([keep[4 * l + 2].reshape(-1, 80) for l in range(nl)]).cpu()
    dbox = torch.cat([keep[4 * l + 3].reshape(-1, 8)[:, :4] for l in range(nl)]).cpu().numpy()
    dctr = torch.cat([keep[4 * l + 3].reshape(-1, 8)[:, 4] for l in range(nl)]).cpu().numpy()
    np.testing.assert_allclose(dcls.reshape(-1)[::17].numpy(), g["dcls_sample"], rtol=2e-3, atol=1e-9)
    np.testing.assert_allclose(dcls.abs().double().sum().item(), float(g["dcls_abs_sum"]), rtol=1e-4)
    np.testing.assert_allclose(dbox, g["dbox"], rtol=2e-3, atol=1e-8)
    np.testing.assert_allclose(dctr, g["dctr"], rtol=2e-3, atol=1e-9)
    _ = C


def test_backward_matches_oracle_autograd(small_net):
    """Parameter gradients of the full student step vs torch autograd through the fp32 oracle (bf16 compute =>
    compared per tensor by relative L2 error)."""
    from oracle import fcos_oracle as O
    net = small_net
    B, H, W = 2, 256, 320
    rng = np.random.RandomState(5)
    img = GI.make_tensor(rng, B, 3, H, W, scale=50.0)
    gts, labels, ignores = GI.make_gt(77, B, H, W, with_ignore=True)
    net.img.copy_(img)
    net.forward()
    _run_loss(net, gts, labels, ignores)
    net.backward()
    torch.cuda.synchronize()
    bb, neck, head = _oracle_state(net, requires_grad=True)
    cs = O.resnet_forward(bb, img, 50)
    ps = O.fpn_forward(neck, cs)
    cls, box, ctr = O.fcos_head_forward(head, ps, training=True)
    out = O.fcos_loss(cls, box, ctr, gts, labels, ignores, loss_weight=3.0)
    sum(out.values()).backward()
    worst = 0.0
    bad = []
    for prefix, d in (("backbone.", bb), ("neck.", neck), ("bbox_head.", head)):
        for k, v in d.items():
            name = prefix + k
            o, n = net.store.offsets.get(name, (None, None))
            if o is None or o + n > net.store.n_train:
                continue
            ref = v.grad if v.grad is not None else torch.zeros_like(v)
            got = net.grad[o:o + n].view(ref.shape).cpu()
            den = ref.norm().item()
            err = (got - ref).norm().item() / (den + 1e-12) if den > 0 else got.norm().item()
            worst = max(worst, err)
            if err > 5e-2:
                bad.append((name, err, den))
    print("worst relative L2 gradient error", worst)
    for b in bad[:20]:
        print("BAD", b)
    assert not bad, f"{len(bad)} parameter gradients off by more than 5e-2 relative L2"


def test_ema_and_sgd_kernels():
    import ctypes as C
    from dsl_b200 import _lib as L
    from oracle import fcos_oracle as O
    n = 1_000_003
    g = torch.Generator(device="cpu").manual_seed(1)
    s, t = torch.randn(n, generator=g), torch.randn(n, generator=g)
    sd, td = s.cuda(), t.cuda()
    k = 0.99
    c_s, c_t = float(torch.tensor(1 - k, dtype=torch.float32)), float(torch.tensor(k, dtype=torch.float32))
    L.check(L.lib.dslb_ema_update(L.ptr(td), L.ptr(sd), n, c_s, c_t, L.cur_stream()), "ema")
    ref = O.ema_update({"w": t}, {"w": s}, k)["w"]
    assert torch.equal(td.cpu(), ref), "EMA must be bit-exact with the reference's fp32 expression"
    p, gr, buf = torch.randn(n, generator=g), torch.randn(n, generator=g), torch.randn(n, generator=g)
    pd, gd, bd = p.cuda(), gr.cuda(), buf.cuda()
    sq = torch.zeros(1, dtype=torch.float64, device="cuda")
    coef = torch.zeros(2, device="cuda")
    L.check(L.lib.dslb_sq_norm(L.ptr(gd), n, L.ptr(sq), L.cur_stream()), "sqnorm")
    L.check(L.lib.dslb_clip_coef(L.ptr(sq), 35.0, L.ptr(coef), L.cur_stream()), "coef")
    L.check(L.lib.dslb_sgd_step(L.ptr(pd), L.ptr(gd), L.ptr(bd), n, L.ptr(coef), None, 0.01, 0.9, 1e-4, 0,
                                L.cur_stream()), "sgd")
    (gc,), total = O.clip_grad_norm([gr], 35.0)
    p_ref, b_ref = O.sgd_momentum_step(p, gc, buf, 0.01, 0.9, 1e-4)
    assert abs(coef[1].item() - total.item()) / total.item() < 1e-5
    assert _rel(pd.cpu(), p_ref) < 1e-6 and _rel(bd.cpu(), b_ref) < 1e-6
    _ = C
