"""LOSSES-registry modules (dsl_b200.losses) vs golden vectors produced by the reference's own FocalLoss / GIoULoss /
CrossEntropyLoss classes (oracle/gen_golden.py::gen_loss_modules, executed from /root/reference in the build
container): values and gradients for every reduction / avg_factor / weight combination."""
import os

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(__file__), "golden")
CASES = [("mean_w_avg", dict(weight=True, avg_factor=37.5, reduction_override=None)),
         ("mean_now", dict(weight=False, avg_factor=None, reduction_override=None)),
         ("sum_w", dict(weight=True, avg_factor=None, reduction_override="sum")),
         ("none_w", dict(weight=True, avg_factor=None, reduction_override="none"))]


def test_loss_modules_construct_like_the_reference():
    from dsl_b200 import losses
    f = losses.FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0)
    assert (f.gamma, f.alpha, f.reduction, f.loss_weight) == (2.0, 0.25, "mean", 1.0)
    g = losses.GIoULoss(loss_weight=1.0)
    assert g.eps == 1e-6
    c = losses.CrossEntropyLoss(use_sigmoid=True, loss_weight=1.0)
    assert c.use_sigmoid and c.reduction == "mean"
    with pytest.raises(AssertionError):
        losses.FocalLoss(use_sigmoid=False)            # the reference asserts the same
    with pytest.raises(NotImplementedError):
        losses.CrossEntropyLoss(use_sigmoid=False)     # softmax CE is not on the path: loud, never a fallback
    with pytest.raises(RuntimeError):
        f(torch.zeros(4, 3), torch.zeros(4, dtype=torch.long))   # CPU tensors: no fallback
    with pytest.raises(ValueError):
        losses._reduce_args("sum", 3.0, 10)            # weight_reduce_loss raises the same


@pytest.mark.gpu
@pytest.mark.parametrize("mname", ["focal", "giou", "bce", "focal_lw"])
def test_loss_modules_match_reference_golden(mname):
    from dsl_b200 import losses
    g = np.load(os.path.join(G, "loss_modules.npz"))
    t = lambda k: torch.from_numpy(g[k]).cuda()  # noqa: E731
    mod, pred, tgt, w = {
        "focal": (losses.FocalLoss(use_sigmoid=True, gamma=2.0, alpha=0.25, loss_weight=1.0), "logits", "labels", "wN"),
        "giou": (losses.GIoULoss(loss_weight=1.0), "b1", "b2", "wn"),
        "bce": (losses.CrossEntropyLoss(use_sigmoid=True, loss_weight=1.0), "ctr", "ctr_t", "wn"),
        "focal_lw": (losses.FocalLoss(use_sigmoid=True, gamma=1.5, alpha=0.4, loss_weight=2.5), "logits", "labels", "wN"),
    }[mname]
    for cname, kw in CASES:
        x = t(pred).clone().requires_grad_(True)
        val = mod(x, t(tgt), weight=t(w) if kw["weight"] else None, avg_factor=kw["avg_factor"],
                  reduction_override=kw["reduction_override"])
        ref = g[f"{mname}_{cname}_val"]
        assert tuple(val.shape) == ref.shape, (mname, cname)
        (val * t(f"{mname}_{cname}_gout")).sum().backward()
        # fp32 element math on both sides (expf/log1pf vs torch's CPU libm: a few ulp), fp64 vs fp32 summation
        assert np.allclose(val.detach().cpu().numpy(), ref, rtol=2e-5, atol=1e-6), (mname, cname)
        assert np.allclose(x.grad.cpu().numpy(), g[f"{mname}_{cname}_grad"], rtol=2e-4, atol=2e-6), (mname, cname)


@pytest.mark.gpu
def test_giou_loss_all_zero_weight_early_out():
    """iou_loss.py:345-348: no positive weight -> (pred * weight).sum(), i.e. 0 with a live graph."""
    from dsl_b200 import losses
    p = torch.rand(5, 4, device="cuda", requires_grad=True)
    out = losses.GIoULoss()(p, torch.rand(5, 4, device="cuda"), weight=torch.zeros(5, device="cuda"))
    out.backward()
    assert float(out.detach()) == 0.0 and torch.equal(p.grad, torch.zeros_like(p))
