"""LOSSES-registry modules of the hot path on libdslb.so: FocalLoss, GIoULoss, CrossEntropyLoss(use_sigmoid=True).

Same constructor kwargs, forward signatures, reduction / avg_factor / loss_weight semantics and exceptions as the
reference (mmdet/models/losses/focal_loss.py:105-181, iou_loss.py:329-366, cross_entropy_loss.py:166-250,
utils.py:27-54). The training step does NOT go through these (FCOSHead.loss is one fused kernel, dslb_fcos_loss); they
exist so that a config or a caller that builds `dict(type='FocalLoss', ...)` by itself still lands on CUDA kernels.
No CPU / eager fallback: CUDA tensors only.
"""
import torch
import torch.nn as nn

from . import _lib as L


def _need_cuda(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"dsl_b200.losses.{what} runs only on CUDA tensors (kernels in libdslb.so; no CPU fallback)")


class _ElementwiseLoss(torch.autograd.Function):
    """forward: one kernel -> (weighted element-wise loss | its sum) and the gradient of the weighted loss;
    backward: that gradient x upstream gradient (x the reduction scale)."""

    @staticmethod
    def forward(ctx, pred, launch, elem_shape, reduce, scale):
        # launch(loss_elem_or_None, loss_sum_or_None, dpred) runs the kernel on the current stream
        dpred = torch.empty_like(pred, dtype=torch.float32)
        if reduce:
            acc = torch.zeros(1, dtype=torch.float64, device=pred.device)
            launch(None, acc, dpred)
            out = (acc[0] * scale).to(torch.float32)
        else:
            out = torch.empty(elem_shape, dtype=torch.float32, device=pred.device)
            launch(out, None, dpred)
            if scale != 1.0:
                out = out * scale
        ctx.save_for_backward(dpred)
        ctx.reduce, ctx.scale, ctx.elem_shape = reduce, scale, elem_shape
        return out

    @staticmethod
    def backward(ctx, gout):
        (dpred,) = ctx.saved_tensors
        if ctx.reduce:
            g = dpred * (gout.to(torch.float32) * ctx.scale)
        else:
            g = gout.to(torch.float32) * ctx.scale
            while g.dim() < dpred.dim():   # (n,) loss of (n, 4) boxes
                g = g.unsqueeze(-1)
            g = dpred * g
        return g, None, None, None, None


def _reduce_args(reduction, avg_factor, numel):
    """weight_reduce_loss (losses/utils.py:27-54) as (reduce to a scalar?, scalar scale)."""
    if avg_factor is None:
        if reduction == "mean":
            return True, 1.0 / max(numel, 1) if numel else float("nan")
        if reduction == "sum":
            return True, 1.0
        return False, 1.0
    if reduction == "mean":
        return True, 1.0 / float(avg_factor)
    if reduction != "none":
        raise ValueError('avg_factor can not be used with reduction="sum"')
    return False, 1.0


class FocalLoss(nn.Module):
    """mmdet/models/losses/focal_loss.py:105-181 (sigmoid focal loss; `target` = class indices in [0, C], C = background)."""

    def __init__(self, use_sigmoid=True, gamma=2.0, alpha=0.25, reduction="mean", loss_weight=1.0):
        super().__init__()
        assert use_sigmoid is True, "Only sigmoid focal loss supported now."
        self.use_sigmoid, self.gamma, self.alpha = use_sigmoid, gamma, alpha
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        assert reduction_override in (None, "none", "mean", "sum")
        reduction = reduction_override if reduction_override else self.reduction
        _need_cuda(pred, "FocalLoss")
        N, C = pred.shape
        x = pred.contiguous().float()
        lab = target.contiguous().to(torch.int64)
        w = None
        if weight is not None:
            if weight.numel() != N:
                raise NotImplementedError("dsl_b200 FocalLoss: per-class weights (FSAF style) are not implemented")
            w = weight.reshape(-1).contiguous().float()
        reduce, scale = _reduce_args(reduction, avg_factor, N * C)

        def launch(elem, acc, dpred):
            L.check(L.lib.dslb_sigmoid_focal_loss(L.ptr(x), L.ptr(lab), L.ptr(w) if w is not None else None, N, C,
                                                  float(self.alpha), float(self.gamma),
                                                  L.ptr(elem) if elem is not None else None,
                                                  L.ptr(acc) if acc is not None else None, L.ptr(dpred),
                                                  L.cur_stream()), "sigmoid_focal_loss")

        return self.loss_weight * _ElementwiseLoss.apply(x, launch, (N, C), reduce, scale)


class GIoULoss(nn.Module):
    """mmdet/models/losses/iou_loss.py:329-366: 1 - GIoU of aligned (x1, y1, x2, y2) boxes."""

    def __init__(self, eps=1e-6, reduction="mean", loss_weight=1.0):
        super().__init__()
        self.eps, self.reduction, self.loss_weight = eps, reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        _need_cuda(pred, "GIoULoss")
        if weight is not None and not torch.any(weight > 0):
            if pred.dim() == weight.dim() + 1:
                weight = weight.unsqueeze(1)
            return (pred * weight).sum()  # 0, keeps the graph (iou_loss.py:345-348)
        assert reduction_override in (None, "none", "mean", "sum")
        reduction = reduction_override if reduction_override else self.reduction
        if weight is not None and weight.dim() > 1:
            assert weight.shape == pred.shape
            weight = weight.mean(-1)
        n = pred.shape[0]
        p = pred.contiguous().float()
        t = target.contiguous().float()
        w = weight.contiguous().float() if weight is not None else None
        reduce, scale = _reduce_args(reduction, avg_factor, n)

        def launch(elem, acc, dpred):
            L.check(L.lib.dslb_giou_loss(L.ptr(p), L.ptr(t), L.ptr(w) if w is not None else None, n, float(self.eps),
                                         L.ptr(elem) if elem is not None else None,
                                         L.ptr(acc) if acc is not None else None, L.ptr(dpred), L.cur_stream()),
                    "giou_loss")

        return self.loss_weight * _ElementwiseLoss.apply(p, launch, (n,), reduce, scale)


class CrossEntropyLoss(nn.Module):
    """mmdet/models/losses/cross_entropy_loss.py:166-250, the use_sigmoid=True branch (binary_cross_entropy on logits of
    the same shape as the float targets — FCOS centerness). Softmax / mask CE are not on the hot path."""

    def __init__(self, use_sigmoid=False, use_mask=False, reduction="mean", class_weight=None, ignore_index=None,
                 loss_weight=1.0):
        super().__init__()
        assert (use_sigmoid is False) or (use_mask is False)
        if not use_sigmoid or use_mask or class_weight is not None:
            raise NotImplementedError("dsl_b200 CrossEntropyLoss: only use_sigmoid=True without class_weight is "
                                      "implemented (the FCOS centerness loss)")
        self.use_sigmoid, self.use_mask = use_sigmoid, use_mask
        self.reduction, self.loss_weight, self.class_weight, self.ignore_index = reduction, loss_weight, None, ignore_index

    def forward(self, cls_score, label, weight=None, avg_factor=None, reduction_override=None, ignore_index=None,
                **kwargs):
        assert reduction_override in (None, "none", "mean", "sum")
        reduction = reduction_override if reduction_override else self.reduction
        _need_cuda(cls_score, "CrossEntropyLoss")
        if cls_score.dim() != label.dim():
            raise NotImplementedError("dsl_b200 CrossEntropyLoss: class-index labels (one-hot expansion) are not "
                                      "implemented; pass float targets of the prediction's shape")
        x = cls_score.contiguous().float()
        y = label.contiguous().float()
        w = weight.contiguous().float() if weight is not None else None
        n = x.numel()
        reduce, scale = _reduce_args(reduction, avg_factor, n)

        def launch(elem, acc, dpred):
            L.check(L.lib.dslb_bce_with_logits(L.ptr(x), L.ptr(y), L.ptr(w) if w is not None else None, n,
                                               L.ptr(elem) if elem is not None else None,
                                               L.ptr(acc) if acc is not None else None, L.ptr(dpred), L.cur_stream()),
                    "bce_with_logits")

        return self.loss_weight * _ElementwiseLoss.apply(x, launch, tuple(x.shape), reduce, scale)


def register(force=True):
    """Register under the reference's LOSSES keys when mmdet is importable; returns the keys registered."""
    try:
        from mmdet.models.builder import LOSSES
    except Exception:
        return []
    for cls in (FocalLoss, GIoULoss, CrossEntropyLoss):
        LOSSES.register_module(name=cls.__name__, force=force, module=cls)
    return ["LOSSES.FocalLoss", "LOSSES.GIoULoss", "LOSSES.CrossEntropyLoss"]
