"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain torch fp32 / numpy) of the reference's algorithm for DSL's
dense teacher-student hot path. It is the checker for the CUDA path: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it. The product (dsl_b200/) never does.

Pinned: every function here is compared, in tests/test_oracle_golden.py, against golden vectors produced by
executing the reference's own source in place (oracle/gen_golden.py -> tests/golden/*.npz), and against the
known-answer tests the reference carries (GIoU: tests/test_metrics/test_box_overlap.py:83-97, distance2bbox:
tests/test_utils/test_misc.py:51-64). All file:line citations are into the reference tree (mmdet/...).

Functional style: networks are evaluated from a state_dict with the reference's parameter names.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

INF = 1e8  # models/dense_heads/fcos_head.py:11

# ------------------------------------------------------------------------------------------------ backbone

RESNET_BLOCKS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}  # models/backbones/resnet.py:358-366


def _bn_eval(sd, prefix, x, eps=1e-5):
    """Frozen BatchNorm2d in eval mode (norm_eval=True, resnet.py:647-656)."""
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                        sd[prefix + ".bias"], False, 0.0, eps)


def bottleneck_forward(sd, prefix, x, stride, has_downsample):
    """Caffe-style Bottleneck: the stride sits on conv1 (resnet.py:153-158, 262-301)."""
    identity = x
    out = F.conv2d(x, sd[prefix + ".conv1.weight"], stride=stride)
    out = F.relu(_bn_eval(sd, prefix + ".bn1", out))
    out = F.conv2d(out, sd[prefix + ".conv2.weight"], padding=1)
    out = F.relu(_bn_eval(sd, prefix + ".bn2", out))
    out = F.conv2d(out, sd[prefix + ".conv3.weight"])
    out = _bn_eval(sd, prefix + ".bn3", out)
    if has_downsample:
        identity = F.conv2d(x, sd[prefix + ".downsample.0.weight"], stride=stride)
        identity = _bn_eval(sd, prefix + ".downsample.1", identity)
    return F.relu(out + identity)


def resnet_forward(sd, x, depth=50, prefix=""):
    """ResNet.forward (resnet.py:630-645): stem conv7x7/2 + BN + ReLU + maxpool3x3/2, four stages; returns C2..C5."""
    x = F.conv2d(x, sd[prefix + "conv1.weight"], stride=2, padding=3)
    x = F.relu(_bn_eval(sd, prefix + "bn1", x))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = []
    for li, nblocks in enumerate(RESNET_BLOCKS[depth]):
        for bi in range(nblocks):
            stride = 2 if (bi == 0 and li > 0) else 1
            x = bottleneck_forward(sd, f"{prefix}layer{li + 1}.{bi}", x, stride, bi == 0)
        outs.append(x)
    return outs


def rla_bottleneck_forward(sd, prefix, x, h, stride, has_downsample):
    """RLA_Bottleneck.forward (backbones/resnet_rla.py:105-137), PyTorch style (stride on conv2, :86). `y = out` at :127
    aliases the tensor that `out += identity` (:134) and the in-place ReLU (:135) then rewrite, so the `y` the reference
    returns IS the block output; returns (out, h') with h' = AvgPool2d(2,2)(h) when the block is strided (:93-95,131-132)."""
    identity = x
    out = F.conv2d(torch.cat((x, h), dim=1), sd[prefix + ".conv1.weight"])
    out = F.relu(_bn_eval(sd, prefix + ".bn1", out))
    out = F.conv2d(out, sd[prefix + ".conv2.weight"], stride=stride, padding=1)
    out = F.relu(_bn_eval(sd, prefix + ".bn2", out))
    out = F.conv2d(out, sd[prefix + ".conv3.weight"])
    out = _bn_eval(sd, prefix + ".bn3", out)
    if has_downsample:
        identity = F.conv2d(x, sd[prefix + ".downsample.0.weight"], stride=stride)
        identity = _bn_eval(sd, prefix + ".downsample.1", identity)
        if stride != 1:
            h = F.avg_pool2d(h, 2, 2)
    return F.relu(out + identity), h


def rla_resnet_forward(sd, x, layers=(3, 4, 6, 3), prefix=""):
    """RLA_ResNet._forward_impl (backbones/resnet_rla.py:289-327): stem, then per block the bottleneck on cat(x, h)
    followed by the recurrent-state update h = recurrent_conv(tanh(bn(h + conv_out(y)))) (:306-311) with conv_out /
    recurrent_conv shared inside a stage (:259-260) and one BatchNorm(32) per block (:284); returns the four stage
    outputs x (the state is not part of the outputs, :312-313)."""
    x = F.conv2d(x, sd[prefix + "conv1.weight"], stride=2, padding=3)
    x = F.relu(_bn_eval(sd, prefix + "bn1", x))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    h = x.new_zeros(x.shape[0], sd[prefix + "recurrent_convs.0.weight"].shape[0], x.shape[2], x.shape[3])
    outs = []
    for li, nblocks in enumerate(layers):
        for bi in range(nblocks):
            stride = 2 if (bi == 0 and li > 0) else 1
            x, h = rla_bottleneck_forward(sd, f"{prefix}stages.{li}.{bi}", x, h, stride, bi == 0)
            h = h + F.conv2d(x, sd[f"{prefix}conv_outs.{li}.weight"])
            h = torch.tanh(_bn_eval(sd, f"{prefix}stage_bns.{li}.{bi}", h))
            h = F.conv2d(h, sd[f"{prefix}recurrent_convs.{li}.weight"], padding=1)
        outs.append(x)
    return outs


def fpn_forward(sd, feats, prefix=""):
    """FPN.forward (necks/fpn.py:151-202) with start_level=1, add_extra_convs='on_output', num_outs=5,
    relu_before_extra_convs=True: laterals on C3..C5, nearest top-down, 3x3 outputs, P6 = conv s2 on P5 output,
    P7 = conv s2 on relu(P6)."""
    ins = feats[1:]
    lat = [F.conv2d(ins[i], sd[f"{prefix}lateral_convs.{i}.conv.weight"], sd[f"{prefix}lateral_convs.{i}.conv.bias"])
           for i in range(3)]
    for i in range(2, 0, -1):
        lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:], mode="nearest")
    outs = [F.conv2d(lat[i], sd[f"{prefix}fpn_convs.{i}.conv.weight"], sd[f"{prefix}fpn_convs.{i}.conv.bias"],
                     padding=1) for i in range(3)]
    outs.append(F.conv2d(outs[-1], sd[f"{prefix}fpn_convs.3.conv.weight"], sd[f"{prefix}fpn_convs.3.conv.bias"],
                         stride=2, padding=1))
    outs.append(F.conv2d(F.relu(outs[-1]), sd[f"{prefix}fpn_convs.4.conv.weight"],
                         sd[f"{prefix}fpn_convs.4.conv.bias"], stride=2, padding=1))
    return outs


def fcos_head_forward(sd, feats, strides=(8, 16, 32, 64, 128), training=True, prefix="", stacked_convs=4,
                      num_groups=32):
    """FCOSHead.forward (fcos_head.py:118-168) over AnchorFreeHead.forward_single (anchor_free_head.py:197-217):
    towers of conv3x3+bias -> GroupNorm(32) -> ReLU; conv_cls on the cls tower, conv_reg and conv_centerness on
    the reg tower (centerness_on_reg=True); bbox = relu(scale_l * reg) (norm_on_bbox=True), x stride in eval."""
    cls_scores, bbox_preds, centernesses = [], [], []
    for lvl, x in enumerate(feats):
        cf, rf = x, x
        for i in range(stacked_convs):
            cf = F.conv2d(cf, sd[f"{prefix}cls_convs.{i}.conv.weight"], sd[f"{prefix}cls_convs.{i}.conv.bias"],
                          padding=1)
            cf = F.relu(F.group_norm(cf, num_groups, sd[f"{prefix}cls_convs.{i}.gn.weight"],
                                     sd[f"{prefix}cls_convs.{i}.gn.bias"], 1e-5))
            rf = F.conv2d(rf, sd[f"{prefix}reg_convs.{i}.conv.weight"], sd[f"{prefix}reg_convs.{i}.conv.bias"],
                          padding=1)
            rf = F.relu(F.group_norm(rf, num_groups, sd[f"{prefix}reg_convs.{i}.gn.weight"],
                                     sd[f"{prefix}reg_convs.{i}.gn.bias"], 1e-5))
        cls = F.conv2d(cf, sd[prefix + "conv_cls.weight"], sd[prefix + "conv_cls.bias"], padding=1)
        reg = F.conv2d(rf, sd[prefix + "conv_reg.weight"], sd[prefix + "conv_reg.bias"], padding=1)
        ctr = F.conv2d(rf, sd[prefix + "conv_centerness.weight"], sd[prefix + "conv_centerness.bias"], padding=1)
        bbox = F.relu((reg * sd[f"{prefix}scales.{lvl}.scale"]).float())
        if not training:
            bbox = bbox * strides[lvl]
        cls_scores.append(cls)
        bbox_preds.append(bbox)
        centernesses.append(ctr)
    return cls_scores, bbox_preds, centernesses


# ------------------------------------------------------------------------------------------------ targets

def get_points(featmap_sizes, strides, dtype=torch.float32, device=None):
    """anchor_free_head.py:287-321 + fcos_head.py:550-560: (x, y) = idx * stride + stride // 2, row-major."""
    pts = []
    for (h, w), s in zip(featmap_sizes, strides):
        ys, xs = torch.meshgrid(torch.arange(h, device=device).to(dtype), torch.arange(w, device=device).to(dtype),
                                indexing="ij")
        pts.append(torch.stack((xs.reshape(-1) * s, ys.reshape(-1) * s), dim=-1) + s // 2)
    return pts


def get_target_single(gt_bboxes, gt_labels, points, regress_ranges, strides_per_point, num_classes,
                      center_sampling=True, radius=1.5):
    """fcos_head.py:623-705 for one image over the concatenated points of all levels.
    regress_ranges: (P,2); strides_per_point: (P,) float32 = level stride."""
    P = points.size(0)
    G = gt_labels.size(0)
    if G == 0:
        return gt_labels.new_full((P,), num_classes), gt_bboxes.new_zeros((P, 4))
    areas = (gt_bboxes[:, 2] - gt_bboxes[:, 0]) * (gt_bboxes[:, 3] - gt_bboxes[:, 1])
    areas = areas[None].repeat(P, 1)
    xs = points[:, 0:1].expand(P, G)
    ys = points[:, 1:2].expand(P, G)
    gb = gt_bboxes[None].expand(P, G, 4)
    left = xs - gb[..., 0]
    right = gb[..., 2] - xs
    top = ys - gb[..., 1]
    bottom = gb[..., 3] - ys
    bbox_targets = torch.stack((left, top, right, bottom), -1)
    if center_sampling:
        cx = (gb[..., 0] + gb[..., 2]) / 2
        cy = (gb[..., 1] + gb[..., 3]) / 2
        st = (strides_per_point * radius)[:, None].expand(P, G)
        x_mins, y_mins, x_maxs, y_maxs = cx - st, cy - st, cx + st, cy + st
        c0 = torch.where(x_mins > gb[..., 0], x_mins, gb[..., 0])
        c1 = torch.where(y_mins > gb[..., 1], y_mins, gb[..., 1])
        c2 = torch.where(x_maxs > gb[..., 2], gb[..., 2], x_maxs)
        c3 = torch.where(y_maxs > gb[..., 3], gb[..., 3], y_maxs)
        center_bbox = torch.stack((xs - c0, ys - c1, c2 - xs, c3 - ys), -1)
        inside = center_bbox.min(-1)[0] > 0
    else:
        inside = bbox_targets.min(-1)[0] > 0
    max_reg = bbox_targets.max(-1)[0]
    in_range = (max_reg >= regress_ranges[:, None, 0]) & (max_reg <= regress_ranges[:, None, 1])
    areas[inside == 0] = INF
    areas[in_range == 0] = INF
    min_area, min_inds = areas.min(dim=1)  # first index wins on ties (fcos_head.py:699)
    labels = gt_labels[min_inds]
    labels[min_area == INF] = num_classes
    bbox_targets = bbox_targets[torch.arange(P, device=min_inds.device), min_inds]
    return labels, bbox_targets


def get_targets(points, gt_bboxes_list, gt_labels_list, strides, regress_ranges, num_classes,
                center_sampling=True, radius=1.5, norm_on_bbox=True):
    """fcos_head.py:562-621. Returns per-level lists (images concatenated inside each level)."""
    num_points = [p.size(0) for p in points]
    rr = torch.cat([points[i].new_tensor(regress_ranges[i])[None].expand_as(points[i]) for i in range(len(points))])
    spp = torch.cat([points[i].new_full((num_points[i],), float(strides[i])) for i in range(len(points))])
    cat_points = torch.cat(points, 0)
    per_img = [get_target_single(b, l, cat_points, rr, spp, num_classes, center_sampling, radius)
               for b, l in zip(gt_bboxes_list, gt_labels_list)]
    labels_l, targets_l = [], []
    for i in range(len(points)):
        labels_l.append(torch.cat([lab.split(num_points, 0)[i] for lab, _ in per_img]))
        t = torch.cat([bt.split(num_points, 0)[i] for _, bt in per_img])
        if norm_on_bbox:
            t = t / strides[i]
        targets_l.append(t)
    return labels_l, targets_l


def centerness_target(pos_bbox_targets):
    """fcos_head.py:707-726."""
    if pos_bbox_targets.numel() == 0:
        return pos_bbox_targets.new_zeros((0,))
    lr = pos_bbox_targets[:, [0, 2]]
    tb = pos_bbox_targets[:, [1, 3]]
    c = (lr.min(-1)[0] / lr.max(-1)[0]) * (tb.min(-1)[0] / tb.max(-1)[0])
    return torch.sqrt(c)


# ------------------------------------------------------------------------------------------------ losses

def distance2bbox(points, distance, max_shape=None):
    """core/bbox/transforms.py:119-162 (max_shape = (H, W[, C]) clips to [0,W] x [0,H])."""
    x1 = points[..., 0] - distance[..., 0]
    y1 = points[..., 1] - distance[..., 1]
    x2 = points[..., 0] + distance[..., 2]
    y2 = points[..., 1] + distance[..., 3]
    b = torch.stack([x1, y1, x2, y2], -1)
    if max_shape is not None:
        h, w = float(max_shape[0]), float(max_shape[1])
        mx = b.new_tensor([w, h, w, h])
        b = torch.where(b < 0, b.new_tensor(0.0), b)
        b = torch.where(b > mx, mx, b)
    return b


def giou_aligned(b1, b2, eps=1e-6):
    """core/bbox/iou_calculators/iou2d_calculator.py:214-260, mode='giou', is_aligned=True."""
    if b1.size(0) == 0:
        return b1.new_zeros((0,))
    a1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    a2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    lt = torch.max(b1[:, :2], b2[:, :2])
    rb = torch.min(b1[:, 2:], b2[:, 2:])
    wh = (rb - lt).clamp(min=0)
    overlap = wh[:, 0] * wh[:, 1]
    e = a1.new_tensor([eps])
    union = torch.max(a1 + a2 - overlap, e)
    ious = overlap / union
    elt = torch.min(b1[:, :2], b2[:, :2])
    erb = torch.max(b1[:, 2:], b2[:, 2:])
    ewh = (erb - elt).clamp(min=0)
    earea = torch.max(ewh[:, 0] * ewh[:, 1], e)
    return ious - (earea - union) / earea


def sigmoid_focal_loss_elem(pred, labels, num_classes, gamma=2.0, alpha=0.25):
    """losses/focal_loss.py:11-56 with the one-hot of :165-167 (label == num_classes => all-negative row)."""
    target = F.one_hot(labels, num_classes=num_classes + 1)[:, :num_classes].type_as(pred)
    p = pred.sigmoid()
    pt = (1 - p) * target + p * (1 - target)
    fw = (alpha * target + (1 - alpha) * (1 - target)) * pt.pow(gamma)
    return F.binary_cross_entropy_with_logits(pred, target, reduction="none") * fw


def unlabeled_weights(num_per_level_img, batch, loss_weight, device=None):
    """fcos_head.py:217-235: per level, the first half of the (image-major) points is 'labeled' (x1), the rest
    'unlabeled' (x loss_weight); with an odd batch (scale-invariant extra image) the labeled part is the first
    (B-1)/2 images."""
    out = []
    for n in num_per_level_img:  # n = B * points_of_level
        w = torch.ones(n, dtype=torch.float32, device=device)
        if batch % 2 == 0:
            w[int(n / 2):] *= loss_weight
        else:
            w[int(n / batch * (batch - 1) / 2):] *= loss_weight
        out.append(w)
    return torch.cat(out)


def fcos_loss(cls_scores, bbox_preds, centernesses, gt_bboxes, gt_labels, gt_bboxes_ignore=None, *,
              strides=(8, 16, 32, 64, 128),
              regress_ranges=((-1, 64), (64, 128), (128, 256), (256, 512), (512, INF)), num_classes=80,
              center_sampling=True, radius=1.5, norm_on_bbox=True, loss_weight=1.0, soft_weight=0.0,
              soft_warm_up=0, cur_iter=0, world_num_pos=None, world_ctr_sum=None, return_aux=False):
    """FCOSHead.loss (fcos_head.py:170-338) incl. the DSL additions (ignore regions :208-215/297-304, unlabeled
    weights :217-235/281-292/305-307, scale-invariant soft loss :312-333). world_* override the (rank-averaged)
    reduce_mean values (:266,274) for multi-rank checks; default = this rank's own values (world size 1)."""
    B = cls_scores[0].size(0)
    sizes = [c.shape[-2:] for c in cls_scores]
    dev = cls_scores[0].device
    points = get_points(sizes, strides, device=dev)
    labels, bbox_targets = get_targets(points, gt_bboxes, gt_labels, strides, regress_ranges, num_classes,
                                       center_sampling, radius, norm_on_bbox)
    ig_labels = None
    if gt_bboxes_ignore is not None:
        ig_lab = [torch.zeros(b.size(0), dtype=torch.int64, device=b.device) + num_classes - 1 for b in gt_bboxes_ignore]
        ig_labels, _ = get_targets(points, gt_bboxes_ignore, ig_lab, strides, regress_ranges, num_classes,
                                   center_sampling, radius, norm_on_bbox)
    As = None
    if loss_weight != 1.0:
        As = unlabeled_weights([l.numel() for l in ig_labels], B, loss_weight, device=dev)

    f_cls = torch.cat([c.permute(0, 2, 3, 1).reshape(-1, num_classes) for c in cls_scores])
    f_box = torch.cat([b.permute(0, 2, 3, 1).reshape(-1, 4) for b in bbox_preds])
    f_ctr = torch.cat([c.permute(0, 2, 3, 1).reshape(-1) for c in centernesses])
    f_lab = torch.cat(labels)
    f_tgt = torch.cat(bbox_targets)
    f_pts = torch.cat([p.repeat(B, 1) for p in points])

    pos = ((f_lab >= 0) & (f_lab < num_classes)).nonzero().reshape(-1)
    n_pos_local = float(len(pos))
    num_pos = max(n_pos_local if world_num_pos is None else world_num_pos, 1.0)
    pos_tgt = f_tgt[pos]
    ctr_t = centerness_target(pos_tgt)
    ctr_sum_local = float(ctr_t.sum()) if len(pos) else 0.0
    denorm = max(ctr_sum_local if world_ctr_sum is None else world_ctr_sum, 1e-6)

    if len(pos) > 0:
        dec_p = distance2bbox(f_pts[pos], f_box[pos])
        dec_t = distance2bbox(f_pts[pos], pos_tgt)
        fw = torch.ones_like(ctr_t)
        if As is not None:
            fw = fw * As[pos]
        w_box = ctr_t * fw
        if not bool((w_box > 0).any()):  # losses/iou_loss.py:345-348
            loss_bbox = (dec_p * w_box[:, None]).sum()
        else:
            loss_bbox = ((1 - giou_aligned(dec_p, dec_t)) * w_box).sum() / denorm
        loss_ctr = (F.binary_cross_entropy_with_logits(f_ctr[pos], ctr_t, reduction="none") * fw).sum() / num_pos
    else:
        loss_bbox = f_box[pos].sum()
        loss_ctr = f_ctr[pos].sum()

    weight = torch.ones_like(f_lab, dtype=torch.float32)
    if ig_labels is not None:
        f_ig = torch.cat(ig_labels).clone()
        inter = ((f_ig - num_classes) * (f_lab - num_classes)).nonzero().reshape(-1)
        if inter.numel() > 0:
            f_ig[inter] = num_classes
        weight = f_ig.float() - num_classes + 1
    if As is not None:
        weight = weight * As
    loss_cls = (sigmoid_focal_loss_elem(f_cls, f_lab, num_classes) * weight[:, None]).sum() / num_pos

    out = dict(loss_cls=loss_cls, loss_bbox=loss_bbox, loss_centerness=loss_ctr)
    if B % 2 != 0 and soft_weight != 0.0:
        si = 0.0
        for i in range(1, len(cls_scores)):
            h, w = cls_scores[i].shape[-2:]
            d = cls_scores[i][B - 2] - cls_scores[i - 1][B - 1][:, :h, :w]
            si = si + (d * d).mean()
        sw = soft_weight * 1.0
        if soft_warm_up >= cur_iter:  # fcos_head.py:325-327 (the caller advances cur_iter while warming up)
            sw = soft_weight / 1000.0
        out["loss_sisoft"] = si * sw
    if return_aux:
        out["_aux"] = dict(labels=f_lab, bbox_targets=f_tgt, weight=weight, pos_inds=pos, centerness_targets=ctr_t,
                           num_pos_local=n_pos_local, ctr_sum_local=ctr_sum_local, points=f_pts)
    return out


# ------------------------------------------------------------------------------------------------ teacher side

def nms_greedy(boxes, scores, thr):
    """mmcv.ops.nms semantics (offset 0): sort by score desc, suppress IoU > thr. numpy, O(n^2); n <= 5000."""
    boxes = np.asarray(boxes, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    order = np.argsort(-scores, kind="stable")
    keep = []
    area = (boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])
    supp = np.zeros(len(boxes), dtype=bool)
    for idx_i, i in enumerate(order):
        if supp[i]:
            continue
        keep.append(i)
        rest = order[idx_i + 1:]
        xx1 = np.maximum(boxes[i, 0], boxes[rest, 0])
        yy1 = np.maximum(boxes[i, 1], boxes[rest, 1])
        xx2 = np.minimum(boxes[i, 2], boxes[rest, 2])
        yy2 = np.minimum(boxes[i, 3], boxes[rest, 3])
        inter = np.maximum(xx2 - xx1, 0) * np.maximum(yy2 - yy1, 0)
        iou = inter / (area[i] + area[rest] - inter)
        supp[rest[iou > thr]] = True
    return np.asarray(keep, dtype=np.int64)


def decode_candidates(cls_scores, bbox_preds, centernesses, img_shapes, scale_factors, strides=(8, 16, 32, 64, 128),
                      nms_pre=1000, score_thr=0.05, rescale=True):
    """FCOSHead._get_bboxes (fcos_head.py:406-527) up to the NMS, plus the gate of multiclass_nms
    (core/post_processing/bbox_nms.py:34-67): per level sigmoid, top-nms_pre of max_c(score*ctr), decode + clip,
    /scale_factor, keep class scores > score_thr (raw score), score *= centerness.
    Returns per image (boxes (n,4), scores (n,), labels (n,), flat candidate index (n,))."""
    B = cls_scores[0].size(0)
    C = cls_scores[0].size(1)
    points = get_points([c.shape[-2:] for c in cls_scores], strides, device=cls_scores[0].device)
    mb, ms, mc = [], [], []
    for cls, box, ctr, pts in zip(cls_scores, bbox_preds, centernesses, points):
        scores = cls.permute(0, 2, 3, 1).reshape(B, -1, C).sigmoid()
        cn = ctr.permute(0, 2, 3, 1).reshape(B, -1).sigmoid()
        bp = box.permute(0, 2, 3, 1).reshape(B, -1, 4)
        pp = pts.expand(B, -1, 2)
        if 0 < nms_pre < bp.shape[1]:
            mx, _ = (scores * cn[..., None]).max(-1)
            _, topk = mx.topk(nms_pre)
            bi = torch.arange(B, device=topk.device).view(-1, 1).expand_as(topk)
            pp, bp, scores, cn = pp[bi, topk], bp[bi, topk], scores[bi, topk], cn[bi, topk]
        boxes = torch.stack([distance2bbox(pp[b], bp[b], max_shape=img_shapes[b]) for b in range(B)])
        mb.append(boxes)
        ms.append(scores)
        mc.append(cn)
    mb, ms, mc = torch.cat(mb, 1), torch.cat(ms, 1), torch.cat(mc, 1)
    if rescale:
        mb = mb / mb.new_tensor(np.asarray(scale_factors, dtype=np.float32)).unsqueeze(1)
    out = []
    for b in range(B):
        sc = ms[b].reshape(-1)
        valid = sc > score_thr
        sc2 = sc * mc[b].view(-1, 1).expand(-1, C).reshape(-1)
        inds = valid.nonzero().squeeze(1)
        out.append((mb[b][:, None].expand(-1, C, 4).reshape(-1, 4)[inds], sc2[inds], (inds % C), inds))
    return out


def multiclass_nms(boxes, scores, labels, iou_thr=0.6, max_per_img=100):
    """mmcv batched_nms (class-offset trick) as called from bbox_nms.py:85-94."""
    if boxes.numel() == 0:
        return boxes.new_zeros((0, 5)), labels
    off = labels.to(boxes) * (boxes.max() + 1)
    keep = nms_greedy((boxes + off[:, None]).numpy(), scores.numpy(), iou_thr)
    keep = torch.from_numpy(keep)[:max_per_img]
    return torch.cat([boxes[keep], scores[keep, None]], -1), labels[keep]


def parse_det_results(dets, labels, score_thr=0.1):
    """runner/hooks/unlabel_pred_hook.py:20-38: keep score >= thr, truncate coordinates with int()."""
    out = []
    for (x1, y1, x2, y2, s), c in zip(dets.tolist(), labels.tolist()):
        if s < score_thr:
            continue
        out.append(dict(bbox=[int(x1), int(y1), int(x2), int(y2)], score=round(float(s), 6), category_index=int(c)))
    return out  # gen_save_json_dict (:55) later sorts by score; order here = input order, as in the reference


def filter_pseudo_labels(rects, scores, cls_ids, img_w, img_h, thres_by_class=None, default_thres=(0.1, 0.3)):
    """datasets/semicoco.py:220-269: drop boxes with no overlap with the image or w/h < 1; a box whose score lies in
    [default_thres[0], thr_c) becomes an ignore region, every other box (incl. score < 0.1) becomes GT.
    thr_c = thres_by_class[c] if present else default_thres[1]."""
    gt, gl, ig = [], [], []
    for (x1, y1, x2, y2), s, c in zip(rects, scores, cls_ids):
        iw = max(0, min(x2, img_w) - max(x1, 0))
        ih = max(0, min(y2, img_h) - max(y1, 0))
        if iw * ih == 0 or x2 - x1 < 1 or y2 - y1 < 1:
            continue
        thr = default_thres[1]
        if thres_by_class is not None and c in thres_by_class:
            thr = thres_by_class[c]
        if s < float(thr) and s >= float(default_thres[0]):
            ig.append([x1, y1, x2, y2])
        else:
            gt.append([x1, y1, x2, y2])
            gl.append(int(c))
    gt = np.array(gt, dtype=np.float32).reshape(-1, 4)
    ig = np.array(ig, dtype=np.float32).reshape(-1, 4)
    return gt, np.array(gl, dtype=np.int64), ig


def hook_saved_boxes(dets, labels, num_classes, infer_score_thr=0.1, iou=0.6):
    """What UnlabelPredHook.save_results2file writes to the image's JSON file, as (rects, scores, class ids):
    runner/hooks/unlabel_pred_hook.py:20-38 (gate score >= thr, int() truncation, round(score, 6)), :55 (stable sort by
    score, descending), :142-165 (per class in range(0, len(id2cat) - 1): the reference's category file carries a
    trailing background entry, tools/coco_convert2_semicoco_json.py:47-48, so that is every one of the num_classes
    real classes — nms(iou, score_threshold=0.1) on the truncated fp32 boxes). These are also the boxes adathres() counts (:315-343)."""
    dets = np.asarray(dets, dtype=np.float32).reshape(-1, 5)
    labels = np.asarray(labels).reshape(-1)
    items = []
    for c in range(num_classes):  # bbox2result groups by class, keeping the detection order inside a class
        for d in dets[labels == c]:
            if float(d[4]) < infer_score_thr:
                continue
            items.append(([int(d[0]), int(d[1]), int(d[2]), int(d[3])], round(float(d[4]), 6), c))
    items.sort(key=lambda t: t[1], reverse=True)  # stable
    rects, scores, cls = [], [], []
    if items:
        b = np.array([t[0] for t in items], dtype=np.float32)
        s = np.array([t[1] for t in items], dtype=np.float32)
        c = np.array([t[2] for t in items], dtype=np.float32)
        for i in range(0, num_classes):
            sel = c == i
            if not sel.any():
                continue
            bi, si = b[sel], s[sel]
            v = si > np.float32(0.1)
            bi, si = bi[v], si[v]
            if len(si) == 0:
                continue
            order = np.argsort(-si, kind="stable")
            bi, si = bi[order], si[order]
            keep = nms_greedy(bi, si, iou)
            for k in keep:
                rects.append(bi[k].tolist())
                scores.append(float(si[k]))
                cls.append(i)
    return rects, scores, cls


def hook_pseudo_labels(dets, labels, num_classes, img_w, img_h, thr_by_class, infer_score_thr=0.1, iou=0.6,
                       default_thres=(0.1, 0.3)):
    """Detections of one image (multiclass_nms order: score descending) -> (gt_bboxes, gt_labels, gt_bboxes_ignore):
    hook_saved_boxes (the JSON the hook writes) followed by datasets/semicoco.py:220-269 (filter_pseudo_labels).
    thr_by_class: sequence of per-class thresholds (fp64)."""
    rects, scores, cls = hook_saved_boxes(dets, labels, num_classes, infer_score_thr, iou)
    thr = {i: float(t) for i, t in enumerate(thr_by_class)}
    return filter_pseudo_labels(rects, scores, cls, img_w, img_h, thr, default_thres)


def adathres(scores_by_class, prev_thres=None, ranges=(0.3, 0.35), gamma1=0.05, gamma2=0.6, base=0.3):
    """runner/hooks/unlabel_pred_hook.py:295-367. scores_by_class: {class: [scores of all pseudo boxes]}.
    A box is counted if score >= 0.3 (first pass) or >= last epoch's thr_c (class absent from history => counted).
    thr_c = clip((sum_c / mean_count)^gamma1 * base, ranges); weight_c = (mean_count / sum_c)^gamma2."""
    dis, cum = {}, {}
    for c, ss in scores_by_class.items():
        for s in ss:
            if prev_thres is None:
                ok = s >= 0.3
            else:
                ok = (c not in prev_thres) or (s >= prev_thres[c])
            if ok:
                dis[c] = dis.get(c, 0) + 1
                cum[c] = cum.get(c, 0.0) + s
    avg = sum(dis.values())
    weights = {c: (avg / len(dis) / cum[c]) ** gamma2 for c in dis}
    thres = {c: max(min((cum[c] / (avg / len(dis))) ** gamma1 * base, ranges[1]), ranges[0]) for c in dis}
    return thres, weights


# ------------------------------------------------------------------------------------------------ step pieces

def ema_update(teacher_sd, student_sd, keep_rate):
    """runner/hooks/semi_epoch_based_runner.py:392-406: T <- (1-k) S + k T over every state_dict entry
    (float math, incl. BN buffers)."""
    return {k: student_sd[k].float() * (1 - keep_rate) + v.float() * keep_rate for k, v in teacher_sd.items()}


def clip_grad_norm(grads, max_norm=35.0, norm_type=2.0):
    """torch.nn.utils.clip_grad_norm_ as called by mmcv OptimizerHook (cfg grad_clip max_norm=35, norm_type=2)."""
    total = torch.norm(torch.stack([torch.norm(g, norm_type) for g in grads]), norm_type)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return [g * coef for g in grads], total


def sgd_momentum_step(p, g, buf, lr, momentum=0.9, weight_decay=1e-4, first=False):
    """torch.optim.SGD (dampening 0, no nesterov): g += wd*p; buf = g (first step) or m*buf + g; p -= lr*buf."""
    g = g + weight_decay * p
    buf = g.clone() if first else momentum * buf + g
    return p - lr * buf, buf


def scale_invariant_input(img, gt_bboxes, gt_bboxes_ignore):
    """semi_epoch_based_runner.py:186-204: bilinear half-resolution copy of the LAST image, zero padded to the
    batch H x W; its boxes / ignore boxes are halved."""
    h, w = img.shape[-2:]
    half = F.interpolate(img[-1:].clone(), (int(h / 2), int(w / 2)), mode="bilinear")
    tmp = torch.zeros_like(img[-1:])
    tmp[:, :, :int(h / 2), :int(w / 2)] = half
    ig = gt_bboxes_ignore[-1].clone()
    if len(ig) > 0:
        ig = ig / 2
    return torch.cat((img, tmp), 0), gt_bboxes[-1].clone() / 2, ig


_ = math


def view_boxes(boxes, labels, sx, sy, img_w, img_h, clip=True, ps_mode=0, ps_crop=0, flip=False):
    """Box part of Resize._resize_bboxes -> PatchShuffle.__call__ -> RandomFlip.bbox_flip('horizontal')
    (mmdet/datasets/pipelines/transforms.py:249-257, 2168-2248, 397-429), NumPy float32 arithmetic left to right.
    ps_mode 0 off / 1 'flip' (vertical cut at column ps_crop) / 2 'flop' (horizontal cut at row ps_crop).
    Returns (boxes (m,4) float32, labels (m,) int64 or None); a box straddling the cut becomes two."""
    import numpy as np
    f = np.float32
    b = np.asarray(boxes, np.float32).reshape(-1, 4) * np.array([sx, sy, sx, sy], dtype=np.float32)
    if clip:
        b[:, 0::2] = np.clip(b[:, 0::2], 0, img_w)
        b[:, 1::2] = np.clip(b[:, 1::2], 0, img_h)
    lab = None if labels is None else list(np.asarray(labels).tolist())
    w, h = f(img_w), f(img_h)
    active = (ps_mode == 1 and ps_crop not in (0, img_w)) or (ps_mode == 2 and ps_crop not in (0, img_h))
    if active and len(b):
        cw = f(ps_crop) if ps_mode == 1 else w
        chh = f(ps_crop) if ps_mode == 2 else h
        ob, ol = [], []
        for i in range(len(b)):
            x1, y1, x2, y2 = (f(v) for v in b[i])
            one = f(1)
            if (x1 - cw + one) * (x2 - cw + one) >= 0 and (y1 - chh + one) * (y2 - chh + one) >= 0:
                if ps_mode == 1:
                    if x1 - cw + one < 0:
                        x1, x2 = x1 + w - cw, x2 + w - cw
                    if x2 - cw + one > 0:
                        x1, x2 = x1 - cw, x2 - cw
                else:
                    if y1 - chh + one < 0:
                        y1, y2 = y1 + h - chh, y2 + h - chh
                    if y2 - chh + one > 0:
                        y1, y2 = y1 - chh, y2 - chh
                ob.append([x1, y1, x2, y2])
                if lab is not None:
                    ol.append(lab[i])
            else:
                if ps_mode == 1:
                    ob += [[x1 + w - cw, y1, w - one, y2], [f(0), y1, x2 - cw, y2]]
                else:
                    ob += [[x1, y1 + h - chh, x2, h - one], [x1, f(0), x2, y2 - chh]]
                if lab is not None:
                    ol += [lab[i], lab[i]]
        b = np.array(ob, dtype=np.float32).reshape(-1, 4)
        lab = ol if lab is not None else None
    if flip and len(b):
        fb = b.copy()
        fb[:, 0] = w - b[:, 2]
        fb[:, 2] = w - b[:, 0]
        b = fb
    return b, (None if lab is None else np.asarray(lab, np.int64))
