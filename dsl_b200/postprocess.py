"""Teacher-side post-processing on the device, through the C ABI (libdslb.so): per-level top-nms_pre + decode + score
gate (FCOSHead._get_bboxes, mmdet/models/dense_heads/fcos_head.py:406-527; multiclass_nms gate,
mmdet/core/post_processing/bbox_nms.py:34-67), class-aware NMS (bbox_nms.py:78-94) and the pseudo-label rule chain of
UnlabelPredHook + SemiCOCODataset (mmdet/runner/hooks/unlabel_pred_hook.py:20-38,142-165; mmdet/datasets/
semicoco.py:220-269). The reference does the last two on the host and through JSON files; here nothing leaves HBM.
"""
import torch

from . import _lib as L


class TeacherPost:
    """Buffers + launches for B images whose head outputs are pixel-major fp32 level tensors:
    cls_out[l] [B, h, w, C] and rc_out[l] [B, h, w, 8] (0-3 distances already x stride, 4 centerness logit)."""

    def __init__(self, B, level_sizes, strides, num_classes, device, nms_pre=1000, score_thr=0.05, iou_thr=0.6,
                 max_per_img=100, cand_cap=8192, max_boxes=1024):
        self.B, self.psize, self.strides, self.C = B, list(level_sizes), tuple(strides), num_classes
        self.dev = torch.device(device)
        self.nms_pre, self.score_thr, self.iou_thr, self.max_per_img = nms_pre, score_thr, iou_thr, max_per_img
        assert max_per_img <= 128
        self.cand_cap = cand_cap
        f32, i32 = torch.float32, torch.int32
        from .arena import Arena
        self.mem = Arena(self.dev, first_chunk=8 << 20)
        z = lambda *s, dtype=f32: self.mem.zeros(*s, dtype=dtype)  # noqa: E731
        self.pt_scores = [z(B, h * w) for (h, w) in self.psize]
        self.cand_boxes = z(B, cand_cap, 4)
        self.cand_scores = z(B, cand_cap)
        self.cand_labels = z(B, cand_cap, dtype=i32)
        self.cand_points = z(B, cand_cap, dtype=i32)
        self.cand_counts = z(B, dtype=i32)
        self.cand_overflow = z(1, dtype=i32)   # sticky: some decode since the last check exceeded cand_cap
        self._tk = None
        self.sel = {}
        self.img_hw = z(B, 2)          # img_shape (H, W): clip range of the decoded boxes
        self.scale_factor = torch.ones(B, 4, dtype=f32, device=self.dev)
        self.ws_bytes = L.lib.dslb_nms_workspace_bytes(B, cand_cap)
        self.ws = z(self.ws_bytes, dtype=torch.uint8)
        self.dets = z(B, max_per_img, 5)
        self.det_labels = z(B, max_per_img, dtype=i32)
        self.det_count = z(B, dtype=i32)
        # pseudo-label rule parameters (cfg data.unlabel_pred.infer_score_thre / eval_config.iou; semicoco default_thres)
        self.thr_class = torch.full((num_classes,), 0.3, dtype=torch.float64, device=self.dev)
        self.img_wh = z(B, 2)          # (width, height) of the ORIGINAL image the rescaled boxes live in
        self.max_boxes = max_boxes
        self.rescale = True
        # adaptive-threshold statistics of the running epoch (unlabel_pred_hook.py:295-343), accumulated on the device
        self.stat_cnt = z(num_classes, dtype=torch.int64)
        self.stat_cum = z(num_classes, dtype=torch.float64)
        self.stat_prev = z(num_classes, dtype=torch.float64)   # last epoch's thresholds (-inf: class not in the history)
        self.have_prev = False
        self.class_weight = z(num_classes, dtype=torch.float64)

    def set_meta(self, img_shapes, scale_factors=None, ori_shapes=None):
        """img_shapes: [(H, W, ...)] per image (img_metas['img_shape']); scale_factors: [4 floats] per image or None."""
        hw = torch.tensor([[float(s[0]), float(s[1])] for s in img_shapes], dtype=torch.float32)
        self.img_hw.copy_(hw, non_blocking=True)
        self.rescale = scale_factors is not None
        if self.rescale:
            self.scale_factor.copy_(torch.tensor([[float(v) for v in sf] for sf in scale_factors], dtype=torch.float32),
                                    non_blocking=True)
        if ori_shapes is None:
            ori_shapes = img_shapes
        self.img_wh.copy_(torch.tensor([[float(s[1]), float(s[0])] for s in ori_shapes], dtype=torch.float32),
                          non_blocking=True)

    def set_class_thresholds(self, thr):
        """Per-class ignore thresholds (adathres.json "thres" of the reference), python floats / fp64."""
        self.thr_class.copy_(torch.as_tensor(thr, dtype=torch.float64), non_blocking=True)

    def _topk_plan(self):
        """Levels with more points than nms_pre: their selections come from ONE dslb_fcos_topk_points launch."""
        import ctypes as C
        lv = [l for l, (h, w) in enumerate(self.psize) if 0 < self.nms_pre < h * w]
        self.sel = {l: torch.zeros(self.B, self.nms_pre, dtype=torch.int64, device=self.dev) for l in lv}
        n = len(lv)
        self._tk = (lv, (C.c_void_p * max(n, 1))(*[self.pt_scores[l].data_ptr() for l in lv]),
                    (C.c_void_p * max(n, 1))(*[self.sel[l].data_ptr() for l in lv]),
                    (C.c_int32 * max(n, 1))(*[self.psize[l][0] * self.psize[l][1] for l in lv]),
                    (C.c_int32 * max(n, 1))(*[self.nms_pre for _ in lv]))

    def decode(self, cls_out, rc_out):
        """Per level: top-nms_pre points by max_c(score * centerness), decode + clip + rescale, score gate."""
        s = L.cur_stream
        if getattr(self, "_tk", None) is None:
            self._topk_plan()
        L.zero(self.cand_counts)
        for l, (h, w) in enumerate(self.psize):
            L.check(L.lib.dslb_fcos_point_scores(L.ptr(cls_out[l]), L.ptr(rc_out[l]), L.ptr(self.pt_scores[l]),
                                                 self.B * h * w, self.C, self.C, s()), "point_scores")
        lv, sc_p, sel_p, n_p, k_p = self._tk
        if lv:   # `max_scores.topk(nms_pre)` (fcos_head.py:452-460) for every level that needs it, one launch
            L.check(L.lib.dslb_fcos_topk_points(sc_p, sel_p, n_p, k_p, len(lv), self.B, s()), "topk_points")
        off = 0
        for l, (h, w) in enumerate(self.psize):
            n = h * w
            if l in self.sel:
                K, selp = self.nms_pre, L.ptr(self.sel[l])
            else:
                K, selp = n, None
            L.check(L.lib.dslb_fcos_decode_gate(
                L.ptr(cls_out[l]), L.ptr(rc_out[l]), selp, self.B, K, self.C, h, w, self.strides[l], self.C,
                L.ptr(self.img_hw), L.ptr(self.scale_factor) if self.rescale else None, float(self.score_thr), off,
                L.ptr(self.cand_boxes), L.ptr(self.cand_scores), L.ptr(self.cand_labels), L.ptr(self.cand_points),
                L.ptr(self.cand_counts), self.cand_cap, L.ptr(self.cand_overflow), s()), "decode_gate")
            off += n

    def overflowed(self, clear=True):
        """True if some image produced more gated candidates than `cand_cap` in any decode since the last check (host
        sync; the flag is sticky on the device). The reference has no cap (up to levels x nms_pre x classes scores can
        pass the gate); beyond it the slots are claimed in arrival order, so WHICH candidates are dropped is not
        deterministic. 8192 is several times what a trained FCOS yields at score_thr 0.05; DSLEngine.end_epoch checks
        the flag and raises."""
        hit = bool(self.cand_overflow.item())
        if hit and clear:
            self.cand_overflow.zero_()
        return hit

    def nms(self):
        """multiclass_nms: survivors in self.dets / det_labels / det_count (score-descending, <= max_per_img)."""
        L.check(L.lib.dslb_multiclass_nms(
            L.ptr(self.cand_boxes), L.ptr(self.cand_scores), L.ptr(self.cand_labels), L.ptr(self.cand_points),
            L.ptr(self.cand_counts), self.B, self.cand_cap, self.C, float(self.iou_thr), self.max_per_img,
            L.ptr(self.ws), self.ws_bytes, L.ptr(self.dets), L.ptr(self.det_labels), L.ptr(self.det_count),
            L.cur_stream()), "multiclass_nms")

    def pseudo_labels(self, gt_boxes, gt_labels, gt_off, ig_boxes, ig_off, infer_score_thr=0.1, hook_iou=0.6,
                      ignore_lo=0.1, accumulate_stats=False):
        """Detections -> packed (pseudo GT, ignore) box lists written straight into a student's target buffers.
        accumulate_stats: also add this batch to the epoch's adathres statistics (per-class count / score sum of the
        boxes the reference's hook would have written to JSON and adathres() would have counted)."""
        L.check(L.lib.dslb_pseudo_labels_stats(
            L.ptr(self.dets), L.ptr(self.det_labels), L.ptr(self.det_count), L.ptr(self.thr_class), L.ptr(self.img_wh),
            self.B, self.max_per_img, self.C, float(infer_score_thr), float(hook_iou), float(ignore_lo),
            int(gt_boxes.shape[0]), L.ptr(gt_boxes), L.ptr(gt_labels), L.ptr(gt_off), L.ptr(ig_boxes), L.ptr(ig_off),
            L.ptr(self.stat_cnt) if accumulate_stats else None, L.ptr(self.stat_cum) if accumulate_stats else None,
            L.ptr(self.stat_prev) if (accumulate_stats and self.have_prev) else None, L.cur_stream()), "pseudo_labels")

    def saved_records(self, image_names, cat_names, infer_score_thr=0.1, hook_iou=0.6, ignore_lo=0.1):
        """What UnlabelPredHook.save_results2file would write for the current detections, one dict per image in the
        reference's per-image JSON layout (formats.pseudo_label_record; unlabel_pred_hook.py:142-175) — for handing the
        teacher's pseudo labels to the reference's SemiCOCODataset. Runs the rule chain once more with the saved-list
        export on (scratch GT / ignore outputs, no statistics) and copies the lists to the host (synchronises)."""
        from . import formats
        assert len(image_names) == self.B
        if getattr(self, "_sv", None) is None:
            z = lambda *s, dtype=torch.float32: torch.zeros(*s, dtype=dtype, device=self.dev)  # noqa: E731
            self._sv = dict(boxes=z(self.B, self.max_per_img, 4), scores=z(self.B, self.max_per_img),
                            labels=z(self.B, self.max_per_img, dtype=torch.int32), count=z(self.B, dtype=torch.int32),
                            gt=z(self.max_boxes, 4), gl=z(self.max_boxes, dtype=torch.int64),
                            go=z(self.B + 1, dtype=torch.int32), ig=z(self.max_boxes, 4),
                            io=z(self.B + 1, dtype=torch.int32))
        v = self._sv
        L.check(L.lib.dslb_pseudo_labels_saved(
            L.ptr(self.dets), L.ptr(self.det_labels), L.ptr(self.det_count), L.ptr(self.thr_class), L.ptr(self.img_wh),
            self.B, self.max_per_img, self.C, float(infer_score_thr), float(hook_iou), float(ignore_lo), self.max_boxes,
            L.ptr(v["gt"]), L.ptr(v["gl"]), L.ptr(v["go"]), L.ptr(v["ig"]), L.ptr(v["io"]), L.ptr(v["boxes"]),
            L.ptr(v["scores"]), L.ptr(v["labels"]), L.ptr(v["count"]), L.cur_stream()), "pseudo_labels_saved")
        cnt = v["count"].cpu().tolist()
        boxes, scores, labels = v["boxes"].cpu(), v["scores"].cpu(), v["labels"].cpu()
        return [formats.pseudo_label_record(image_names[b], boxes[b, :cnt[b]].tolist(), scores[b, :cnt[b]].tolist(),
                                            labels[b, :cnt[b]].tolist(), cat_names) for b in range(self.B)]

    def adathres_reset(self, forget_history=False):
        self.stat_cnt.zero_()
        self.stat_cum.zero_()
        if forget_history:
            self.have_prev = False

    def adathres_update(self, gamma1=0.05, gamma2=0.6, base=0.3, ranges=(0.3, 0.35), default_thres=0.3):
        """End of an epoch: adathres() of the reference (unlabel_pred_hook.py:295-367) on the device. Statistics are
        summed over ranks first (the reference's rank 0 reads every rank's JSON files). Installs the new per-class
        thresholds for the pseudo-label rule, keeps them as next epoch's counting gate, clears the accumulators."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.stat_cnt)
            dist.all_reduce(self.stat_cum)
        L.check(L.lib.dslb_adathres_finalize(
            L.ptr(self.stat_cnt), L.ptr(self.stat_cum), self.C, float(gamma1), float(gamma2), float(base),
            float(ranges[0]), float(ranges[1]), float(default_thres), L.ptr(self.thr_class), L.ptr(self.class_weight),
            L.ptr(self.stat_prev), L.cur_stream()), "adathres_finalize")
        self.have_prev = True
        self.adathres_reset()
        return self.thr_class, self.class_weight

    def adathres_state(self):
        """Host copy of the epoch's thresholds: (thr per class, class weight per class — 0 = class not counted, i.e.
        absent from the reference's adathres.json). formats.adathres_to_json turns it into the reference's file."""
        return self.thr_class.cpu().tolist(), self.class_weight.cpu().tolist()

    def load_adathres(self, thr, counted):
        """Resume from the reference's adathres.json (formats.adathres_from_json): installs the per-class thresholds and
        makes them the history the next epoch's counting gate reads (absent classes: every box counts, as in
        unlabel_pred_hook.py:328-337). Engines that captured a graph without a history must capture again."""
        t = torch.as_tensor(thr, dtype=torch.float64)
        self.thr_class.copy_(t, non_blocking=False)
        prev = torch.where(torch.as_tensor(counted, dtype=torch.bool), t, torch.full_like(t, float("-inf")))
        self.stat_prev.copy_(prev, non_blocking=False)
        self.have_prev = True

    def results(self):
        """Host copy: [(dets (n,5) fp32, labels (n,) int64)] per image — what FCOSHead.get_bboxes returns."""
        cnt = self.det_count.cpu().tolist()
        dets, labels = self.dets.cpu(), self.det_labels.cpu()
        return [(dets[b, :cnt[b]].clone(), labels[b, :cnt[b]].to(torch.int64)) for b in range(self.B)]
