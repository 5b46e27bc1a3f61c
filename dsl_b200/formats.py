"""The reference's on-disk formats either side of the teacher -> student hand-over, so a run can move between the
reference and dsl_b200 in either direction (host-only helpers; the arithmetic stays on the device):

* `adathres.json` — what UnlabelPredHook's adathres() writes once per epoch (mmdet/runner/hooks/unlabel_pred_hook.py:
  344-367) and SemiCOCODataset._parse_ann_info re-reads per box (mmdet/datasets/semicoco.py:243-258): three maps over the
  classes that were COUNTED this epoch — "cat" {category name: class weight} sorted by name, "id" {class id: class
  weight}, "thres" {category name: ignore threshold}. Classes that were never counted are absent (the dataset then
  uses its default_thres pair; the next adathres() counts every box of such a class).
* per-image pseudo-label files — what save_results2file writes (unlabel_pred_hook.py:164-175) and the dataset loads:
  {"imageName", "targetNum", "rects" [[x1, y1, x2, y2]], "tags" [category name], "masks" [[]], "scores"}.
"""
import json
import math


def adathres_to_json(thr, weight, cat_names):
    """Per-class thresholds / class weights (TeacherPost.adathres_update: weight == 0 marks a class that was not counted)
    -> the dict adathres() dumps."""
    thr = [float(v) for v in thr]
    weight = [float(v) for v in weight]
    assert len(thr) == len(weight) and len(cat_names) >= len(thr)
    counted = [c for c in range(len(thr)) if weight[c] > 0.0 and not math.isnan(weight[c])]
    by_name = sorted(counted, key=lambda c: cat_names[c])
    return {"cat": {cat_names[c]: weight[c] for c in by_name},
            "id": {int(c): weight[c] for c in by_name},
            "thres": {cat_names[c]: thr[c] for c in counted}}


def adathres_from_json(obj, cat_names, absent_thr=0.3):
    """The reference's adathres.json (dict or path) -> (thr per class id, counted mask): classes absent from "thres" get
    `absent_thr` (SemiCOCODataset's default_thres[1]) and counted False, i.e. no history for the next epoch's gate."""
    if isinstance(obj, str):
        with open(obj, "r") as f:
            obj = json.load(f)
    thres = obj["thres"]
    unknown = [k for k in thres if k not in cat_names]
    if unknown:
        raise KeyError(f"adathres.json names categories the label map does not have: {unknown}")
    thr = [float(thres.get(n, absent_thr)) for n in cat_names]
    return thr, [n in thres for n in cat_names]


def save_adathres(path, thr, weight, cat_names):
    with open(path, "w") as f:
        json.dump(adathres_to_json(thr, weight, cat_names), f, indent=4, ensure_ascii=False)


def pseudo_label_record(image_name, boxes, scores, labels, cat_names):
    """One image's pseudo boxes (as UnlabelPredHook keeps them: integer-valued corners, class-ascending, score-descending)
    -> the dict save_results2file dumps."""
    boxes = [[float(v) for v in b] for b in boxes]
    return {"imageName": image_name, "targetNum": len(boxes), "rects": boxes,
            "tags": [cat_names[int(c)] for c in labels], "masks": [[] for _ in boxes],
            "scores": [float(s) for s in scores]}


def read_pseudo_label_record(obj, cat_names):
    """A per-image pseudo-label file (dict or path) -> (rects [[x1, y1, x2, y2]], scores or None for the score-less
    labeled files, class ids) over its first targetNum entries (semicoco.py:220-224); a tag the label map does not have
    raises KeyError like the dataset's own cat2id lookup."""
    if isinstance(obj, str):
        with open(obj, "r") as f:
            obj = json.load(f)
    ids = {n: i for i, n in enumerate(cat_names)}
    n = int(obj["targetNum"])
    rects = [[float(v) for v in obj["rects"][i]] for i in range(n)]
    scores = [float(obj["scores"][i]) for i in range(n)] if "scores" in obj else None
    return rects, scores, [ids[obj["tags"][i]] for i in range(n)]
