"""dsl_b200/formats.py (the reference's on-disk formats around the teacher -> student hand-over) on CPU: the
adathres.json writer reproduces the file the reference's own adathres() wrote (golden misc.npz, JSON kept verbatim), the
reader inverts it, and — where the reference tree is present — its adathres() and SemiCOCODataset._parse_ann_info consume
the per-image files written here exactly like the ones its own hook writes."""
import json
import os

import numpy as np
import pytest

from dsl_b200 import formats as FM
from oracle import fcos_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")
CATS = [f"cat{i}" for i in range(6)]


def _arrays(thres, weights, C=6):
    thr, wgt = [0.3] * C, [0.0] * C                      # TeacherPost's convention for classes that were not counted
    for c, v in thres.items():
        thr[c] = v
        wgt[c] = weights[c]
    return thr, wgt


def _same(a, b, tol=1e-12):
    assert set(a) == set(b) and all(set(a[k]) == set(b[k]) for k in a), (a, b)
    for k in a:
        for kk in a[k]:
            assert abs(a[k][kk] - b[k][kk]) <= tol * max(1.0, abs(b[k][kk])), (k, kk, a[k][kk], b[k][kk])


def test_adathres_json_matches_the_file_the_reference_wrote():
    g = np.load(os.path.join(G, "misc.npz"))
    scores = {CATS.index(c): v for c, v in json.loads(bytes(g["ada_scores_json"]).decode()).items()}
    first = json.loads(bytes(g["ada_first_json"]).decode())
    second = json.loads(bytes(g["ada_second_json"]).decode())
    thres, weights = O.adathres(scores)
    ours = json.loads(json.dumps(FM.adathres_to_json(*_arrays(thres, weights), CATS)))     # through JSON: "id" keys -> str
    _same(ours, first)
    assert list(ours["cat"]) == list(first["cat"])                 # "cat" / "id" are written sorted by category name
    assert list(ours["id"]) == list(first["id"])
    thres2, weights2 = O.adathres(scores, prev_thres=thres)
    _same(json.loads(json.dumps(FM.adathres_to_json(*_arrays(thres2, weights2), CATS))), second)
    # reader: the reference's file -> per-class vector + counted mask, absent classes on the dataset's default
    thr, counted = FM.adathres_from_json(first, CATS, absent_thr=0.3)
    for c, n in enumerate(CATS):
        assert counted[c] == (n in first["thres"]) and thr[c] == first["thres"].get(n, 0.3)
    with pytest.raises(KeyError):
        FM.adathres_from_json({"thres": {"zebra": 0.31}}, CATS)


def test_adathres_json_leaves_uncounted_classes_out(tmp_path):
    thr = [0.31, 0.3, 0.35, 0.3]
    wgt = [1.2, 0.0, 0.7, float("nan")]
    names = ["d", "c", "b", "a"]
    p = str(tmp_path / "adathres.json")
    FM.save_adathres(p, thr, wgt, names)
    d = json.load(open(p))
    assert d == {"cat": {"b": 0.7, "d": 1.2}, "id": {"2": 0.7, "0": 1.2}, "thres": {"d": 0.31, "b": 0.35}}
    assert list(d["cat"]) == ["b", "d"]
    assert FM.adathres_from_json(p, names, absent_thr=0.3) == ([0.31, 0.3, 0.35, 0.3], [True, False, True, False])


def test_pseudo_label_record_round_trip():
    boxes = np.array([[1, 2, 30, 40], [5, 6, 70, 80]], dtype=np.float32)
    rec = FM.pseudo_label_record("sub/a.jpg", boxes, np.array([0.9, 0.25], np.float32), [3, 0], CATS)
    assert rec["targetNum"] == 2 and rec["tags"] == ["cat3", "cat0"] and rec["masks"] == [[], []]
    rec = json.loads(json.dumps(rec))
    rects, scores, cls = FM.read_pseudo_label_record(rec, CATS)
    assert rects == boxes.tolist() and cls == [3, 0] and scores == [float(np.float32(0.9)), 0.25]
    labeled = dict(rec)
    del labeled["scores"]                                         # the labeled set's files carry no scores
    assert FM.read_pseudo_label_record(labeled, CATS)[1] is None
    rec["tags"][1] = "zebra"
    with pytest.raises(KeyError):
        FM.read_pseudo_label_record(rec, CATS)


def test_reference_consumes_the_files_written_here(tmp_path):
    """Live: per-image files from pseudo_label_record -> the reference's own adathres() (twice: without and with the
    history file) -> same JSON as formats.adathres_to_json over the oracle's statistics; the reference's
    SemiCOCODataset._parse_ann_info reads the same files + the adathres.json written HERE and returns the oracle's
    GT / ignore split."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference tree not present (GPU box): covered by misc.npz")
    import types
    from oracle.gen_golden import _extract_method
    _, ref_adathres = ref_loader.load_hook_functions()
    cat2id = {c: i for i, c in enumerate(CATS)}
    id2cat = {str(i): c for i, c in enumerate(CATS)}
    rng = np.random.RandomState(123)
    os.makedirs(tmp_path / "sub")
    files, by_class, recs = [], {}, []
    for k in range(10):
        n = int(rng.randint(0, 30))
        x1, y1 = rng.randint(0, 500, n), rng.randint(0, 380, n)
        boxes = np.stack([x1, y1, x1 + rng.randint(0, 140, n), y1 + rng.randint(0, 100, n)], 1).astype(np.float32)
        scores = np.round(rng.rand(n) * 0.9 + 0.1, 6)
        labels = rng.choice(5, size=n, p=[0.5, 0.2, 0.15, 0.1, 0.05])          # class 5 never appears
        rec = FM.pseudo_label_record(f"sub/im{k}.jpg", boxes, scores, labels, CATS)
        json.dump(rec, open(tmp_path / "sub" / f"im{k}.jpg.json", "w"))
        files.append(f"sub/im{k}.jpg\n")
        recs.append(rec)
        for c, s in zip(labels, scores):
            by_class.setdefault(int(c), []).append(float(s))
    ref_file = str(tmp_path / "adathres.json")
    prev = None
    for _ in range(2):
        ref_adathres(0, True, ref_file, id2cat, cat2id, files, str(tmp_path / "sub"), {})
        thres, weights = O.adathres(by_class, prev_thres=prev)
        ours = json.loads(json.dumps(FM.adathres_to_json(*_arrays(thres, weights), CATS)))
        _same(ours, json.load(open(ref_file)))
        assert "cat5" not in ours["thres"]
        prev = thres
    # the dataset side: our adathres.json + our per-image files through the reference's _parse_ann_info
    mine = str(tmp_path / "mine.json")
    FM.save_adathres(mine, *_arrays(thres, weights), CATS)
    parse_ann = _extract_method(os.path.join(ref_loader.REF_ROOT, "mmdet/datasets/semicoco.py"), "SemiCOCODataset",
                                "_parse_ann_info", {"os": os, "json": json, "np": np})
    thr_vec, _ = FM.adathres_from_json(mine, CATS, absent_thr=0.3)
    for k, rec in enumerate(recs):
        slf = types.SimpleNamespace(ann_path=str(tmp_path / "sub"), thres=mine, default_thres=[0.1, 0.3],
                                    thres_list_by_class={}, labelmapper=dict(cat2id=cat2id))
        ann = parse_ann(slf, dict(filename=f"im{k}.jpg", width=640, height=480), None)
        rects, scores, cls = FM.read_pseudo_label_record(rec, CATS)
        gt, lab, ig = O.filter_pseudo_labels(rects, scores, cls, 640, 480, {c: t for c, t in enumerate(thr_vec)},
                                             (0.1, 0.3))
        assert np.array_equal(ann["bboxes"], np.asarray(gt, np.float32).reshape(-1, 4)), k
        assert np.array_equal(ann["labels"], np.asarray(lab, np.int64)), k
        assert np.array_equal(ann["bboxes_ignore"], np.asarray(ig, np.float32).reshape(-1, 4)), k


def test_pseudo_label_files_match_the_files_the_reference_hook_wrote():
    """Golden saved_files.npz = the per-image JSON files UnlabelPredHook.save_results2file wrote (verbatim) for the
    hook_chain.npz detections: the oracle's hook_saved_boxes + formats.pseudo_label_record reproduce them exactly
    (same boxes, same float scores, same order, same keys), and the reader inverts the file."""
    g = np.load(os.path.join(G, "saved_files.npz"))
    ncase, C = int(g["meta"][0]), int(g["meta"][1])
    total = 0
    for k in range(ncase):
        ref = json.loads(bytes(g[f"c{k}_saved_json"]).decode())
        rects, scores, cls = O.hook_saved_boxes(g[f"c{k}_dets"], g[f"c{k}_labels"], C)
        rec = FM.pseudo_label_record("sub/a.jpg", rects, scores, cls, CATS)
        assert json.loads(json.dumps(rec)) == ref, k
        assert list(rec) == list(ref)                               # key order of the file
        r2, s2, c2 = FM.read_pseudo_label_record(ref, CATS)
        assert (r2, s2, c2) == ([[float(v) for v in b] for b in rects], [float(s) for s in scores], list(cls)), k
        total += ref["targetNum"]
    assert total > 50
