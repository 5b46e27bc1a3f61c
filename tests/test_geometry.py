"""View geometry (Resize -> PatchShuffle -> RandomFlip box mapping, padded batch assembly): the oracle restatement and
the CUDA kernels against golden vectors produced by the reference's own pipeline classes
(oracle/gen_golden.py::gen_view_geometry)."""
import os

import numpy as np
import pytest
import torch

G = os.path.join(os.path.dirname(__file__), "golden")


def _cases():
    g = np.load(os.path.join(G, "view_geometry.npz"))
    return g, int(g["meta"][0])


def test_oracle_view_boxes_matches_reference_golden():
    from oracle import fcos_oracle as O
    g, n = _cases()
    splits = 0
    for k in range(n):
        sx, sy, w, h, clip, mode, crop, flip = g["views"][k]
        kw = dict(sx=np.float32(sx), sy=np.float32(sy), img_w=int(w), img_h=int(h), clip=bool(clip), ps_mode=int(mode),
                  ps_crop=int(crop), flip=bool(flip))
        b, l = O.view_boxes(g[f"c{k}_boxes"], g[f"c{k}_labels"], **kw)
        assert np.array_equal(b, g[f"c{k}_out_boxes"]), k          # bit-exact float32
        assert np.array_equal(l, g[f"c{k}_out_labels"]), k
        bi, _ = O.view_boxes(g[f"c{k}_ignore"], None, **kw)
        assert np.array_equal(bi, g[f"c{k}_out_ignore"]), k
        splits += len(b) - len(g[f"c{k}_boxes"])
    assert splits > 20    # the golden set really exercises the box split at the PatchShuffle cut


@pytest.mark.gpu
def test_view_boxes_kernel_matches_reference_golden():
    from dsl_b200.geometry import View, ViewGeometry
    g, n = _cases()
    geo = ViewGeometry(n, max_boxes=1024)
    geo.set_views([View(float(np.float32(v[0])), float(np.float32(v[1])), int(v[2]), int(v[3]), int(v[4]), int(v[5]),
                        int(v[6]), int(v[7])) for v in g["views"]])
    for key, with_labels in (("boxes", True), ("ignore", False)):
        lens = [len(g[f"c{k}_{key}"]) for k in range(n)]
        off = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
        boxes = torch.from_numpy(np.concatenate([g[f"c{k}_{key}"].reshape(-1, 4) for k in range(n)])).cuda()
        labels = torch.from_numpy(np.concatenate([g[f"c{k}_labels"] for k in range(n)])).cuda() if with_labels else None
        ob, ol, oo = geo.run(boxes, labels, off)
        torch.cuda.synchronize()
        oo = oo.cpu().tolist()
        for k in range(n):
            want = g[f"c{k}_out_{key}"].reshape(-1, 4)
            got = ob[oo[k]:oo[k + 1]].cpu().numpy()
            assert np.array_equal(got, want), (key, k)             # bit-exact, same order (split boxes adjacent)
            if with_labels:
                assert np.array_equal(ol[oo[k]:oo[k + 1]].cpu().numpy(), g[f"c{k}_out_labels"]), k


@pytest.mark.gpu
def test_pad_batch_matches_torch():
    from dsl_b200.geometry import pad_batch
    torch.manual_seed(0)
    imgs = [torch.randn(3, 70, 101, device="cuda"), torch.randn(3, 96, 64, device="cuda"), torch.randn(3, 33, 128, device="cuda")]
    out = pad_batch(imgs)
    assert tuple(out.shape) == (3, 3, 96, 128)
    ref = torch.zeros_like(out)
    for b, im in enumerate(imgs):
        ref[b, :, :im.shape[1], :im.shape[2]] = im
    assert torch.equal(out, ref)
