#!/usr/bin/env python
"""2-rank check of the bucketed gradient all-reduce (run under torchrun --nproc-per-node 2): one step from the same
initial state with the three asynchronous bucket all-reduces vs one all-reduce of the whole gradient; the updated
weights must agree (up to the fp32 atomics order of the split-K weight gradients) and be identical on both ranks."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from dsl_b200.trainer import DSLEngine
from tests.golden import inputs as GI

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, H, W = 2, 256, 320
res = {}
for graphs in (False, True, "one-graph"):
    for bucketed in (True, False):
        if graphs == "one-graph" and not bucketed:
            continue
        eng = DSLEngine(B, H, W, depth=50, seed=0, use_graphs=bool(graphs))
        eng.bucketed = bucketed
        eng.graph_nccl = graphs == "one-graph"     # collectives captured inside ONE graph vs six graphs around them
        rng = np.random.RandomState(10 + rank)
        img = GI.make_tensor(rng, B, 3, H, W, scale=50.0)
        gts, labels, ignores = GI.make_gt(20 + rank, B, H, W, with_ignore=True)
        eng.set_inputs(img, gts, labels, ignores, teacher_img=img)
        for _ in range(2):
            losses = eng.step()
        torch.cuda.synchronize()
        flat = eng.student.store.flat.clone()
        other = [torch.empty_like(flat) for _ in range(2)]
        dist.all_gather(other, flat)
        assert torch.equal(other[0], other[1]), "ranks diverged"
        res[(graphs, bucketed)] = (flat, {k: float(v) for k, v in losses.items()})
        del eng
        torch.cuda.empty_cache()
for graphs in (False, True, "one-graph"):
    a, la = res[(graphs, True)]
    b, lb = res[(True if graphs == "one-graph" else graphs, False)]
    d = (a - b).abs().max().item()
    ref = b.abs().max().item()
    if rank == 0:
        print(f"graphs={graphs}: max |w_bucketed - w_single| = {d:.3e} (max |w| {ref:.3f}); losses {la} vs {lb}")
    assert d <= 1e-5 * ref, d
if rank == 0:
    print("bucketed all-reduce OK")

# A rank-local graph capture must not talk to the peers (multi-scale training: ranks meet new padded shapes at different
# iterations): rank 1 builds + captures the plan of a NEW shape and steps it while rank 0 replays its cached graph. Every
# rank issues exactly one step's collectives; the weights must stay identical across ranks.
engA = DSLEngine(B, H, W, depth=50, seed=0, use_graphs=True)
rng = np.random.RandomState(10 + rank)
img = GI.make_tensor(rng, B, 3, H, W, scale=50.0)
gts, labels, ignores = GI.make_gt(20 + rank, B, H, W, with_ignore=True)
engA.set_inputs(img, gts, labels, ignores, teacher_img=img)
engA.step()
torch.cuda.synchronize()
if rank == 1:
    H2, W2 = 192, 256
    engB = DSLEngine(B, H2, W2, depth=50, seed=0, use_graphs=True, student_store=engA.student.store,
                     teacher_store=engA.teacher.store)
    engB.mom, engB.lr_scale = engA.mom, engA.lr_scale
    img2 = GI.make_tensor(rng, B, 3, H2, W2, scale=50.0)
    g2, l2, i2 = GI.make_gt(30, B, H2, W2, with_ignore=True)
    engB.set_inputs(img2, g2, l2, i2, teacher_img=img2)
    engB.step()          # capture (collective-free warm-up) + first replay
else:
    engA.step()          # plain replay
torch.cuda.synchronize()
flat = engA.student.store.flat.clone()
other = [torch.empty_like(flat) for _ in range(2)]
dist.all_gather(other, flat)
assert torch.equal(other[0], other[1]), "ranks diverged when only one of them captured a new shape"
engA.step() if rank == 0 else engB.step()
torch.cuda.synchronize()
if rank == 0:
    print("rank-local capture OK")
from dsl_b200 import dist_ops as _D  # noqa: E402
_D.shutdown(engA, engB if rank == 1 else None)
