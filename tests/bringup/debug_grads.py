#!/usr/bin/env python
"""TEST INFRASTRUCTURE (run by hand on a GPU box): per-parameter and per-FPN-level gradient errors of the CUDA backward vs
torch autograd on the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

from dsl_b200.engine import FCOSNet
from oracle import fcos_oracle as O
from tests.golden import inputs as GI
from tests.test_gpu_parity import _oracle_state, _run_loss, _nchw

B, H, W = 2, 256, 320
net = FCOSNet(B, H, W, depth=50, train=True, seed=3, loss_weight=3.0, parity_outputs=True)
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
for t in list(cs) + list(ps):
    t.retain_grad()
mode = os.environ.get("ORACLE_HEAD_INPUT", "oracle")
cls, box, ctr = O.fcos_head_forward(head, ps, training=True)
out = O.fcos_loss(cls, box, ctr, gts, labels, ignores, loss_weight=3.0)
sum(out.values()).backward()


def rl2(got, ref):
    den = ref.norm().item()
    return (got - ref).norm().item() / (den + 1e-30), den


for l in range(5):
    e, d = rl2(_nchw(net.dp[l], 256), ps[l].grad)
    print(f"dP{l + 3}: rel L2 {e:.4f} (|ref| {d:.3e})")
    gc = cls[l].grad if cls[l].grad is not None else None
for prefix, dct in (("bbox_head.", head), ("neck.", neck), ("backbone.", bb)):
    for k, v in dct.items():
        name = prefix + k
        o, n = net.store.offsets.get(name, (None, None))
        if o is None or o + n > net.store.n_train:
            continue
        ref = v.grad if v.grad is not None else torch.zeros_like(v)
        got = net.grad[o:o + n].view(ref.shape).cpu()
        e, d = rl2(got, ref)
        cos = torch.nn.functional.cosine_similarity(got.flatten().double(), ref.flatten().double(), dim=0).item()
        print(f"{name:50s} relL2 {e:.4f} cos {cos:.5f} |ref| {d:.3e} |got| {got.norm().item():.3e}")
