"""TEST / BENCH INFRASTRUCTURE ONLY — one DSL teacher+student step on the HOST cores with the oracle restatement
(plain torch fp32 + autograd), i.e. the reference's CPU arithmetic for the path: teacher forward (no_grad, eval) +
decode gate, student forward + FCOSHead.loss + backward, clip-grad-norm 35 + momentum SGD (bias lr x2 / wd 0) and
the EMA body of SemiEpochBasedRunner.EMA. Used by bench.py for `cpu_baseline` and `--impl reference` (kind "port":
the reference's own Python cannot travel to the GPU box — no mmcv there — see DESIGN.md)."""
import contextlib
import time

import numpy as np
import torch

from oracle import fcos_oracle as O


def _split(sd):
    bb = {k[len("backbone."):]: v for k, v in sd.items() if k.startswith("backbone.")}
    neck = {k[len("neck."):]: v for k, v in sd.items() if k.startswith("neck.")}
    head = {k[len("bbox_head."):]: v for k, v in sd.items() if k.startswith("bbox_head.")}
    return bb, neck, head


class CpuStep:
    def __init__(self, B, H, W, depth=50, seed=0, threads=None, backbone="resnet", device="cpu"):
        """device "cpu": the reference's CPU arithmetic (cpu_baseline / --impl reference). device "cuda": the SAME
        restatement run by stock torch eager + cuDNN on the GPU (bench.py --impl eager-gpu; SURVEY section 8(d)'s
        honest GPU comparator) — still only a baseline, never part of the product."""
        from dsl_b200.params import RESNET_BLOCKS, ParamStore, fpn_spec, head_spec, resnet_spec, rla_resnet_spec
        if threads:
            torch.set_num_threads(threads)
        self.B, self.H, self.W, self.depth = B, H, W, depth
        self.layers = RESNET_BLOCKS[depth]
        self.backbone = backbone     # "rla": RLA_ResNet, the shipped configs' backbone (oracle: rla_resnet_forward)
        bb_spec = rla_resnet_spec(self.layers) if backbone == "rla" else resnet_spec(depth)
        store = ParamStore(bb_spec + fpn_spec() + head_spec(), "cpu").init_reference(seed)
        self.spec = store.spec
        self.student = {p.name: store.views[p.name].clone() for p in store.spec}
        self.teacher = {k: v.clone() for k, v in self.student.items()}
        self.train_names = [p.name for p in store.spec if p.region in ("A", "B")]
        self.bias_names = {p.name for p in store.spec if p.region == "B"}
        self.mom = {n: torch.zeros_like(self.student[n]) for n in self.train_names}
        rng = np.random.RandomState(seed)
        self.img_s = torch.from_numpy((rng.randn(B, 3, H, W) * 50).astype(np.float32))
        self.img_t = torch.from_numpy((rng.randn(B, 3, H, W) * 50).astype(np.float32))
        from tests.golden import inputs as GI
        self.gts, self.labels, self.ignores = GI.make_gt(seed + 1, B, H, W, with_ignore=True)
        self.device = torch.device(device)
        self.autocast = None         # e.g. torch.bfloat16: network forwards under torch.autocast, losses / decode in fp32
        if self.device.type != "cpu":
            mv = lambda t: t.to(self.device)  # noqa: E731
            self.student = {k: mv(v) for k, v in self.student.items()}
            self.teacher = {k: mv(v) for k, v in self.teacher.items()}
            self.mom = {k: mv(v) for k, v in self.mom.items()}
            self.img_s, self.img_t = mv(self.img_s), mv(self.img_t)
            self.gts, self.labels = [mv(t) for t in self.gts], [mv(t) for t in self.labels]
            self.ignores = [mv(t) for t in self.ignores]

    def _backbone(self, bb, img):
        if self.backbone == "rla":
            return O.rla_resnet_forward(bb, img, self.layers)
        return O.resnet_forward(bb, img, self.depth)

    def _amp(self):
        if self.autocast is None:
            return contextlib.nullcontext()
        return torch.autocast(self.device.type, dtype=self.autocast)

    def step(self, lr=0.01):
        B = self.B
        f32 = lambda xs: [x.float() for x in xs]  # noqa: E731  (no-op without autocast)
        with torch.no_grad():
            bb, neck, head = _split(self.teacher)
            with self._amp():
                ps = O.fpn_forward(neck, self._backbone(bb, self.img_t))
                cls, box, ctr = O.fcos_head_forward(head, ps, training=False)
            cls, box, ctr = f32(cls), f32(box), f32(ctr)
            O.decode_candidates(cls, box, ctr, [(self.H, self.W, 3)] * B, [[1.0] * 4] * B, nms_pre=1000,
                                score_thr=0.05, rescale=True)
        params = {n: self.student[n].detach().requires_grad_(True) for n in self.train_names}
        sd = dict(self.student)
        sd.update(params)
        bb, neck, head = _split(sd)
        with self._amp():
            ps = O.fpn_forward(neck, self._backbone(bb, self.img_s))
            cls, box, ctr = O.fcos_head_forward(head, ps, training=True)
        cls, box, ctr = f32(cls), f32(box), f32(ctr)
        losses = O.fcos_loss(cls, box, ctr, self.gts, self.labels, self.ignores, loss_weight=3.0)
        sum(losses.values()).backward()
        with torch.no_grad():
            grads = [params[n].grad if params[n].grad is not None else torch.zeros_like(params[n])
                     for n in self.train_names]
            grads, _ = O.clip_grad_norm(grads, 35.0)
            for n, g in zip(self.train_names, grads):
                bias = n in self.bias_names
                p, m = O.sgd_momentum_step(self.student[n], g, self.mom[n], lr * (2.0 if bias else 1.0), 0.9,
                                           0.0 if bias else 1e-4)
                self.student[n], self.mom[n] = p, m
            self.teacher = O.ema_update(self.teacher, self.student, 0.99)
        return {k: float(v.detach()) for k, v in losses.items()}


def time_cpu_steps(B, H, W, steps=1, warmup=0, depth=50, threads=None, backbone="resnet"):
    cs = CpuStep(B, H, W, depth=depth, threads=threads, backbone=backbone)
    for _ in range(warmup):
        cs.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        cs.step()
    dt = time.perf_counter() - t0
    return dict(seconds_per_step=dt / steps, images_per_sec=B * steps / dt, cores=torch.get_num_threads())
