"""One DSL teacher+student training step on one GPU (one process per GPU; pure data parallel).

Step = { EMA-teacher forward (no_grad, eval mode: bbox x stride) on the weak-aug images + decode / score gate,
         student forward + loss + backward on the strong-aug images,
         [gradient all-reduce over ranks],  clip-grad-norm 35 + momentum SGD (bias lr x2 / wd 0),  EMA update,
         refresh of the bf16 operand caches }
which is what the reference spreads over SemiEpochBasedRunner.train / run_iter (mmdet/runner/hooks/
semi_epoch_based_runner.py:169-283), OptimizerHook, EMAOWNHook -> runner.EMA (:368-409) and UnlabelPredHook's
per-iteration teacher inference (mmdet/runner/hooks/unlabel_pred_hook.py:512-562) — minus the JSON / disk round trip.

With world_size == 1 the whole step is captured once into a CUDA graph and replayed; with more ranks it is split into
three graphs around the two collectives (the packed 2-scalar normaliser all-reduce and the gradient all-reduce).
"""
import os

import torch
import torch.distributed as dist

from . import _lib as L
from . import dist_ops
from .engine import FCOSNet, STRIDES
from .postprocess import TeacherPost


_COMM_WARM = False   # the process group's NCCL communicator exists (its creation cannot be captured into a graph)


class DSLEngine:
    def __init__(self, B, H, W, depth=50, num_classes=80, device="cuda", seed=0, lr=0.01, momentum=0.9,
                 weight_decay=1e-4, bias_lr_mult=2.0, bias_decay_mult=0.0, max_grad_norm=35.0, ema_keep=0.99,
                 loss_weight=3.0, teacher_B=None, use_graphs=True, nms_pre=1000, score_thr=0.05, two_streams=True,
                 student_store=None, teacher_store=None, scale_invariant=False, soft_weight=0.0, soft_warm_up=0,
                 head_kwargs=None, backbone="resnet"):
        """student_store / teacher_store: existing ParamStores (e.g. of two plugin.FCOS modules) to train in place.
        scale_invariant: the student batch gets the reference's extra half-resolution copy of its last image
        (semi_epoch_based_runner.py:186-204) -> B + 1 images, and the SI soft loss (fcos_head.py:312-333) with
        soft_weight (soft_weight / 1000 for the first soft_warm_up + 1 steps)."""
        self.dev = torch.device(device)
        self.B, self.H, self.W = B, H, W
        self.scale_invariant = bool(scale_invariant)
        self.soft_weight, self.soft_warm_up, self.cur_iter = float(soft_weight), int(soft_warm_up), 0
        sB = B + 1 if self.scale_invariant else B
        hk = dict(head_kwargs or {})
        hk["backbone"] = backbone   # "rla": RLA_ResNet (configs/fcos_semi/RLA_*.py:3-13), see engine_rla.py
        self.student = FCOSNet(sB, H, W, depth, num_classes, train=True, device=device, seed=seed,
                               loss_weight=loss_weight, soft_weight=soft_weight, store=student_store, **hk)
        tB = teacher_B or B
        self.teacher = FCOSNet(tB, H, W, depth, num_classes, train=False, device=device, seed=seed,
                               store=teacher_store, **hk)
        if teacher_store is None:
            # teacher starts as a copy of the student (reference: both built from the same config, load_checkpoint
            # loads the same file into both, semi_epoch_based_runner.py:350-366)
            self.teacher.store.flat.copy_(self.student.store.flat)
            self.teacher.repack()
        st = self.student.store
        self.mom = torch.zeros(st.n_train, dtype=torch.float32, device=self.dev)
        self.lr, self.momentum, self.wd = lr, momentum, weight_decay
        self.bias_lr_mult, self.bias_decay_mult = bias_lr_mult, bias_decay_mult
        self.max_grad_norm = max_grad_norm
        self.ema_keep = ema_keep
        self.ema_in_step = True   # False: the EMA is applied by an explicit DSLEngine.ema() call (EMAOWNHook epoch mode)
        self.sqnorm = torch.zeros(1, dtype=torch.float64, device=self.dev)
        self.coef = torch.ones(2, dtype=torch.float32, device=self.dev)
        self.lr_scale = torch.ones(1, dtype=torch.float32, device=self.dev)
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.student.world_size = float(self.world)
        self.use_graphs = use_graphs
        self.graphs = None
        self.nms_pre, self.score_thr = nms_pre, score_thr
        self.adathres_stats = True   # accumulate the adaptive-threshold statistics in every teacher pass
        # world > 1: all-reduce the gradient in buckets as the backward produces them (DSLB_NO_BUCKETS=1: one all-reduce)
        self.bucketed = os.environ.get("DSLB_NO_BUCKETS") is None
        # world > 1: capture the collectives INSIDE one CUDA graph (NCCL is capturable): the step is then ONE graph
        # launch instead of six graphs with eager ncclAllReduce launches between them; the bucket all-reduces run on a
        # communication stream forked off the backward, so they still overlap the later buckets. DSLB_GRAPH_NCCL=0: off.
        self.graph_nccl = os.environ.get("DSLB_GRAPH_NCCL", "1") != "0"
        self.s_comm = torch.cuda.Stream()
        self._comm_evs = None
        # |g|^2 bucket by bucket as the gradients become final (one GPU: on the weight-gradient side stream behind each
        # bucket's unpack; several: on the communication stream behind each bucket's all-reduce) instead of one pass over
        # all 128 MB between the backward and clip -> SGD. DSLB_BUCKET_SQNORM=0: off.
        self.bucket_sqnorm = os.environ.get("DSLB_BUCKET_SQNORM", "1") != "0"
        self._sq_partial = False
        # several GPUs, one-graph step: the dgrad stream does not join the weight-gradient stream at bucket boundaries —
        # only the communication stream waits for a bucket to be final. DSLB_BUCKET_JOIN=1: join as before.
        self.lazy_bucket_join = os.environ.get("DSLB_BUCKET_JOIN", "0") != "1"
        # DSLB_PREP_UNDER_FWD=1: gradient-buffer memsets + target assignment on the weight-gradient stream under the
        # forward pass instead of at the head of the backward. Opt-in: measured slower on B200 (8.78 vs 8.70 ms over three
        # interleaved runs) — the 256 MB of memset writes compete with the HBM-bound layer1 convs they overlap.
        # DSLB_PREP_UNDER_FWD=towers: the same, forked where the FCOSHead towers start (tensor-bound launches, HBM idle).
        # DSLB_PREP_UNDER_FWD=loss: only the two memsets, on the weight-gradient stream BESIDE the loss kernel (which reads
        # and writes 63 MB and none of those buffers) instead of in front of the backward.
        _m = os.environ.get("DSLB_PREP_UNDER_FWD", "0")
        self.prep_mode = {"1": "start", "start": "start", "towers": "towers", "loss": "loss"}.get(_m, "off")
        self.prep_under_forward = self.prep_mode in ("start", "towers")
        self.student.zero_in_bwd = not (self.prep_under_forward or (self.prep_mode == "loss" and two_streams))
        self._prep_ev0, self._prep_ev1 = torch.cuda.Event(), torch.cuda.Event()
        self.fuse_ema = os.environ.get("DSLB_FUSE_EMA", "1") != "0"   # EMA of the trainable regions inside the SGD kernel
        self.student.bucket_hook = self._bucket_sqnorm if (self.world == 1 and self.bucket_sqnorm) else None
        global _COMM_WARM
        if self.world > 1 and not _COMM_WARM:
            # first engine of the process = a point every rank reaches together: force the (lazy) communicator creation
            # here, eagerly, on the stream the captured collectives will use
            with torch.cuda.stream(self.s_comm):
                dist.all_reduce(torch.zeros(1, device=self.dev))
            torch.cuda.synchronize()
            _COMM_WARM = True
        self._build_teacher_post()
        # Tried and kept as an opt-in (DSLB_JOINT_FWD=1): teacher + student forward as JOINT launches — layer k of the two
        # networks as ONE persistent conv launch with the segments of both. Measured slower (9.52-9.59 ms vs 9.37 ms, B=4 at
        # 800x1344; R101 bs 2: 7.86 vs 7.80 ms): two independent passes on two streams let one kernel's tail wave overlap the
        # other's prologue and let the small elementwise kernels co-run, which a merged launch serialises again.
        self.joint_fwd = self._build_joint_forward() if os.environ.get("DSLB_JOINT_FWD", "0") == "1" else None
        self.launches_per_step = None
        self.two_streams = two_streams
        self.s2 = torch.cuda.Stream()   # teacher branch
        self.s3 = torch.cuda.Stream()   # weight gradients of the student backward
        self._fork_ev, self._join_ev = torch.cuda.Event(), torch.cuda.Event()
        self._fork2_ev, self._join2_ev = torch.cuda.Event(), torch.cuda.Event()
        # double-buffered input staging for prefetch_inputs(): H2D of batch i+1 on a copy stream under step i
        self.copy_stream = torch.cuda.Stream()
        self._stage = None
        self._stage_ev, self._consumed_ev = torch.cuda.Event(), torch.cuda.Event()
        self._consumed_ev.record()
        self._stage_pending = False
        self._h_gt_off = torch.zeros(B + 1, dtype=torch.int32).pin_memory()
        self._h_ig_off = torch.zeros(B + 1, dtype=torch.int32).pin_memory()

    # ---------------------------------------------------------------------------------------- teacher decode
    def _build_teacher_post(self):
        t = self.teacher
        self.post = TeacherPost(t.B, t.psize, STRIDES, t.C, self.dev, nms_pre=self.nms_pre, score_thr=self.score_thr)
        self.post.set_meta([(self.H, self.W, 3)] * t.B, [[1.0] * 4] * t.B)
        self.cand_counts = self.post.cand_counts
        # detections -> pseudo GT / ignore boxes of the NEXT student batch (reference: JSON files read back by the
        # dataloader `preload` iterations later); kept in separate buffers unless feed_pseudo_labels is set
        mb = self.student.max_boxes
        self.pl_gt_boxes = torch.zeros(mb, 4, dtype=torch.float32, device=self.dev)
        self.pl_gt_labels = torch.zeros(mb, dtype=torch.int64, device=self.dev)
        self.pl_gt_off = torch.zeros(t.B + 1, dtype=torch.int32, device=self.dev)
        self.pl_ig_boxes = torch.zeros(mb, 4, dtype=torch.float32, device=self.dev)
        self.pl_ig_off = torch.zeros(t.B + 1, dtype=torch.int32, device=self.dev)

    def _build_joint_forward(self):
        from .engine import ConvPlan
        s_ops, t_ops = self.student.fwd_ops, self.teacher.fwd_ops
        if len(s_ops) != len(t_ops):
            return None
        joint, merged = [], 0
        self._joint_plans = []
        for a, b in zip(s_ops, t_ops):
            pa, pb = getattr(a, "__self__", None), getattr(b, "__self__", None)
            ok = isinstance(pa, ConvPlan) and isinstance(pb, ConvPlan) and len(pa.segs) + len(pb.segs) <= L.MAX_SEGS
            if ok and len(pa.segs) == 1:
                g = pa.segs[0]      # single narrow 3x3 stride-1 convs run on the halo-tile kernel, which takes one segment
                if g.get("R") == 3 and g.get("stride", 1) == 1 and g.get("Cin") in (64, 128) and g.get("Cout") in (64, 128):
                    ok = False
            if ok:
                plan = ConvPlan(pa.segs + pb.segs, pa.what + "+teacher")
                self._joint_plans.append(plan)
                joint.append(plan.run)
                merged += 1
            else:
                joint.append(a)
                joint.append(b)
        return joint if merged else None

    def _forward_both(self):
        """Teacher (no_grad, eval) and student forward, then the teacher's post-processing forked onto the second stream
        (nothing in this step's student pass consumes it)."""
        with torch.no_grad():
            for op in self.joint_fwd:
                op()
        if not self.two_streams:
            with torch.no_grad():
                self.teacher_decode()
            return
        main = torch.cuda.current_stream()
        self._fork_ev.record(main)
        self.s2.wait_event(self._fork_ev)
        with torch.cuda.stream(self.s2):
            with torch.no_grad():
                self.teacher_decode()
            self._join_ev.record(self.s2)

    def teacher_decode(self):
        """Teacher head outputs -> gated candidates -> NMS -> pseudo GT / ignore boxes, all on the device
        (fcos_head.py:406-548; bbox_nms.py:34-94; unlabel_pred_hook.py:20-38,142-165; semicoco.py:220-269)."""
        t = self.teacher
        self.post.decode(t.cls_out, t.rc_out)
        self.post.nms()
        self.post.pseudo_labels(self.pl_gt_boxes, self.pl_gt_labels, self.pl_gt_off, self.pl_ig_boxes, self.pl_ig_off,
                                accumulate_stats=self.adathres_stats)

    def end_epoch(self, **adathres_kw):
        """Per-epoch adaptive thresholds (UnlabelPredHook.before_train_epoch -> adathres, unlabel_pred_hook.py:447-449):
        turn the statistics the teacher branch accumulated on the device into next epoch's per-class thresholds."""
        had = self.post.have_prev
        if self.post.overflowed():
            raise RuntimeError(f"teacher decode: an image produced more than cand_cap={self.post.cand_cap} gated candidates "
                               "during this epoch; the surplus was dropped in arrival order (the reference has no cap). "
                               "Raise TeacherPost(cand_cap=...) or score_thr.")
        out = self.post.adathres_update(**adathres_kw)
        if self.post.have_prev != had:
            # the captured pseudo-label launch carries the "no history yet" gate (a NULL pointer for last epoch's
            # thresholds, unlabel_pred_hook.py:315-343): capture again now that a history exists
            self.graphs = None
        return out

    # ---------------------------------------------------------------------------------------- step pieces
    def _teacher_branch(self):
        with torch.no_grad():
            self.teacher.forward()
            self.teacher_decode()

    def _fork_teacher(self):
        """Teacher forward + post-processing on a second stream: nothing in this step's student pass consumes it (the
        reference's pseudo labels reach the student `preload` iterations later), so it only has to be done before the
        EMA overwrites the teacher weights. Tail waves and the small latency-bound post-processing kernels then overlap
        with the student's launches."""
        if not self.two_streams:
            self._teacher_branch()
            return
        main = torch.cuda.current_stream()
        self._fork_ev.record(main)
        self.s2.wait_event(self._fork_ev)
        with torch.cuda.stream(self.s2):
            self._teacher_branch()
            self._join_ev.record(self.s2)

    def _join_teacher(self):
        if self.two_streams:
            torch.cuda.current_stream().wait_event(self._join_ev)

    def _fork_prep(self, targets):
        """What the backward needs and the forward does not touch — cleared gradient buffers (two 128 MB memsets) and, with
        `targets`, the target assignment (it reads only the GT / ignore boxes) — on the weight-gradient stream, idle during
        the forward pass, instead of between the forward and the loss on the critical path."""
        if not self.prep_under_forward:
            return
        with torch.no_grad():
            if not self.two_streams:
                self.student.zero_state()
                if targets:
                    self.student.run_targets()
                return
            main = torch.cuda.current_stream()
            self._prep_ev0.record(main)
            self.s3.wait_event(self._prep_ev0)
            with torch.cuda.stream(self.s3):
                self.student.zero_state()
                if targets:
                    self.student.run_targets()
                self._prep_ev1.record(self.s3)

    def _join_prep(self, targets):
        if not self.prep_under_forward:
            if targets:
                with torch.no_grad():
                    self.student.run_targets()
            return
        if self.two_streams:
            torch.cuda.current_stream().wait_event(self._prep_ev1)

    def _student_forward(self, targets):
        """Student forward pass; prep_mode "towers" forks the backward's preparation (_fork_prep) where the head starts."""
        with torch.no_grad():
            if self.prep_mode == "towers":
                k = self.student.head_op_start
                self.student.forward(end=k)
                self._fork_prep(targets)
                self.student.forward(start=k)
            else:
                self.student.forward()

    def _phase_a(self):
        if self.prep_mode == "start" or self.joint_fwd is not None:
            self._fork_prep(targets=True)
        if self.joint_fwd is not None:
            self._forward_both()
        else:
            self._fork_teacher()
            self._student_forward(targets=True)
        self._join_prep(targets=True)
        if self.world > 1:
            self._join_teacher()   # each captured graph must re-join its forked stream

    def _phase_t(self):
        """Target assignment alone: it needs only the GT / ignore boxes, so with several ranks it runs FIRST and the
        packed normaliser all-reduce hides under the forward passes instead of sitting between forward and loss."""
        with torch.no_grad():
            self.student.run_targets()

    def _phase_a_fwd(self):
        if self.prep_mode == "start" or self.joint_fwd is not None:
            self._fork_prep(targets=False)
        if self.joint_fwd is not None:
            self._forward_both()
        else:
            self._fork_teacher()
            self._student_forward(targets=False)
        self._join_prep(targets=False)
        self._join_teacher()

    def _run_loss(self):
        """Loss + head-output gradients; prep_mode "loss": the gradient-buffer memsets run beside it on the side stream."""
        if self.prep_mode == "loss" and not self.student.zero_in_bwd and not self.two_streams:
            self.student.zero_state()      # (instrumented passes serialise the streams)
            self.student.run_loss()
        elif self.prep_mode == "loss" and not self.student.zero_in_bwd:
            main = torch.cuda.current_stream()
            self._prep_ev0.record(main)
            self.s3.wait_event(self._prep_ev0)
            with torch.cuda.stream(self.s3):
                self.student.zero_state()
                self._prep_ev1.record(self.s3)
            self.student.run_loss()
            main.wait_event(self._prep_ev1)
        else:
            self.student.run_loss()

    def _phase_b(self):
        self._run_loss()
        if self.student.bucket_hook is not None:
            L.zero(self.sqnorm)    # the buckets add their shares as they become final (_bucket_sqnorm)
            self._sq_partial = True
        self.student.backward(self.s3 if self.two_streams else None)

    def _bucket_sqnorm(self, k, lo, hi):
        """bucket_hook of a single-GPU step: the bucket's share of |g|^2, on the side stream behind its unpack, so only
        the last (smallest) bucket's share sits between the backward and clip -> SGD."""
        g = self.student.grad
        L.check(L.lib.dslb_sq_norm(L.ptr(g[lo:hi]), hi - lo, L.ptr(self.sqnorm), L.cur_stream()), "sq_norm bucket")

    def _bucket_ranges(self):
        """Backward op ranges ending at the gradient-bucket boundaries: [(op_start, op_end, grad_lo, grad_hi)]."""
        out, start = [], 0
        for end, lo, hi in self.student.bwd_buckets:
            out.append((start, end, lo, hi))
            start = end
        out[-1] = (out[-1][0], len(self.student.bwd_ops), out[-1][2], out[-1][3])
        return out

    def _phase_b_part(self, k, join=True):
        """Bucket k of the backward (k = 0 also evaluates the loss): when it returns, grad[lo:hi] of that bucket is final
        (join=False: final once the returned side-stream event has been reached, see FCOSNet.backward)."""
        start, end, _, _ = self._bucket_ranges()[k]
        if k == 0:
            self._run_loss()
        return self.student.backward(self.s3 if self.two_streams else None, start=start, end=end, join=join)

    def _phase_c(self):
        if self.world == 1:
            self._join_teacher()
        s = L.cur_stream()
        st, tt = self.student.store, self.teacher.store
        g = self.student.grad
        if not self._sq_partial:
            L.zero(self.sqnorm)
            L.check(L.lib.dslb_sq_norm(L.ptr(g), g.numel(), L.ptr(self.sqnorm), s), "sq_norm")
        self._sq_partial = False
        # max_grad_norm None = no clipping (optimizer_config.grad_clip=None): a bound no fp32 norm reaches gives coef 1
        L.check(L.lib.dslb_clip_coef(L.ptr(self.sqnorm), float(self.max_grad_norm if self.max_grad_norm is not None
                                                                 else 3.0e38), L.ptr(self.coef), s), "clip_coef")
        a0, a1 = st.region_range["A"]
        b0, b1 = st.region_range["B"]
        k = float(self.ema_keep)
        c_s = float(torch.tensor(1 - k, dtype=torch.float32))  # fp32(1 - keep_rate), as torch's scalar promotion does
        c_t = float(torch.tensor(k, dtype=torch.float32))
        # per-iteration EMA: the teacher copies of the trainable regions are updated by the SGD kernel itself (same
        # arithmetic on the new weights, one read of them less); the flat EMA then only covers what SGD never touches
        fused = self.ema_in_step and self.fuse_ema
        for (r0, r1, lr, wd, what) in ((a0, a1, self.lr, self.wd, "sgd A"),
                                       (b0, b1, self.lr * self.bias_lr_mult, self.wd * self.bias_decay_mult, "sgd B")):
            args = (L.ptr(st.flat[r0:r1]), L.ptr(g[r0:r1]), L.ptr(self.mom[r0:r1]), r1 - r0, L.ptr(self.coef),
                    L.ptr(self.lr_scale), lr, self.momentum, wd, 0)
            if fused:
                L.check(L.lib.dslb_sgd_ema_step(*args, L.ptr(tt.flat[r0:r1]), c_s, c_t, s), what)
            else:
                L.check(L.lib.dslb_sgd_step(*args, s), what)
        assert a0 == 0 and b0 == a1 and b1 == st.n_train

        def teacher_side():
            if not self.ema_in_step:
                return
            lo = st.n_train if fused else 0    # (frozen weights, BatchNorm buffers: everything behind the trainable range)
            L.check(L.lib.dslb_ema_update(L.ptr(tt.flat[lo:]), L.ptr(st.flat[lo:]), st.numel - lo, c_s, c_t,
                                          L.cur_stream()), "ema")
            self.teacher.repack(everything=True)    # the reference's EMA touches every state_dict entry

        if self.two_streams:
            # EMA + teacher operand refresh on the second stream, student operand refresh on this one: both only READ the
            # student's new weights, and each alone is latency bound rather than bandwidth bound
            main = torch.cuda.current_stream()
            self._fork2_ev.record(main)
            self.s2.wait_event(self._fork2_ev)
            with torch.cuda.stream(self.s2):
                teacher_side()
                self._join2_ev.record(self.s2)
            self.student.repack(everything=False)  # frozen stem / layer1 / BatchNorm operands never change
            main.wait_event(self._join2_ev)
        else:
            teacher_side()
            self.student.repack(everything=False)

    def ema(self, keep_rate=None):
        """runner.EMA() as an explicit call (semi_epoch_based_runner.py:368-409): T <- (1 - k) S + k T over every
        state_dict entry + refresh of the teacher's derived operands. The captured step does this itself while
        `ema_in_step` is set."""
        k = float(self.ema_keep if keep_rate is None else keep_rate)
        c_s = float(torch.tensor(1 - k, dtype=torch.float32))
        c_t = float(torch.tensor(k, dtype=torch.float32))
        st, tt = self.student.store, self.teacher.store
        L.check(L.lib.dslb_ema_update(L.ptr(tt.flat), L.ptr(st.flat), st.numel, c_s, c_t, L.cur_stream()), "ema")
        self.teacher.repack(everything=True)

    def _allreduce_counts(self):
        dist_ops.allreduce_sum_(self.student.counts)   # packed (num_pos, sum ctr-targets): one collective

    def _allreduce_grads(self):
        dist_ops.allreduce_mean_(self.student.grad)    # DDP semantics: mean over ranks (mmdet/apis/train.py:88-96)

    def _allreduce_bucket_async(self, k):
        """Mean all-reduce of gradient bucket k, asynchronous: it overlaps the backward of the later buckets (the
        reference's DDP does the same with its 25 MB buckets, mmdet/apis/train.py:88-102)."""
        _, _, lo, hi = self._bucket_ranges()[k]
        return dist_ops.allreduce_mean_async_(self.student.grad[lo:hi])

    def _run_eager(self, collectives=True):
        """One step without graphs. collectives=False: the same launches without any all-reduce — the warm-up pass of
        a graph capture. A rank captures whenever IT first meets a padded shape (multi-scale batches differ between
        ranks), so a warm-up that talked to its peers would pair its collectives with whatever the other ranks happen to
        be sending (a gradient bucket against a 2-double normaliser: hang or corrupted gradients). The captured graphs
        exclude the collectives anyway and the warm-up's state changes are rolled back."""
        if self.world > 1 and self.bucketed:
            self._phase_t()
            wc = dist_ops.allreduce_sum_async_(self.student.counts) if collectives else None
            self._phase_a_fwd()
            if wc is not None:
                wc.wait()
            works = []
            for k in range(len(self.student.bwd_buckets)):
                self._phase_b_part(k)
                if collectives:
                    works.append(self._allreduce_bucket_async(k))
            for w in works:
                if w is not None:
                    w.wait()
        else:
            self._phase_a()
            if collectives:
                self._allreduce_counts()
            self._phase_b()
            if collectives:
                self._allreduce_grads()
        self._phase_c()

    def _step_with_collectives(self):
        """The whole multi-rank step as one stream-ordered sequence (capturable): target assignment, the packed
        normaliser all-reduce on the communication stream under the forward passes, backward bucket by bucket with each
        bucket's mean all-reduce forked onto the communication stream as soon as the bucket is final, join, optimizer."""
        main = torch.cuda.current_stream()
        nb = len(self.student.bwd_buckets)
        if self._comm_evs is None:
            self._comm_evs = [torch.cuda.Event() for _ in range(2 * (nb + 1))]
        ev = self._comm_evs
        self._phase_t()
        ev[0].record(main)
        self.s_comm.wait_event(ev[0])
        with torch.cuda.stream(self.s_comm):
            dist_ops.allreduce_sum_(self.student.counts)
            ev[1].record(self.s_comm)
        if self.prep_mode == "start" or self.joint_fwd is not None:
            self._fork_prep(targets=False)
        if self.joint_fwd is not None:
            self._forward_both()
        else:
            self._fork_teacher()   # joined before the optimizer / EMA (the teacher branch overlaps the whole backward)
            self._student_forward(targets=False)
        self._join_prep(targets=False)
        main.wait_event(ev[1])
        lazy = self.lazy_bucket_join and self.two_streams
        if self.bucket_sqnorm:
            L.zero(self.sqnorm)
            self._sq_partial = True
        for k in range(nb):
            # lazy: the main (dgrad) stream does not wait for the bucket's weight gradients — only the communication
            # stream does, so the dgrad chain of the next bucket starts at once, as in the single-GPU step
            side_ev = self._phase_b_part(k, join=not lazy or k == nb - 1)
            _, _, lo, hi = self._bucket_ranges()[k]
            ev[2 + 2 * k].record(main)
            self.s_comm.wait_event(ev[2 + 2 * k])
            if side_ev is not None:
                self.s_comm.wait_event(side_ev)
            with torch.cuda.stream(self.s_comm):
                dist_ops.allreduce_mean_(self.student.grad[lo:hi])
                if self.bucket_sqnorm:   # the bucket's share of |mean gradient|^2, behind its all-reduce
                    self._bucket_sqnorm(k, lo, hi)
                ev[3 + 2 * k].record(self.s_comm)
        for k in range(nb):
            main.wait_event(ev[3 + 2 * k])
        self._join_teacher()
        self._phase_c()

    def _capture(self):
        torch.cuda.synchronize()
        # The warm-up pass torch's graph capture requires must not count as a training step: every piece of state a step
        # mutates (weights, momentum, EMA teacher, adathres statistics) is put back afterwards, so the first replay IS
        # the first step.
        state = [self.student.store.flat, self.teacher.store.flat, self.mom, self.post.stat_cnt, self.post.stat_cum]
        saved = [t.clone() for t in state]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up on a side stream, as torch's graph capture requires
            self._run_eager(collectives=False)   # no peer traffic: capture is a rank-local event (see _run_eager)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for t, v in zip(state, saved):
            t.copy_(v)
        self.student.repack(everything=True)
        self.teacher.repack(everything=True)
        torch.cuda.synchronize()
        if self.world == 1:
            phases = [[self._phase_a, self._phase_b, self._phase_c]]
        elif self.bucketed and self.graph_nccl:
            phases = [[self._step_with_collectives]]
        elif self.bucketed:
            nb = len(self.student.bwd_buckets)
            phases = [[self._phase_t], [self._phase_a_fwd]] + [[(lambda k=k: self._phase_b_part(k))] for k in range(nb)] + \
                [[self._phase_c]]
        else:
            phases = [[self._phase_a], [self._phase_b], [self._phase_c]]
        graphs = []
        pool = None
        for fns in phases:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                for f in fns:
                    f()
            pool = g.pool()
            graphs.append(g)
        self.graphs = graphs

    # ---------------------------------------------------------------------------------------- public
    def set_inputs(self, student_img, gt_bboxes, gt_labels, gt_bboxes_ignore=None, teacher_img=None):
        """Copy one batch into the engine's static input buffers (host tensors should be pinned for async H2D).
        student_img None: the images are already there (set_images_from_sources)."""
        if self.scale_invariant:
            # reference: SemiEpochBasedRunner.train builds the extra image on the host (:186-204); here the batch of B
            # lands in the first B slots and the half-resolution copy of the last one is written by a kernel
            B = self.B
            if student_img is not None:
                self.student.img[:B].copy_(student_img, non_blocking=True)
            L.check(L.lib.dslb_si_half_image(L.ptr(self.student.img[B - 1]), L.ptr(self.student.img[B]), 3, self.H,
                                             self.W, L.cur_stream()), "si_half_image")
            gt_bboxes = list(gt_bboxes) + [gt_bboxes[-1] / 2]
            gt_labels = list(gt_labels) + [gt_labels[-1]]
            if gt_bboxes_ignore is not None:
                gt_bboxes_ignore = list(gt_bboxes_ignore) + [gt_bboxes_ignore[-1] / 2]
            self._si_weight_tick()
        elif student_img is not None:
            self.student.img.copy_(student_img, non_blocking=True)
        self.student.set_targets(gt_bboxes, gt_labels, gt_bboxes_ignore)
        if teacher_img is not None:
            self.teacher.img.copy_(teacher_img, non_blocking=True)

    def _si_weight_tick(self):
        """Soft-loss warm-up (fcos_head.py:325-327): weight / 1000 while soft_warm_up >= cur_iter."""
        if self.soft_weight != 0.0:
            sw = self.soft_weight / 1000.0 if self.soft_warm_up >= self.cur_iter else self.soft_weight
            if sw != self.student.si_weight:
                self.student.si_weight = sw
                self.graphs = None      # the weight is a launch constant of the captured loss kernel
            if self.soft_warm_up >= self.cur_iter:
                self.cur_iter += 1

    def set_inputs_with_pseudo_labels(self, student_img, gt_bboxes, gt_labels, gt_bboxes_ignore, views, teacher_img=None,
                                      pl=None):
        """set_inputs() for the reference's labeled + unlabeled batch mix, with the unlabeled part labelled ON THE
        DEVICE: `gt_*` cover only the first B - tB (labeled) images; the last tB images are the strong views of the
        images the EMA teacher labelled in an earlier pass (`pl` = (gt_boxes, gt_labels, gt_off, ig_boxes, ig_off) in
        original-image coordinates, default: this engine's own self.pl_* of the previous step), carried into each
        strong view by `views` (geometry.View per unlabeled image: Resize scale, PatchShuffle cut, flip). Replaces the
        reference's JSON round trip (UnlabelPredHook.save_results2file -> SemiCOCODataset -> pipelines) for the
        transforms geometry.py covers. With scale_invariant the extra half-resolution copy of the last image and its
        halved box lists (semi_epoch_based_runner.py:186-204) are built on the device as well. No host sync."""
        from .geometry import ViewGeometry
        st, tB = self.student, self.teacher.B
        BL = self.B - tB
        assert BL >= 0 and len(gt_bboxes) == BL and len(gt_labels) == BL and len(views) == tB
        assert gt_bboxes_ignore is not None and len(gt_bboxes_ignore) == BL
        if getattr(self, "_geo", None) is None:
            self._geo = ViewGeometry(tB, max_boxes=st.max_boxes, device=self.dev)
            self._geo_off = torch.zeros(tB + 1, dtype=torch.int32, device=self.dev)
        self._geo.set_views(views)
        B = self.B
        if student_img is not None:      # None: already rendered in place by set_images_from_sources
            st.img[:B].copy_(student_img, non_blocking=True)
        if teacher_img is not None:
            self.teacher.img.copy_(teacher_img, non_blocking=True)
        pl_gt_boxes, pl_gt_labels, pl_gt_off, pl_ig_boxes, pl_ig_off = pl if pl is not None else (
            self.pl_gt_boxes, self.pl_gt_labels, self.pl_gt_off, self.pl_ig_boxes, self.pl_ig_off)
        # labeled part from the host lists
        offs, ioffs = [0], [0]
        for b in gt_bboxes:
            offs.append(offs[-1] + int(b.shape[0]))
        for b in gt_bboxes_ignore:
            ioffs.append(ioffs[-1] + int(b.shape[0]))
        nL, nI = offs[-1], ioffs[-1]
        if nL:
            st.gt_boxes[:nL].copy_(torch.cat([b.reshape(-1, 4) for b in gt_bboxes]).to(torch.float32), non_blocking=True)
            st.gt_labels[:nL].copy_(torch.cat([l.reshape(-1) for l in gt_labels]).to(torch.int64), non_blocking=True)
        if nI:
            st.ig_boxes[:nI].copy_(torch.cat([b.reshape(-1, 4) for b in gt_bboxes_ignore]).to(torch.float32),
                                   non_blocking=True)
        st.gt_off[:BL + 1].copy_(torch.tensor(offs, dtype=torch.int32), non_blocking=True)
        st.ig_off[:BL + 1].copy_(torch.tensor(ioffs, dtype=torch.int32), non_blocking=True)
        st.use_ignore = True
        # unlabeled part: teacher's pseudo GT / ignore lists -> strong views, appended behind the labeled boxes
        self._geo.run(pl_gt_boxes, pl_gt_labels, pl_gt_off, out_boxes=st.gt_boxes[nL:],
                      out_labels=st.gt_labels[nL:], out_off=self._geo_off)
        st.gt_off[BL + 1:B + 1].copy_(self._geo_off[1:] + nL)
        self._geo.run(pl_ig_boxes, None, pl_ig_off, out_boxes=st.ig_boxes[nI:], out_off=self._geo_off)
        st.ig_off[BL + 1:B + 1].copy_(self._geo_off[1:] + nI)
        if self.scale_invariant:
            s = L.cur_stream()
            L.check(L.lib.dslb_si_half_image(L.ptr(st.img[B - 1]), L.ptr(st.img[B]), 3, self.H, self.W, s), "si_half_image")
            L.check(L.lib.dslb_append_scaled_boxes(L.ptr(st.gt_boxes), L.ptr(st.gt_labels), L.ptr(st.gt_off), B, 0.5,
                                                   st.max_boxes, s), "si boxes")
            L.check(L.lib.dslb_append_scaled_boxes(L.ptr(st.ig_boxes), None, L.ptr(st.ig_off), B, 0.5, st.max_boxes, s),
                    "si ignore boxes")
            self._si_weight_tick()

    def set_images_from_sources(self, student_srcs, student_views, teacher_srcs=None, teacher_views=None,
                                mean=(123.675, 116.28, 103.53), std=(58.395, 57.12, 57.375), to_rgb=True):
        """Render the step's images straight into the plans' static input buffers from uint8 HWC source images on the
        device (cv2.imread order) and the draws of the reference's pipelines (geometry.image_view: Resize scale,
        PatchShuffle cut, flip): Resize -> PatchShuffle -> RandomFlip -> Normalize -> Pad -> collate
        (configs/fcos_semi/*.py:66-92, 108-121) as ONE launch per network, bit-exact with the cv2 / mmcv CPU pipeline.
        The host sends ~1 byte per source pixel instead of 12 bytes per padded pixel. Boxes follow through set_inputs /
        set_inputs_with_pseudo_labels (the scale-invariant extra image is built by set_inputs from slot B - 1)."""
        from .geometry import view_images
        assert len(student_srcs) == self.B
        view_images(student_srcs, student_views, mean, std, to_rgb=to_rgb, H=self.H, W=self.W, out=self.student.img[:self.B])
        if teacher_srcs is not None:
            assert len(teacher_srcs) == self.teacher.B
            view_images(teacher_srcs, teacher_views, mean, std, to_rgb=to_rgb, H=self.H, W=self.W, out=self.teacher.img)

    def prefetch_inputs(self, student_img, gt_bboxes, gt_labels, gt_bboxes_ignore=None, teacher_img=None):
        """Asynchronous set_inputs(): the pinned host batch is copied H2D on a separate copy stream into staging
        buffers (so it overlaps the step that is still running); the next step() moves it into the plan's static input
        buffers with device-side copies before launching. Call it right after step() for the NEXT batch."""
        st = self.student
        assert not self.scale_invariant, "prefetch_inputs does not build the scale-invariant extra image: use set_inputs"
        if self._stage is None:
            self._stage = dict(img_s=torch.empty_like(st.img), img_t=torch.empty_like(self.teacher.img),
                               gt_boxes=torch.empty_like(st.gt_boxes), gt_labels=torch.empty_like(st.gt_labels),
                               gt_off=torch.empty_like(st.gt_off), ig_boxes=torch.empty_like(st.ig_boxes),
                               ig_off=torch.empty_like(st.ig_off))
        sg = self._stage
        offs, ioffs = [0], [0]
        for b in gt_bboxes:
            offs.append(offs[-1] + int(b.shape[0]))
        assert offs[-1] <= st.max_boxes, "too many GT boxes for the preallocated buffer"
        use_ignore = gt_bboxes_ignore is not None
        if use_ignore:
            for b in gt_bboxes_ignore:
                ioffs.append(ioffs[-1] + int(b.shape[0]))
            assert ioffs[-1] <= st.max_boxes
        cs = self.copy_stream
        cs.wait_event(self._consumed_ev)   # the previous step has moved the staging buffers into its inputs
        with torch.cuda.stream(cs):
            sg["img_s"].copy_(student_img, non_blocking=True)
            if teacher_img is not None:
                sg["img_t"].copy_(teacher_img, non_blocking=True)
            o = 0
            for b, l in zip(gt_bboxes, gt_labels):
                n = int(b.shape[0])
                if n:
                    sg["gt_boxes"][o:o + n].copy_(b, non_blocking=True)
                    sg["gt_labels"][o:o + n].copy_(l, non_blocking=True)
                o += n
            self._h_gt_off.copy_(torch.tensor(offs, dtype=torch.int32))
            sg["gt_off"].copy_(self._h_gt_off, non_blocking=True)
            if use_ignore:
                o = 0
                for b in gt_bboxes_ignore:
                    n = int(b.shape[0])
                    if n:
                        sg["ig_boxes"][o:o + n].copy_(b, non_blocking=True)
                    o += n
                self._h_ig_off.copy_(torch.tensor(ioffs, dtype=torch.int32))
                sg["ig_off"].copy_(self._h_ig_off, non_blocking=True)
            self._stage_ev.record(cs)
        self._stage_pending = True
        self._stage_teacher = teacher_img is not None
        assert use_ignore == st.use_ignore or self.graphs is None, "ignore boxes on/off is fixed once the graph exists"
        st.use_ignore = use_ignore

    def _consume_stage(self):
        if not self._stage_pending:
            return
        st, sg = self.student, self._stage
        torch.cuda.current_stream().wait_event(self._stage_ev)
        st.img.copy_(sg["img_s"], non_blocking=True)
        if self._stage_teacher:
            self.teacher.img.copy_(sg["img_t"], non_blocking=True)
        for k in ("gt_boxes", "gt_labels", "gt_off", "ig_boxes", "ig_off"):
            getattr(st, k).copy_(sg[k], non_blocking=True)
        self._consumed_ev.record()
        self._stage_pending = False

    def step(self):
        """Run one teacher+student step on the inputs last given to set_inputs() / prefetch_inputs()."""
        self._consume_stage()
        if not self.use_graphs:
            self._run_eager()
        else:
            if self.graphs is None:
                self._capture()
            if self.world == 1 or (self.bucketed and self.graph_nccl):
                self.graphs[0].replay()
            elif self.bucketed:
                self.graphs[0].replay()
                wc = dist_ops.allreduce_sum_async_(self.student.counts)
                self.graphs[1].replay()
                if wc is not None:
                    wc.wait()
                works = []
                for k in range(len(self.student.bwd_buckets)):
                    self.graphs[2 + k].replay()
                    works.append(self._allreduce_bucket_async(k))
                for w in works:
                    if w is not None:
                        w.wait()
                self.graphs[-1].replay()
            else:
                self.graphs[0].replay()
                self._allreduce_counts()
                self.graphs[1].replay()
                self._allreduce_grads()
                self.graphs[2].replay()
        return self.student.losses()

    @staticmethod
    def split_by_bound(rows, ridge):
        """rows: (ms, flops, bytes) per implicit-GEMM launch. A launch whose arithmetic intensity (algorithmic FLOP per
        algorithmic HBM byte) is below `ridge` (= tensor peak / HBM peak) cannot reach the tensor roofline however good
        the kernel is: it is accounted against the HBM roofline instead. Returns per class the summed ms / flops /
        bytes / launch count."""
        out = dict(tensor=dict(ms=0.0, flops=0.0, bytes=0.0, n=0), hbm=dict(ms=0.0, flops=0.0, bytes=0.0, n=0))
        for ms, flops, nbytes in rows:
            c = out["hbm" if (nbytes > 0 and flops / nbytes < ridge) else "tensor"]
            c["ms"] += ms
            c["flops"] += flops
            c["bytes"] += nbytes
            c["n"] += 1
        return out

    def profile_kernels(self, steps=3, ridge=210.0):
        """Instrumented EAGER pass over the same workload: CUDA events (on the launching stream) around every
        implicit-GEMM conv / wgrad launch. Returns summed device ms and algorithmic FLOPs per step for the two
        tensor-core kernel families, the FCOSHead tower share, the eager step time and the C-ABI launch count, and the
        implicit-GEMM launches split into tensor-bound and HBM-bound ones (`by_bound`, see split_by_bound)."""
        from .engine import ConvPlan, WgradPlan
        nets = (self.teacher, self.student)
        recs = []

        def wrap(op):
            plan = getattr(op, "__self__", None)
            if not isinstance(plan, (ConvPlan, WgradPlan)):
                return op

            def timed(_op=op, _plan=plan):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _op()
                e1.record()
                recs.append((_plan, e0, e1))

            return timed

        saved = [(n, n.fwd_ops, n.bwd_ops) for n in nets]
        two, self.two_streams = self.two_streams, False   # serialise the teacher branch: per-kernel times must not overlap
        joint, self.joint_fwd = self.joint_fwd, None      # per-plan accounting: the two networks' launches one by one
        for n in nets:
            n.fwd_ops = [wrap(o) for o in n.fwd_ops]
            n.bwd_ops = [wrap(o) for o in n.bwd_ops]
        try:
            self._run_eager()  # warm
            torch.cuda.synchronize()
            recs.clear()
            L.reset_launch_count()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(steps):
                self._run_eager()
            t1.record()
            torch.cuda.synchronize()
        finally:
            for n, f, b in saved:
                n.fwd_ops, n.bwd_ops = f, b
            self.two_streams = two
            self.joint_fwd = joint
        out = dict(step_ms=t0.elapsed_time(t1) / steps, launches_per_step=L.launch_count / steps)
        fam = dict(conv_igemm=dict(ms=0.0, flops=0.0, n=0), conv_wgrad=dict(ms=0.0, flops=0.0, n=0))
        tower = dict(ms=0.0, flops=0.0)
        per_plan = {}
        rows = []
        for idx, (plan, e0, e1) in enumerate(recs):
            k = "conv_wgrad" if isinstance(plan, WgradPlan) else "conv_igemm"
            ms = e0.elapsed_time(e1)
            if k == "conv_igemm":
                rows.append((ms / steps, plan.flops / steps, float(getattr(plan, "bytes", 0)) / steps))
            key = (idx % (len(recs) // steps), plan.what, k)
            a = per_plan.setdefault(key, [0.0, plan.flops])
            a[0] += ms / steps
            fam[k]["ms"] += ms / steps
            fam[k]["flops"] += plan.flops / steps
            fam[k]["n"] += 1.0 / steps
            if k == "conv_igemm" and plan.what.startswith("head.tower") and "grad" not in plan.what:
                tower["ms"] += ms / steps
                tower["flops"] += plan.flops / steps
        out["plans"] = [dict(i=i, what=w, kind=k, us=round(1e3 * v[0], 2), gflop=round(v[1] / 1e9, 2),
                             tflops=round(v[1] / max(v[0], 1e-9) / 1e9, 1)) for (i, w, k), v in sorted(per_plan.items())]
        for k in fam:
            fam[k]["n"] = int(round(fam[k]["n"]))
        out.update(fam)
        out["rows"] = rows
        try:
            out["by_bound"] = self.split_by_bound(rows, ridge)
            for c in out["by_bound"].values():
                c["n"] = int(round(c["n"] / steps))
        except Exception:   # bookkeeping only
            out["by_bound"] = None
        if tower["ms"] > 0:
            out["head_tower"] = dict(tflops=round(tower["flops"] / (tower["ms"] * 1e-3) / 1e12, 1),
                                     ms_per_step=round(tower["ms"], 3), flops_per_step=tower["flops"],
                                     note="8 tower convs x 5 levels, teacher + student forward launches")
        return out

    def flops_per_step(self):
        return self.teacher.flops_fwd + self.student.flops_fwd + self.student.flops_bwd
