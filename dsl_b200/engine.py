"""Host-side execution plan of the FCOS detector on the C-ABI kernels (libdslb.so).

One `FCOSNet` = one set of weights (student or EMA teacher) + every activation / gradient buffer and every TMA launch
plan for a FIXED input shape (B, H, W). Building it allocates everything once; after that a forward or a
forward+backward is a flat list of pre-bound C calls on the current CUDA stream — no allocation, no host sync — so the
whole teacher+student step can be captured in a CUDA graph (see DSLEngine).

Reference functions replaced (paths relative to the reference root):
  ResNet.forward / Bottleneck.forward      mmdet/models/backbones/resnet.py:630-645, 262-301
  FPN.forward                              mmdet/models/necks/fpn.py:151-202
  FCOSHead.forward / forward_single        mmdet/models/dense_heads/fcos_head.py:118-168, anchor_free_head.py:197-217
  FCOSHead.loss (+ get_targets, ...)       mmdet/models/dense_heads/fcos_head.py:170-338, 562-726
  autograd backward of all of the above    (torch autograd in the reference)
"""
import ctypes as C
import math
import os

import torch

from . import _lib as L
from .arena import Arena
from .params import RESNET_BLOCKS, ParamStore, fpn_spec, head_spec, resnet_spec, rla_resnet_spec

BF16 = torch.bfloat16
STRIDES = (8, 16, 32, 64, 128)
REGRESS_RANGES = ((-1, 64), (64, 128), (128, 256), (256, 512), (512, 1e8))
INF = 1e8


def ceil_to(x, m):
    return (x + m - 1) // m * m


def conv_out(h, k, s, p):
    return (h + 2 * p - k) // s + 1


def seg_bytes(s):
    """Algorithmic HBM bytes of one conv segment (dict of dslb_conv_seg_t fields): every operand and the output touched
    exactly once — input pixels the taps reach, packed weights, output, residual and mask reads (DESIGN section 3:
    2 * npix * (Cin + Cout * (1 + residual + mask)) for the bf16 1x1 convs)."""
    N, H, W, Cin, Cout = (int(s[k]) for k in ("N", "H", "W", "Cin", "Cout"))
    R, S = int(s.get("R") or 1), int(s.get("S") or 1)
    stride, pad = int(s.get("stride") or 1), int(s.get("pad") or 0)
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    npo = N * Ho * Wo
    npi = npo if (R == 1 and S == 1) else N * H * W          # a strided 1x1 conv reads only the pixels it keeps
    out_elem = 4 if s.get("out_fp32") else 2
    extra = (1 if s.get("residual") is not None else 0) + (1 if s.get("relu_mask") is not None else 0)
    return 2 * npi * Cin + 2 * R * S * Cout * Cin + npo * Cout * (out_elem + 2 * extra)


class ConvPlan:
    """dslb_conv_plan_* wrapper; `segs` is a list of dicts of dslb_conv_seg_t fields (tensors for pointers)."""

    def __init__(self, segs, what="conv"):
        self.what = what
        self.keep = []
        self.segs = [dict(s) for s in segs]     # kept so that plans of two networks can be merged (DSLEngine joint forward)
        arr = (L.ConvSeg * len(segs))()
        for a, s in zip(arr, segs):
            for k, v in s.items():
                if isinstance(v, torch.Tensor):
                    self.keep.append(v)
                    setattr(a, k, v.data_ptr())
                elif v is not None:
                    setattr(a, k, v)
        self.plan = C.c_void_p()
        self._lib = L.lib   # the library that owns the plan handle
        L.check(L.lib.dslb_conv_plan_create(arr, len(segs), C.byref(self.plan)), what)
        self.flops = L.lib.dslb_conv_plan_flops(self.plan)
        try:   # algorithmic HBM traffic of one run: roofline bookkeeping only, never a reason to fail a plan
            self.bytes = sum(seg_bytes(s) for s in segs)
        except (KeyError, TypeError, ValueError):
            self.bytes = 0

    def run(self):
        L.check(L.lib.dslb_conv_plan_run(self.plan, L.cur_stream()), self.what)

    def __del__(self):
        try:
            if self.plan:
                self._lib.dslb_conv_plan_destroy(self.plan)
        except Exception:
            pass


class WgradPlan:
    def __init__(self, segs, what="wgrad"):
        self.what = what
        self.keep = []
        arr = (L.WgradSeg * len(segs))()
        for a, s in zip(arr, segs):
            for k, v in s.items():
                if isinstance(v, torch.Tensor):
                    self.keep.append(v)
                    setattr(a, k, v.data_ptr())
                else:
                    setattr(a, k, v)
        self.plan = C.c_void_p()
        self._lib = L.lib   # the library that owns the plan handle
        L.check(L.lib.dslb_wgrad_plan_create(arr, len(segs), C.byref(self.plan)), what)
        self.flops = L.lib.dslb_wgrad_plan_flops(self.plan)

    def run(self):
        L.check(L.lib.dslb_wgrad_plan_run(self.plan, L.cur_stream()), self.what)

    def __del__(self):
        try:
            if self.plan:
                self._lib.dslb_wgrad_plan_destroy(self.plan)
        except Exception:
            pass


class TablePlan:
    """dslb_pack_plan_* / dslb_unpack_plan_* wrapper: one launch over a table of per-conv descriptors."""

    def __init__(self, descs, kind, what):
        self.what = what
        self.keep = []
        cls = L.PackDesc if kind == "pack" else L.UnpackDesc
        arr = (cls * len(descs))()
        for a, d in zip(arr, descs):
            for k, v in d.items():
                if isinstance(v, torch.Tensor):
                    self.keep.append(v)
                    setattr(a, k, v.data_ptr())
                elif v is not None:
                    setattr(a, k, v)
        self.plan = C.c_void_p()
        self._lib = L.lib   # the library that owns the plan handle
        fn = L.lib.dslb_pack_plan_create if kind == "pack" else L.lib.dslb_unpack_plan_create
        L.check(fn(arr, len(descs), C.byref(self.plan)), what)

    def run(self):
        L.check(L.lib.dslb_table_plan_run(self.plan, L.cur_stream()), self.what)

    def __del__(self):
        try:
            if self.plan:
                self._lib.dslb_table_plan_destroy(self.plan)
        except Exception:
            pass


class ConvW:
    """Derived device state of one conv: packed bf16 operands (frozen BatchNorm folded in), shift vector, wgrad slot."""

    def __init__(self, net, wname, bias=None, bn=None, stride=1, pad=0, need_dgrad=False, trainable=False,
                 dy_ld=None):
        st = net.store
        self.net = net
        self.wname = wname
        self.w = st[wname]
        self.O, self.I, self.R, self.S = self.w.shape
        self.stride, self.pad = stride, pad
        self.cout_pad = ceil_to(self.O, 16)
        dev = st.device
        self.wp = net.mem.zeros(self.R * self.S, self.cout_pad, self.I, dtype=BF16)
        self.bn = bn
        self.scale = None
        if bn is not None:
            self.scale = torch.empty(self.O, dtype=torch.float32, device=dev)
            self.shift = torch.empty(self.O, dtype=torch.float32, device=dev)
        else:
            self.shift = st[bias] if bias is not None else None
        self.bias_name = bias
        self.need_dgrad = need_dgrad
        self.trainable = trainable
        self.dy_ld = dy_ld or ceil_to(self.O, 64)  # channel stride of the gradient buffer feeding dgrad/wgrad
        if need_dgrad:
            self.wpT = net.mem.zeros(self.R * self.S, ceil_to(self.I, 16), self.dy_ld, dtype=BF16)
        self.dw = None  # fp32 packed wgrad [taps][O][I]: a view into the net's zero arena (FCOSNet._alloc_arena)
        if trainable:
            net.want_arena(self, "dw", self.R * self.S * self.O * self.I, (self.R * self.S, self.O, self.I))

    def pack_descs(self):
        """dslb_pack_desc_t fields of the fprop (and dgrad) operands of this conv."""
        st = self.net.store
        bn = {}
        if self.bn is not None:
            bn = dict(bn_gamma=st[self.bn + ".weight"], bn_beta=st[self.bn + ".bias"],
                      bn_mean=st[self.bn + ".running_mean"], bn_var=st[self.bn + ".running_var"], bn_eps=1e-5,
                      scale_out=self.scale, shift_out=self.shift)
        out = [dict(w=self.w, out=self.wp, O=self.O, I=self.I, R=self.R, S=self.S, rows_pad=self.cout_pad,
                    cols_pad=self.I, mode=0, fill_padding=1, **bn)]
        if self.need_dgrad:
            out.append(dict(w=self.w, out=self.wpT, O=self.O, I=self.I, R=self.R, S=self.S,
                            rows_pad=self.wpT.shape[1], cols_pad=self.wpT.shape[2], mode=1, fill_padding=1, **bn))
        return out

    def unpack_desc(self):
        """packed fp32 wgrad -> OIHW gradient view (x folded BN scale)."""
        st = self.net.store
        d = dict(dw=self.dw, g=self.net.grad_view(self.wname), O=self.O, I=self.I, R=self.R, S=self.S, rows=self.O,
                 row_off=0)
        if self.bn is not None:
            d.update(bn_gamma=st[self.bn + ".weight"], bn_var=st[self.bn + ".running_var"], bn_eps=1e-5)
        return d

    # ---- segment builders -------------------------------------------------------------------------------
    def fseg(self, x, y, N, H, W, **kw):
        """fprop segment: y = conv(x)."""
        d = dict(x=x, w=self.wp, y=y, N=N, H=H, W=W, Cin=self.I, Cout=self.O, cout_pad=self.cout_pad, R=self.R,
                 S=self.S, stride=self.stride, pad=self.pad, ldc=self.O, shift=self.shift)
        d.update(kw)
        return d

    def dseg(self, dy, dx, N, Ho, Wo, Hin, Win, **kw):
        """dgrad segment: dx = conv_transpose(dy). Stride-2 1x1 convs scatter into the (Hin, Win) map."""
        d = dict(x=dy, w=self.wpT, y=dx, N=N, H=Ho, W=Wo, Cin=self.dy_ld, Cout=self.I, cout_pad=self.wpT.shape[1],
                 R=self.R, S=self.S, stride=1, pad=self.R - 1 - self.pad, ldc=self.I)
        if self.stride == 2:
            assert self.R == 1, "strided 3x3 convs go through dseg_upsampled"
            d.update(scatter2=1, Hs=Hin, Ws=Win)
        d.update(kw)
        return d

    def dseg_upsampled(self, dy_up, dx, N, Hin, Win, **kw):
        """dgrad of a stride-2 conv as a stride-1 dgrad over the zero-upsampled dY ([N][Hin][Win][dy_ld])."""
        d = dict(x=dy_up, w=self.wpT, y=dx, N=N, H=Hin, W=Win, Cin=self.dy_ld, Cout=self.I,
                 cout_pad=self.wpT.shape[1], R=self.R, S=self.S, stride=1, pad=self.R - 1 - self.pad, ldc=self.I)
        d.update(kw)
        return d

    def wseg(self, x, dy, N, H, W):
        return dict(x=x, dy=dy, dw=self.dw, N=N, H=H, W=W, Cin=self.I, Cout=self.O, ldy=self.dy_ld, dw_rows=self.O,
                    R=self.R, S=self.S, stride=self.stride, pad=self.pad)


class FCOSNet:
    """ResNet-50/101 + FPN + FCOSHead for a fixed (B, H, W); `train=True` also builds the backward plan."""

    def __init__(self, B, H, W, depth=50, num_classes=80, train=True, store=None, device="cuda", seed=0,
                 loss_weight=1.0, soft_weight=0.0, center_sampling=True, radius=1.5, norm_on_bbox=True,
                 max_boxes=1024, parity_outputs=False, parts="all", level_sizes=None, strides=STRIDES,
                 regress_ranges=REGRESS_RANGES, backbone="resnet", head_precision="bf16"):
        """parts="all": backbone + FPN + head on a (B, 3, H, W) image. parts="head": FCOSHead only, on caller-filled
        FPN maps self.p[l] of `level_sizes` (the standalone HEADS-registry module); backward then ends at self.dp.
        parts="backbone": ResNet only (BACKBONES module): stage outputs in self.stage_out, backward seeded by the caller
        in self.gc (gradients w.r.t. C3..C5, unmasked). parts="neck": FPN only (NECKS module) on caller-filled C3..C5
        maps of `level_sizes` (first three entries), outputs self.p, backward seeded in self.dp, input gradients in
        self.gc."""
        assert parts in ("all", "head", "backbone", "neck")
        # backbone="rla": RLA_ResNet (resnet_rla.py:140-400, the backbone of configs/fcos_semi/RLA_*.py) with
        # layers = RESNET_BLOCKS[depth]; see engine_rla.py
        assert backbone in ("resnet", "rla")
        self.backbone = backbone
        # "bf16x3": the accurate inference mode of the FCOSHead (split-bf16 operands, fp32 tower maps; _build_head_split)
        assert head_precision in ("bf16", "bf16x3")
        assert head_precision == "bf16" or not train, "the bf16x3 head is an inference mode (teacher / simple_test)"
        self.head_precision = head_precision
        assert parts in ("head", "neck") or (H % 32 == 0 and W % 32 == 0), \
            "inputs are padded to a multiple of 32 (Pad size_divisor=32)"
        self.parts = parts
        self.bb_prefix = "" if parts == "backbone" else "backbone."
        self.neck_prefix = "" if parts == "neck" else "neck."
        self.strides, self.regress_ranges = tuple(strides), tuple(regress_ranges)
        self.B, self.H, self.W, self.depth, self.C = B, H, W, depth, num_classes
        self.train = train
        self.dev = torch.device(device)
        self.loss_weight, self.soft_weight = float(loss_weight), float(soft_weight)
        self.center_sampling, self.radius, self.norm_on_bbox = center_sampling, radius, norm_on_bbox
        self.parity_outputs = parity_outputs
        if store is None:
            bb_spec = resnet_spec if backbone == "resnet" else \
                (lambda depth, prefix="backbone.": rla_resnet_spec(RESNET_BLOCKS[depth], prefix=prefix))
            spec = {"head": lambda: head_spec(num_classes), "backbone": lambda: bb_spec(depth, prefix=""),
                    "neck": lambda: fpn_spec(prefix=""),
                    "all": lambda: bb_spec(depth) + fpn_spec() + head_spec(num_classes)}[parts]()
            store = ParamStore(spec, device).init_reference(seed)
        self.store = store
        self.mem = Arena(self.dev)   # every static buffer of this plan: a few memsets instead of one per buffer
        self._arena_wants = []
        self.fwd_ops, self.bwd_ops, self.repack_ops = [], [], []
        self.bwd_meta = []
        self.extra_pack_descs = []
        self.convs = []
        self.flops_fwd = 0.0
        self.flops_bwd = 0.0
        if train:
            self.grad = self.mem.zeros(store.n_train, dtype=torch.float32)
        if parts in ("all", "backbone"):
            self._build_backbone()
        if parts == "neck":
            chans = (512, 1024, 2048)
            self.stage_out = [None] + [(self.buf(B, h, w, c), h, w, c) for (h, w), c in zip(level_sizes[:3], chans)]
        if parts in ("all", "neck"):
            self._build_fpn()
        if parts == "head":
            self.psize = [tuple(hw) for hw in level_sizes]
            self.p = [self.buf(B, h, w, 256) for (h, w) in self.psize]
        self.head_op_start = len(self.fwd_ops)
        if parts in ("all", "head"):
            self._build_head()
        if train:
            self.bwd_buckets = []   # (index one past the bucket's last backward op, flat grad range lo, hi)
            self._unpacked = set()
            if parts in ("all", "head"):
                self._build_loss()
                self._alloc_arena()
                self._build_head_bwd()
            else:
                self._alloc_arena()

                def zero_state():
                    L.zero(self.grad)
                    L.zero(self.arena)

                self.add_bwd(zero_state)
            if parts == "neck":
                self.dp = [self.buf(B, h, w, 256) for (h, w) in self.psize]   # seeds: gradients w.r.t. P3..P7
            if parts in ("all", "neck"):
                self._build_fpn_bwd()
            if parts == "all":
                self._emit_bucket(("neck.", "bbox_head."), head=True)   # + every bias (region B follows region A)
            if parts == "backbone":
                cs = self.stage_out[1:]
                self.gc = [self.buf(B, h, w, c) for (_, h, w, c) in cs]          # seeds: gradients w.r.t. C3..C5
                # the fused plan masks dC5 in the lateral dgrad epilogue; standalone, the last ReLU's mask is applied here
                self.add_bwd(self.ew("dslb_relu_family", self.gc[2], cs[2][0], self.gc[2], self.gc[2].numel(), 1))
            if parts in ("all", "backbone"):
                self._build_backbone_bwd()
            self._build_finish_bwd()
        self._build_pack_plans()
        self.repack(everything=True)

    # ------------------------------------------------------------------------------------------ helpers
    def buf(self, *shape, dtype=BF16):
        return self.mem.zeros(*shape, dtype=dtype)

    def want_arena(self, obj, attr, n, shape, dtype=torch.float32):
        """Reserve `n` elements of the zero arena (one memset at the start of every backward clears all of it);
        `obj.attr` becomes a view of `shape` once _alloc_arena() has run."""
        self._arena_wants.append((obj, attr, n, shape, dtype))

    def _alloc_arena(self):
        off = 0
        plan = []
        for obj, attr, n, shape, dtype in self._arena_wants:
            nf = n * (2 if dtype == torch.float64 else 1)
            plan.append((obj, attr, off, nf, shape, dtype))
            off += ceil_to(nf, 64)
        self.arena = self.mem.zeros(max(off, 64), dtype=torch.float32)
        for obj, attr, o, nf, shape, dtype in plan:
            v = self.arena[o:o + nf]
            if dtype == torch.float64:
                v = v.view(torch.float64)
            setattr(obj, attr, v.view(shape))

    def grad_view(self, name):
        o, n = self.store.offsets[name]
        assert o + n <= self.store.n_train, name + " is not trainable"
        return self.grad[o:o + n]

    def conv(self, *a, **k):
        c = ConvW(self, *a, **k)
        self.convs.append(c)
        return c

    def add_fwd(self, fn):
        self.fwd_ops.append(fn)

    def add_bwd(self, fn, side=False, tag=None, wait=None):
        """Append a backward op. side=True: independent of the dgrad critical path (weight / bias / GroupNorm-parameter
        gradients: they only read buffers no later op of the step rewrites) — may run on a second stream; `tag` names its completion event, `wait` names a side op the main stream must have finished
        before this op runs (buffer re-use), "__all__" = every side op issued so far."""
        self.bwd_ops.append(fn)
        self.bwd_meta.append((side, tag, wait))

    def plan_fwd(self, segs, what):
        p = ConvPlan(segs, what)
        self.flops_fwd += p.flops
        self.add_fwd(p.run)
        return p

    def plan_bwd(self, segs, what):
        p = ConvPlan(segs, what)
        self.flops_bwd += p.flops
        self.add_bwd(p.run)
        return p

    def plan_wgrad(self, segs, what):
        p = WgradPlan(segs, what)
        self.flops_bwd += p.flops
        self.add_bwd(p.run, side=True, tag=what)
        return p

    def ew(self, fn_name, *args):
        """Bind an elementwise C call (pointers resolved now, stream at call time)."""
        fn = getattr(L.lib, fn_name)
        cargs = [L.ptr(a) if isinstance(a, torch.Tensor) else a for a in args]
        keep = [a for a in args if isinstance(a, torch.Tensor)]

        def run(_fn=fn, _a=cargs, _k=keep, _n=fn_name):
            L.check(_fn(*_a, L.cur_stream()), _n)

        return run

    # ------------------------------------------------------------------------------------------ backbone
    def _build_backbone(self):
        if self.backbone == "rla":
            from . import engine_rla
            return engine_rla.build_backbone(self)
        B, H, W = self.B, self.H, self.W
        st = self.store
        self.img = self.buf(B, 3, H, W, dtype=torch.float32)  # NCHW fp32 input, as the reference feeds it
        H2, W2 = conv_out(H, 7, 2, 3), conv_out(W, 7, 2, 3)
        H4, W4 = conv_out(H2, 3, 2, 1), conv_out(W2, 3, 2, 1)
        # stem: one fused kernel (im2col tile assembled in shared memory -> tcgen05 MMA -> folded BN + ReLU), max-pool
        self.stem_out = self.buf(B, H2, W2, 64)
        self.x0 = self.buf(B, H4, W4, 64)
        self.img4 = self.buf(B, H, W, 4)   # NHWC bf16 copy of the image, channels padded 3 -> 4 (stem workspace)
        self.add_fwd(self.ew("dslb_stem_conv", self.img, st[self.bb_prefix + "conv1.weight"],
                             st[self.bb_prefix + "bn1.weight"], st[self.bb_prefix + "bn1.bias"],
                             st[self.bb_prefix + "bn1.running_mean"], st[self.bb_prefix + "bn1.running_var"],
                             1e-5, self.img4, self.stem_out, B, H, W))
        self.flops_fwd += 2.0 * B * H2 * W2 * 64 * 147
        self.add_fwd(self.ew("dslb_maxpool3x3s2", self.stem_out, self.x0, B, H2, W2, 64))

        self.blocks = []
        x, h, w, inpl = self.x0, H4, W4, 64
        self.stage_out = []
        for li, nb in enumerate(RESNET_BLOCKS[self.depth]):
            planes = 64 * 2 ** li
            trainable = self.train and li >= 1  # frozen_stages = 1
            for bi in range(nb):
                s = 2 if (bi == 0 and li > 0) else 1
                ho, wo = conv_out(h, 1, s, 0), conv_out(w, 1, s, 0)
                p = f"{self.bb_prefix}layer{li + 1}.{bi}"
                # dgrad into the block input is needed unless that input is the frozen layer1 output
                dgrad_in = trainable and not (li == 1 and bi == 0)
                blk = dict(li=li, bi=bi, stride=s, hin=h, win=w, h=ho, w=wo, cin=inpl, planes=planes, xin=x,
                           trainable=trainable, dgrad_in=dgrad_in)
                blk["c1"] = self.conv(p + ".conv1.weight", bn=p + ".bn1", stride=s, need_dgrad=dgrad_in,
                                      trainable=trainable)
                blk["c2"] = self.conv(p + ".conv2.weight", bn=p + ".bn2", pad=1, need_dgrad=trainable,
                                      trainable=trainable)
                blk["c3"] = self.conv(p + ".conv3.weight", bn=p + ".bn3", need_dgrad=trainable, trainable=trainable)
                blk["a1"] = self.buf(B, ho, wo, planes)
                blk["a2"] = self.buf(B, ho, wo, planes)
                blk["out"] = self.buf(B, ho, wo, planes * 4)
                segs = [blk["c1"].fseg(x, blk["a1"], B, h, w, relu_nch=planes)]
                if bi == 0:
                    blk["ds"] = self.conv(p + ".downsample.0.weight", bn=p + ".downsample.1", stride=s,
                                          need_dgrad=dgrad_in, trainable=trainable)
                    blk["idn"] = self.buf(B, ho, wo, planes * 4)
                    segs.append(blk["ds"].fseg(x, blk["idn"], B, h, w))
                self.plan_fwd(segs, p + ".conv1")
                self.plan_fwd([blk["c2"].fseg(blk["a1"], blk["a2"], B, ho, wo, relu_nch=planes)], p + ".conv2")
                self.plan_fwd([blk["c3"].fseg(blk["a2"], blk["out"], B, ho, wo,
                                              residual=blk.get("idn", x), relu_nch=planes * 4)], p + ".conv3")
                self.blocks.append(blk)
                x, h, w, inpl = blk["out"], ho, wo, planes * 4
            self.stage_out.append((x, h, w, inpl))

    # ------------------------------------------------------------------------------------------ FPN
    def _build_fpn(self):
        B = self.B
        tr = self.train
        self.lat, self.fpnc = [], []
        self.lm, self.p = [], []
        cs = self.stage_out[1:]  # C3, C4, C5
        segs = []
        for i, (x, h, w, c) in enumerate(cs):
            cw = self.conv(f"{self.neck_prefix}lateral_convs.{i}.conv.weight",
                           bias=f"{self.neck_prefix}lateral_convs.{i}.conv.bias",
                           need_dgrad=tr, trainable=tr)
            self.lat.append(cw)
            self.lm.append(self.buf(B, h, w, 256))
            segs.append(cw.fseg(x, self.lm[i], B, h, w))
        self.plan_fwd(segs, "fpn.laterals")
        for i in (2, 1):
            (_, h, w, _), (_, hs, ws, _) = cs[i - 1], cs[i]
            self.add_fwd(self.ew("dslb_upsample_add", self.lm[i - 1], self.lm[i], B, h, w, hs, ws, 256))
        segs = []
        self.psize = []
        for i, (x, h, w, c) in enumerate(cs):
            cw = self.conv(f"{self.neck_prefix}fpn_convs.{i}.conv.weight",
                           bias=f"{self.neck_prefix}fpn_convs.{i}.conv.bias", pad=1,
                           need_dgrad=tr, trainable=tr)
            self.fpnc.append(cw)
            self.p.append(self.buf(B, h, w, 256))
            self.psize.append((h, w))
            segs.append(cw.fseg(self.lm[i], self.p[i], B, h, w))
        self.plan_fwd(segs, "fpn.out")
        h5, w5 = self.psize[2]
        h6, w6 = conv_out(h5, 3, 2, 1), conv_out(w5, 3, 2, 1)
        h7, w7 = conv_out(h6, 3, 2, 1), conv_out(w6, 3, 2, 1)
        self.psize += [(h6, w6), (h7, w7)]
        for i in (3, 4):
            self.fpnc.append(self.conv(f"{self.neck_prefix}fpn_convs.{i}.conv.weight",
                                       bias=f"{self.neck_prefix}fpn_convs.{i}.conv.bias",
                                       stride=2, pad=1, need_dgrad=tr, trainable=tr))
        self.p.append(self.buf(B, h6, w6, 256))
        self.p.append(self.buf(B, h7, w7, 256))
        self.r6 = self.buf(B, h6, w6, 256)
        self.plan_fwd([self.fpnc[3].fseg(self.p[2], self.p[3], B, h5, w5)], "fpn.p6")
        self.add_fwd(self.ew("dslb_relu_family", self.p[3], None, self.r6, self.p[3].numel(), 0))
        self.plan_fwd([self.fpnc[4].fseg(self.r6, self.p[4], B, h6, w6)], "fpn.p7")

    # ------------------------------------------------------------------------------------------ head
    def _build_head_split(self):
        """FCOSHead forward with split-bf16 operands (north_star: FCOSHead outputs within 1e-3 of the fp32 reference,
        fcos_head.py:118-168; plain bf16 measures 1.4e-3 on identical inputs). Every activation travels as [hi | lo | hi]
        (3 x 256 bf16 channels), every weight as [w_hi | w_hi | w_lo]: the SAME tcgen05 implicit-GEMM launches with
        K = 3 x 2304 accumulate hi*w_hi + lo*w_hi + hi*w_lo in fp32; the tower maps are stored in fp32 (conv epilogue:
        fp32 direct store + GroupNorm statistics) and dslb_gn_apply_relu_split turns them into the next split operand.
        3x the tower FLOPs, inference only."""
        B, C = self.B, self.C
        st = self.store
        nl = len(self.psize)
        f32 = torch.float32
        self.tower = {"cls": [], "reg": []}
        split_w = []      # (packed split operand, [(fp32 master weight, first row)], rows_pad)

        def split_operand(parts, rows_pad):
            wp = self.mem.zeros(9, rows_pad, 768, dtype=BF16)
            split_w.append((wp, parts, rows_pad))
            return wp

        for br in ("cls", "reg"):
            for i in range(4):
                self.tower[br].append(split_operand([(st[f"bbox_head.{br}_convs.{i}.conv.weight"], 0)], 256))
        self.gn_stats = self.mem.zeros(2, 4, nl, B, 32, L.GN_STAT_STRIDE, dtype=torch.float64)
        self.gn_mr = self.mem.zeros(2, 4, nl, B, 32, 4, dtype=f32)
        self.add_fwd(lambda: L.zero(self.gn_stats))
        self.ps = [self.buf(B, h, w, 768) for (h, w) in self.psize]
        self.y = {br: [[self.buf(B, h, w, 256, dtype=f32) for (h, w) in self.psize] for _ in range(4)] for br in ("cls", "reg")}
        self.z = {br: [[self.buf(B, h, w, 768) for (h, w) in self.psize] for _ in range(4)] for br in ("cls", "reg")}
        for l, (h, w) in enumerate(self.psize):
            self.add_fwd(self.ew("dslb_bf16_to_split", self.p[l], self.ps[l], B * h * w, 256))
        for i in range(4):
            segs, gsegs = [], []
            for bi_, br in enumerate(("cls", "reg")):
                for l, (h, w) in enumerate(self.psize):
                    x = self.ps[l] if i == 0 else self.z[br][i - 1][l]
                    segs.append(dict(x=x, w=self.tower[br][i], y=self.y[br][i][l], N=B, H=h, W=w, Cin=768, Cout=256,
                                     cout_pad=256, R=3, S=3, stride=1, pad=1, ldc=256, out_fp32=1,
                                     shift=st[f"bbox_head.{br}_convs.{i}.conv.bias"], gn_stats=self.gn_stats[bi_, i, l],
                                     gn_cpg=8))
                    gsegs.append(dict(x=self.y[br][i][l], y=self.z[br][i][l], stats=self.gn_stats[bi_, i, l],
                                      gamma=st[f"bbox_head.{br}_convs.{i}.gn.weight"],
                                      beta=st[f"bbox_head.{br}_convs.{i}.gn.bias"], mr=self.gn_mr[bi_, i, l], N=B,
                                      HW=h * w))
            p = ConvPlan(segs, f"head.tower{i}.bf16x3")
            self.flops_fwd += p.flops / 3.0      # algorithmic FLOPs: the reference's conv, not the three partial products
            self.add_fwd(p.run)
            arr = (L.GnSeg * len(gsegs))()
            keep = []
            for a, g in zip(arr, gsegs):
                for k, v in g.items():
                    if isinstance(v, torch.Tensor):
                        keep.append(v)
                        setattr(a, k, v.data_ptr())
                    else:
                        setattr(a, k, v)

            def gn_run(_arr=arr, _n=len(gsegs), _k=keep):
                L.check(L.lib.dslb_gn_apply_relu_split(_arr, _n, 256, 32, 1e-5, L.cur_stream()), "gn_apply_split")

            self.add_fwd(gn_run)
        cls_wp = split_operand([(st["bbox_head.conv_cls.weight"], 0)], ceil_to(C, 16))
        rc_wp = split_operand([(st["bbox_head.conv_reg.weight"], 0), (st["bbox_head.conv_centerness.weight"], 4)], 16)
        self.rc_scale = torch.ones(nl, 8, dtype=f32, device=self.dev)
        self.rc_shift = self.mem.zeros(nl, 8, dtype=f32)
        self.scale_vals = torch.ones(nl, dtype=f32, device=self.dev)
        self.level_mult = torch.tensor([float(s) for s in self.strides[:nl]], dtype=f32, device=self.dev)   # eval: x stride
        offs = [st.offsets[f"bbox_head.scales.{l}.scale"][0] for l in range(nl)]
        self.scale_stride = offs[1] - offs[0]

        def repack_split():
            # [w_hi | w_hi | w_lo]: w_hi is the bf16 rounding the pack kernel applies to the first two thirds, w_lo the
            # rounding of the remainder (torch ops: this optional path is refreshed only after a weight update)
            for wp, parts, rows_pad in split_w:
                rows = []
                for wt, r0 in parts:
                    wf = wt.detach().float()
                    wl = wf - wf.to(BF16).float()
                    rows.append((r0, torch.cat([wf, wf, wl], dim=1)))
                O = max(r0 + c.shape[0] for r0, c in rows)
                cat = torch.zeros(O, 768, 3, 3, dtype=f32, device=self.dev)
                for r0, c in rows:
                    cat[r0:r0 + c.shape[0]] = c
                L.check(L.lib.dslb_pack_weight(L.ptr(cat), L.ptr(wp), O, 768, 3, 3, rows_pad, 768, None, 0, L.cur_stream()),
                        "pack split")
                self._split_keep = cat   # stays alive until the (stream-ordered) next allocation re-uses it
            L.check(L.lib.dslb_fcos_regctr_affine(
                L.ptr(st["bbox_head.scales.0.scale"]), self.scale_stride, L.ptr(st["bbox_head.conv_reg.bias"]),
                L.ptr(st["bbox_head.conv_centerness.bias"]), L.ptr(self.level_mult), L.ptr(self.rc_scale),
                L.ptr(self.rc_shift), L.ptr(self.scale_vals), nl, L.cur_stream()), "regctr_affine")

        self.repack_ops.append(repack_split)
        self.cls_out = [self.buf(B, h, w, C, dtype=f32) for (h, w) in self.psize]
        self.rc_out = [self.buf(B, h, w, 8, dtype=f32) for (h, w) in self.psize]
        segs = []
        for l, (h, w) in enumerate(self.psize):
            segs.append(dict(x=self.z["cls"][3][l], w=cls_wp, y=self.cls_out[l], N=B, H=h, W=w, Cin=768, Cout=C,
                             cout_pad=ceil_to(C, 16), R=3, S=3, stride=1, pad=1, ldc=C, out_fp32=1,
                             shift=st["bbox_head.conv_cls.bias"]))
        for l, (h, w) in enumerate(self.psize):
            segs.append(dict(x=self.z["reg"][3][l], w=rc_wp, y=self.rc_out[l], N=B, H=h, W=w, Cin=768, Cout=5,
                             cout_pad=16, R=3, S=3, stride=1, pad=1, ldc=8, out_fp32=1, relu_nch=4,
                             scale=self.rc_scale[l], shift=self.rc_shift[l]))
        p = ConvPlan(segs, "head.predictors.bf16x3")
        self.flops_fwd += p.flops / 3.0
        self.add_fwd(p.run)

    def _build_head(self):
        if self.head_precision == "bf16x3":
            return self._build_head_split()
        B, C = self.B, self.C
        tr = self.train
        st = self.store
        nl = len(self.psize)
        self.tower = {"cls": [], "reg": []}
        for br in ("cls", "reg"):
            for i in range(4):
                self.tower[br].append(self.conv(f"bbox_head.{br}_convs.{i}.conv.weight",
                                                bias=f"bbox_head.{br}_convs.{i}.conv.bias", pad=1, need_dgrad=tr,
                                                trainable=tr))
        # GroupNorm statistics of all (branch, layer, level) maps in one buffer -> one memset per pass
        self.gn_stats = self.mem.zeros(2, 4, nl, B, 32, L.GN_STAT_STRIDE, dtype=torch.float64)
        self.gn_mr = self.mem.zeros(2, 4, nl, B, 32, 4, dtype=torch.float32)
        self.add_fwd(lambda: L.zero(self.gn_stats))
        self.y = {br: [[self.buf(B, h, w, 256) for (h, w) in self.psize] for _ in range(4)] for br in ("cls", "reg")}
        self.z = {br: [[self.buf(B, h, w, 256) for (h, w) in self.psize] for _ in range(4)] for br in ("cls", "reg")}
        for i in range(4):
            segs, gsegs = [], []
            for bi_, br in enumerate(("cls", "reg")):
                for l, (h, w) in enumerate(self.psize):
                    x = self.p[l] if i == 0 else self.z[br][i - 1][l]
                    segs.append(self.tower[br][i].fseg(x, self.y[br][i][l], B, h, w,
                                                       gn_stats=self.gn_stats[bi_, i, l], gn_cpg=8))
                    gsegs.append(dict(x=self.y[br][i][l], y=self.z[br][i][l], stats=self.gn_stats[bi_, i, l],
                                      gamma=st[f"bbox_head.{br}_convs.{i}.gn.weight"],
                                      beta=st[f"bbox_head.{br}_convs.{i}.gn.bias"], mr=self.gn_mr[bi_, i, l], N=B,
                                      HW=h * w))
            self.plan_fwd(segs, f"head.tower{i}")
            self.add_fwd(self._gn_apply(gsegs))
        # predictors: conv_cls (N=80, fp32) on the cls tower; conv_reg + conv_centerness fused (N=5 -> 16, fp32
        # rows of 8) on the reg tower with bbox = relu(scale_l * (conv + b)) (x stride in eval mode)
        self.cls_w = self.conv("bbox_head.conv_cls.weight", bias="bbox_head.conv_cls.bias", pad=1, need_dgrad=tr,
                               trainable=tr, dy_ld=128)
        self.rc_wp = self.mem.zeros(9, 16, 256, dtype=BF16)   # rows 0-3 conv_reg, 4 conv_centerness
        self.rc_wpT = self.mem.zeros(9, 256, 64, dtype=BF16)
        self.rc_scale = torch.ones(nl, 8, dtype=torch.float32, device=self.dev)
        self.rc_shift = self.mem.zeros(nl, 8, dtype=torch.float32)
        self.scale_vals = torch.ones(nl, dtype=torch.float32, device=self.dev)
        self.level_mult = torch.tensor([float(s) if not self.train else 1.0 for s in self.strides[:nl]],
                                       dtype=torch.float32, device=self.dev)
        offs = [st.offsets[f"bbox_head.scales.{l}.scale"][0] for l in range(nl)]
        self.scale_stride = offs[1] - offs[0]
        assert all(offs[l] == offs[0] + l * self.scale_stride for l in range(nl))
        for wname, row in (("bbox_head.conv_reg.weight", 0), ("bbox_head.conv_centerness.weight", 4)):
            w = st[wname]
            self.extra_pack_descs.append((tr, dict(w=w, out=self.rc_wp, O=w.shape[0], I=256, R=3, S=3, rows_pad=16,
                                                   cols_pad=256, row_off=row, mode=0, fill_padding=0)))
            if tr:
                self.extra_pack_descs.append((tr, dict(w=w, out=self.rc_wpT, O=w.shape[0], I=256, R=3, S=3,
                                                       rows_pad=256, cols_pad=64, col_off=row, mode=1,
                                                       fill_padding=0)))

        def repack_regctr():
            L.check(L.lib.dslb_fcos_regctr_affine(
                L.ptr(st["bbox_head.scales.0.scale"]), self.scale_stride, L.ptr(st["bbox_head.conv_reg.bias"]),
                L.ptr(st["bbox_head.conv_centerness.bias"]), L.ptr(self.level_mult), L.ptr(self.rc_scale),
                L.ptr(self.rc_shift), L.ptr(self.scale_vals), nl, L.cur_stream()), "regctr_affine")

        self.repack_ops.append(repack_regctr)
        self.cls_out = [self.buf(B, h, w, C, dtype=torch.float32) for (h, w) in self.psize]
        self.rc_out = [self.buf(B, h, w, 8, dtype=torch.float32) for (h, w) in self.psize]
        segs = []
        for l, (h, w) in enumerate(self.psize):
            segs.append(self.cls_w.fseg(self.z["cls"][3][l], self.cls_out[l], B, h, w, out_fp32=1))
        for l, (h, w) in enumerate(self.psize):
            segs.append(dict(x=self.z["reg"][3][l], w=self.rc_wp, y=self.rc_out[l], N=B, H=h, W=w, Cin=256, Cout=5,
                             cout_pad=16, R=3, S=3, stride=1, pad=1, ldc=8, out_fp32=1, relu_nch=4,
                             scale=self.rc_scale[l], shift=self.rc_shift[l]))
        self.plan_fwd(segs, "head.predictors")

    def _gn_apply(self, gsegs):
        arr = (L.GnSeg * len(gsegs))()
        keep = []
        for a, s in zip(arr, gsegs):
            for k, v in s.items():
                if isinstance(v, torch.Tensor):
                    keep.append(v)
                    setattr(a, k, v.data_ptr())
                else:
                    setattr(a, k, v)
        n = len(gsegs)
        nb = L.lib.dslb_gn_bwd_blocks(arr, n)
        host = (C.c_int * (2 * nb))()
        L.check(L.lib.dslb_gn_bwd_plan(arr, n, host), "gn_plan")
        tab = torch.tensor(list(host), dtype=torch.int32, device=self.dev)
        keep.append(tab)

        def run(_arr=arr, _k=keep, _n=n, _tab=tab, _nb=nb):
            L.check(L.lib.dslb_gn_apply_relu_tab(_arr, _n, 256, 32, 1e-5, L.ptr(_tab), _nb, L.cur_stream()), "gn_apply")

        return run

    # ------------------------------------------------------------------------------------------ loss
    def _fill_levels(self, with_grads):
        nl = len(self.psize)
        arr = (L.FcosLevel * nl)()
        for l, (h, w) in enumerate(self.psize):
            a = arr[l]
            a.cls = self.cls_out[l].data_ptr()
            a.regctr = self.rc_out[l].data_ptr()
            if with_grads:
                a.dcls_bf16 = self.dcls[l].data_ptr()
                a.dregctr_bf16 = self.drc[l].data_ptr()
                if self.parity_outputs:
                    a.dcls_f32 = self.dcls_f32[l].data_ptr()
                    a.dregctr_f32 = self.drc_f32[l].data_ptr()
            a.h, a.w, a.stride = h, w, self.strides[l]
            a.ld_cls, a.ld_dcls, a.ld_dreg = self.C, 128, 64
            a.rr_lo, a.rr_hi = float(self.regress_ranges[l][0]), float(self.regress_ranges[l][1])
            a.scale = 1.0
            a.cs_radius = float(self.strides[l] * self.radius)
        return arr

    def _build_loss(self):
        B, C = self.B, self.C
        nl = len(self.psize)
        self.npoints = B * sum(h * w for (h, w) in self.psize)
        self.dcls = [self.buf(B, h, w, 128) for (h, w) in self.psize]   # bf16, cols >= C stay zero
        self.drc = [self.buf(B, h, w, 64) for (h, w) in self.psize]     # bf16, cols >= 5 stay zero
        if self.parity_outputs:
            self.dcls_f32 = [self.buf(B, h, w, C, dtype=torch.float32) for (h, w) in self.psize]
            self.drc_f32 = [self.buf(B, h, w, 8, dtype=torch.float32) for (h, w) in self.psize]
        P = self.npoints
        self.labels = self.mem.zeros(P, dtype=torch.int64)
        self.bbox_targets = self.mem.zeros(P, 4, dtype=torch.float32)
        self.weights = self.mem.zeros(P, dtype=torch.float32)
        self.ctr_targets = self.mem.zeros(P, dtype=torch.float32)
        self.counts = self.mem.zeros(2, dtype=torch.float64)
        self.norm = torch.ones(2, dtype=torch.float32, device=self.dev)
        self.loss_acc = self.mem.zeros(16, dtype=torch.float32)  # one memset for both accumulators
        self.loss_sums = self.loss_acc[:8].view(torch.float64)
        self.dscale = self.loss_acc[8:16]
        self.loss_f32 = self.mem.zeros(4, dtype=torch.float32)   # the four loss scalars as the detector returns them
        self.max_boxes = 1024
        self.gt_boxes = self.mem.zeros(self.max_boxes, 4, dtype=torch.float32)
        self.gt_labels = self.mem.zeros(self.max_boxes, dtype=torch.int64)
        self.gt_off = self.mem.zeros(B + 1, dtype=torch.int32)
        self.ig_boxes = self.mem.zeros(self.max_boxes, 4, dtype=torch.float32)
        self.ig_off = self.mem.zeros(B + 1, dtype=torch.int32)
        self.use_ignore = True
        self.levels_arr = self._fill_levels(True)
        self.world_size = 1.0
        # labeled images come first (fcos_head.py:227-233): B/2 for an even batch, (B-1)/2 with the SI extra image
        self.n_labeled = B // 2 if B % 2 == 0 else (B - 1) // 2
        self.si_weight = 0.0
        self.want_arena(self, "gn_red", 2 * 4 * nl * B * 256 * 2, (2, 4, nl, B, 256, 2), torch.float64)
        # GroupNorm-backward group sums (sum gamma*dy, sum gamma*dy*xhat) left by the epilogue of the dgrad that produces dz
        self.want_arena(self, "gn_bsum", 2 * 4 * nl * B * 32 * L.GN_STAT_STRIDE, (2, 4, nl, B, 32, L.GN_STAT_STRIDE),
                        torch.float64)
        self.want_arena(self, "rc_dw", 9 * 5 * 256, (9, 5, 256))
        self.want_arena(self, "rc_db", 8, (8,))
        scale_idx = [self.store.offsets[f"bbox_head.scales.{l}.scale"][0] for l in range(nl)]
        self.scale_idx = torch.tensor(scale_idx, dtype=torch.long, device=self.dev)

    def set_targets(self, gt_bboxes, gt_labels, gt_bboxes_ignore=None):
        """Upload per-image box lists (torch tensors, any device) into the fixed device buffers."""
        B = self.B
        assert len(gt_bboxes) == B and len(gt_labels) == B
        offs = [0]
        for b in gt_bboxes:
            offs.append(offs[-1] + int(b.shape[0]))
        assert offs[-1] <= self.max_boxes, "too many GT boxes for the preallocated buffer"
        if offs[-1] > 0:
            self.gt_boxes[:offs[-1]].copy_(torch.cat([b.reshape(-1, 4) for b in gt_bboxes]).to(torch.float32),
                                           non_blocking=True)
            self.gt_labels[:offs[-1]].copy_(torch.cat([l.reshape(-1) for l in gt_labels]).to(torch.int64),
                                            non_blocking=True)
        self.gt_off.copy_(torch.tensor(offs, dtype=torch.int32), non_blocking=True)
        self.use_ignore = gt_bboxes_ignore is not None
        if self.use_ignore:
            ioffs = [0]
            for b in gt_bboxes_ignore:
                ioffs.append(ioffs[-1] + int(b.shape[0]))
            assert ioffs[-1] <= self.max_boxes
            if ioffs[-1] > 0:
                self.ig_boxes[:ioffs[-1]].copy_(torch.cat([b.reshape(-1, 4) for b in gt_bboxes_ignore])
                                                .to(torch.float32), non_blocking=True)
            self.ig_off.copy_(torch.tensor(ioffs, dtype=torch.int32), non_blocking=True)

    def run_targets(self):
        """kernel 1 of the loss: labels / bbox_targets / weights / centerness targets + local normaliser sums."""
        L.zero(self.counts)
        lw = self.loss_weight
        L.check(L.lib.dslb_fcos_targets(
            self.levels_arr, len(self.psize), self.B, self.C, L.ptr(self.gt_boxes), L.ptr(self.gt_labels),
            L.ptr(self.gt_off), L.ptr(self.ig_boxes) if self.use_ignore else None,
            L.ptr(self.ig_off) if self.use_ignore else None, int(self.center_sampling), int(self.norm_on_bbox), lw,
            self.n_labeled, L.ptr(self.labels), L.ptr(self.bbox_targets), L.ptr(self.weights),
            L.ptr(self.ctr_targets), L.ptr(self.counts), L.cur_stream()), "fcos_targets")

    def run_loss(self):
        """kernel 2 (after `counts` has been all-reduced over ranks): losses + gradients of the head outputs."""
        s = L.cur_stream()
        L.check(L.lib.dslb_fcos_norm(L.ptr(self.counts), float(self.world_size), L.ptr(self.norm), s), "fcos_norm")
        L.zero(self.loss_acc)
        nl = len(self.psize)
        L.check(L.lib.dslb_fcos_loss(
            self.levels_arr, nl, self.B, self.C, L.ptr(self.labels), L.ptr(self.bbox_targets), L.ptr(self.weights),
            L.ptr(self.ctr_targets), L.ptr(self.norm), 0.25, 2.0, self.loss_weight, self.n_labeled,
            float(self.si_weight), L.ptr(self.scale_vals), L.ptr(self.loss_sums), L.ptr(self.dscale), s), "fcos_loss")
        L.check(L.lib.dslb_f64_to_f32(L.ptr(self.loss_sums), L.ptr(self.loss_f32), 4, s), "loss scalars")

    # ------------------------------------------------------------------------------------------ head backward
    def _build_head_bwd(self):
        B = self.B
        st = self.store
        nl = len(self.psize)
        br_names = ("cls", "reg")
        self.dz = {br: [self.buf(B, h, w, 256) for (h, w) in self.psize] for br in br_names}
        # two sets, alternating between tower layers: the (side-stream) wgrad of layer i may still be reading its dY
        # while the GroupNorm backward of layer i-1 writes the other set
        self.dy2 = [{br: [self.buf(B, h, w, 256) for (h, w) in self.psize] for br in br_names} for _ in range(2)]
        self.dp = [self.buf(B, h, w, 256) for (h, w) in self.psize]  # gradient w.r.t. the FPN outputs

        # gradient buffers cleared at the head of the backward — unless a trainer does it earlier, off the critical path
        # (zero_in_bwd = False + its own zero_state() call under the forward pass)
        self.zero_in_bwd = True
        self.add_bwd(lambda: self.zero_state() if self.zero_in_bwd else None)
        # --- predictors: wgrad (10 segs), bias grads, dgrad into the last tower outputs
        wsegs = []
        for l, (h, w) in enumerate(self.psize):
            wsegs.append(self.cls_w.wseg(self.z["cls"][3][l], self.dcls[l], B, h, w))
        for l, (h, w) in enumerate(self.psize):
            wsegs.append(dict(x=self.z["reg"][3][l], dy=self.drc[l], dw=self.rc_dw, N=B, H=h, W=w, Cin=256, Cout=5,
                              ldy=64, dw_rows=5, R=3, S=3, stride=1, pad=1))
        # order of a layer's weight-gradient launch (side stream) and its dgrad (main stream), DSLB_WGRAD_AFTER_DGRAD:
        # 0 = wgrad issued first (both start together and time-slice the SMs), 1 = tower wgrads issued behind their dgrad
        # (they then run under the next layer's GroupNorm backward, which otherwise leaves the tensor pipe idle), 2 = the
        # predictors' too
        wg_mode = int(os.environ.get("DSLB_WGRAD_AFTER_DGRAD", "1"))

        def pred_wgrads(wsegs=wsegs):
            self.plan_wgrad(wsegs, "head.predictors.wgrad")
            g_cls_b = self.grad_view("bbox_head.conv_cls.bias")
            for l, (h, w) in enumerate(self.psize):
                self.add_bwd(self.ew("dslb_colsum", self.dcls[l], g_cls_b, B * h * w, 128, self.C), side=True, tag="colsum")
                self.add_bwd(self.ew("dslb_colsum", self.drc[l], self.rc_db, B * h * w, 64, 5), side=True, tag="colsum")

        if wg_mode < 2:
            pred_wgrads()
        # DSLB_GN_BWD_FUSED=1: the dgrad that produces dz of tower layer i also accumulates that layer's GroupNorm-backward
        # group sums in its epilogue (dslb_conv_seg_t::gnb_*), so dslb_gn_bwd is ONE pass over (x, dz) instead of reduce +
        # apply. Opt-in: measured on B200 the 8 epilogue warps become the bound of the dgrad (126 -> 168 us per tower
        # layer) and eat what the dropped reduce pass (58 us alone) gives back: 8.998 vs 8.980 ms per step over three
        # interleaved runs (profiles/r02g_gn_bwd_fused_ab.txt).
        fuse_gnb = os.environ.get("DSLB_GN_BWD_FUSED", "0") == "1"

        def gnb(bi_, br, i, l):
            if not fuse_gnb:
                return {}
            return dict(gnb_x=self.y[br][i][l], gnb_mr=self.gn_mr[bi_, i, l],
                        gnb_gamma=st[f"bbox_head.{br}_convs.{i}.gn.weight"],
                        gnb_beta=st[f"bbox_head.{br}_convs.{i}.gn.bias"], gnb_sums=self.gn_bsum[bi_, i, l], gn_cpg=8)

        dsegs = []
        for l, (h, w) in enumerate(self.psize):
            dsegs.append(self.cls_w.dseg(self.dcls[l], self.dz["cls"][l], B, h, w, h, w, **gnb(0, "cls", 3, l)))
        for l, (h, w) in enumerate(self.psize):
            dsegs.append(dict(x=self.drc[l], w=self.rc_wpT, y=self.dz["reg"][l], N=B, H=h, W=w, Cin=64, Cout=256,
                              cout_pad=256, R=3, S=3, stride=1, pad=1, ldc=256, **gnb(1, "reg", 3, l)))
        self.plan_bwd(dsegs, "head.predictors.dgrad")
        if wg_mode >= 2:
            pred_wgrads()
        # --- towers, last layer first
        for i in (3, 2, 1, 0):
            self.dy = self.dy2[i & 1]
            gsegs = []
            for bi_, br in enumerate(br_names):
                for l, (h, w) in enumerate(self.psize):
                    gsegs.append(dict(x=self.y[br][i][l], y=self.dy[br][l], dz=self.dz[br][l],
                                      stats=self.gn_stats[bi_, i, l],
                                      gamma=st[f"bbox_head.{br}_convs.{i}.gn.weight"],
                                      beta=st[f"bbox_head.{br}_convs.{i}.gn.bias"], red=self.gn_red[bi_, i, l],
                                      mr=self.gn_mr[bi_, i, l],
                                      dbias=self.grad_view(f"bbox_head.{br}_convs.{i}.conv.bias"), N=B, HW=h * w,
                                      **({"gsums": self.gn_bsum[bi_, i, l]} if fuse_gnb else {})))
            self.add_bwd(self._gn_bwd(gsegs), wait=f"head.tower{i + 2}.wgrad" if i + 2 <= 3 else None)
            for bi_, br in enumerate(br_names):
                # dgamma / dbeta: sum over levels and images of the per-(n,c) sums
                red = self.gn_red[bi_, i].reshape(nl * B, 256, 2)
                self.add_bwd(self.ew("dslb_gn_bwd_params", red, self.grad_view(f"bbox_head.{br}_convs.{i}.gn.weight"),
                                     self.grad_view(f"bbox_head.{br}_convs.{i}.gn.bias"), nl * B, 256),
                             side=True, tag="gn_params")
            wsegs = []
            for br in br_names:
                for l, (h, w) in enumerate(self.psize):
                    x = self.p[l] if i == 0 else self.z[br][i - 1][l]
                    wsegs.append(self.tower[br][i].wseg(x, self.dy[br][l], B, h, w))
            # wg_mode >= 1: issue the layer's dgrad first, so the side-stream wgrad (ordered behind the main stream's
            # position at its issue) starts when the dgrad has finished and runs under the NEXT layer's GroupNorm backward
            # instead of time-slicing the SMs with its own dgrad (8.907 vs 8.971 ms per step, three interleaved pairs)
            wgrad_late = wg_mode >= 1 and i > 0
            if not wgrad_late:
                self.plan_wgrad(wsegs, f"head.tower{i}.wgrad")
            if i > 0:
                dsegs = []
                for bi_, br in enumerate(br_names):
                    for l, (h, w) in enumerate(self.psize):
                        dsegs.append(self.tower[br][i].dseg(self.dy[br][l], self.dz[br][l], B, h, w, h, w,
                                                            **gnb(bi_, br, i - 1, l)))
                self.plan_bwd(dsegs, f"head.tower{i}.dgrad")
                if wgrad_late:
                    self.plan_wgrad(wsegs, f"head.tower{i}.wgrad")
            else:
                # both towers read the FPN output: cls-tower dgrad writes dP, reg-tower dgrad accumulates into it
                self.plan_bwd([self.tower["cls"][0].dseg(self.dy["cls"][l], self.dp[l], B, h, w, h, w)
                               for l, (h, w) in enumerate(self.psize)], "head.tower0.dgrad.cls")
                self.plan_bwd([self.tower["reg"][0].dseg(self.dy["reg"][l], self.dp[l], B, h, w, h, w,
                                                         residual=self.dp[l])
                               for l, (h, w) in enumerate(self.psize)], "head.tower0.dgrad.reg")

    def zero_state(self):
        """Clear what the backward accumulates into: the flat gradient and the zero arena (every packed wgrad accumulator,
        the GroupNorm backward sums, rc_dw / rc_db)."""
        L.zero(self.grad)
        L.zero(self.arena)

    def _gn_bwd(self, gsegs):
        n = len(gsegs)
        arr = (L.GnSeg * n)()
        keep = []
        for a, s in zip(arr, gsegs):
            for k, v in s.items():
                if isinstance(v, torch.Tensor):
                    keep.append(v)
                    setattr(a, k, v.data_ptr())
                else:
                    setattr(a, k, v)
        nb = L.lib.dslb_gn_bwd_blocks(arr, n)
        host = (C.c_int * (2 * nb))()
        L.check(L.lib.dslb_gn_bwd_plan(arr, n, host), "gn_bwd_plan")
        tab = torch.tensor(list(host), dtype=torch.int32, device=self.dev)
        keep.append(tab)

        def run(_arr=arr, _n=n, _tab=tab, _nb=nb, _k=keep):
            L.check(L.lib.dslb_gn_bwd(_arr, _n, 256, 32, 1e-5, L.ptr(_tab), _nb, L.cur_stream()), "gn_bwd")

        return run

    # ------------------------------------------------------------------------------------------ FPN backward
    def _build_fpn_bwd(self):
        B = self.B
        cs = self.stage_out[1:]
        (h5, w5), (h6, w6), (h7, w7) = self.psize[2], self.psize[3], self.psize[4]
        gb = lambda i: self.grad_view(f"{self.neck_prefix}fpn_convs.{i}.conv.bias")  # noqa: E731
        # P7 = conv_s2(relu(P6)); P6 = conv_s2(P5)
        self.tmp6 = self.buf(B, h6, w6, 256)
        self.plan_wgrad([self.fpnc[4].wseg(self.r6, self.dp[4], B, h6, w6)], "fpn.p7.wgrad")
        self.add_bwd(self.ew("dslb_colsum", self.dp[4], gb(4), B * h7 * w7, 256, 256), side=True, tag="colsum")
        # dgrad of the two stride-2 3x3 convs: zero-upsample dY, then a stride-1 tensor-core dgrad
        self.up7 = self.buf(B, h6, w6, 256)
        self.up6 = self.buf(B, h5, w5, 256)
        self.add_bwd(self.ew("dslb_zero_upsample2", self.dp[4], self.up7, B, h7, w7, h6, w6, 256))
        self.plan_bwd([self.fpnc[4].dseg_upsampled(self.up7, self.tmp6, B, h6, w6, relu_mask=self.p[3])],
                      "fpn.p7.dgrad")
        self.add_bwd(self.ew("dslb_relu_family", self.dp[3], self.tmp6, self.dp[3], self.tmp6.numel(), 2))
        self.plan_wgrad([self.fpnc[3].wseg(self.p[2], self.dp[3], B, h5, w5)], "fpn.p6.wgrad")
        self.add_bwd(self.ew("dslb_colsum", self.dp[3], gb(3), B * h6 * w6, 256, 256), side=True, tag="colsum")
        self.add_bwd(self.ew("dslb_zero_upsample2", self.dp[3], self.up6, B, h6, w6, h5, w5, 256))
        self.plan_bwd([self.fpnc[3].dseg_upsampled(self.up6, self.dp[2], B, h5, w5, residual=self.dp[2])],
                      "fpn.p6.dgrad")
        # P3..P5 = conv3x3(merged laterals)
        self.plan_wgrad([self.fpnc[i].wseg(self.lm[i], self.dp[i], B, cs[i][1], cs[i][2]) for i in range(3)],
                        "fpn.out.wgrad")
        for i in range(3):
            self.add_bwd(self.ew("dslb_colsum", self.dp[i], gb(i), B * cs[i][1] * cs[i][2], 256, 256), side=True, tag="colsum")
        self.dl = [self.buf(B, cs[i][1], cs[i][2], 256) for i in range(3)]
        self.plan_bwd([self.fpnc[i].dseg(self.dp[i], self.dl[i], B, cs[i][1], cs[i][2], cs[i][1], cs[i][2])
                       for i in range(3)], "fpn.out.dgrad")
        # top-down path backward: d l4m += down(d l3m); d l5m += down(d l4m)
        for i in (1, 2):
            (_, h, w, _), (_, hs, ws, _) = cs[i - 1], cs[i]
            self.add_bwd(self.ew("dslb_upsample_add_bwd", self.dl[i], self.dl[i - 1], B, h, w, hs, ws, 256))
        # laterals
        self.plan_wgrad([self.lat[i].wseg(cs[i][0], self.dl[i], B, cs[i][1], cs[i][2]) for i in range(3)],
                        "fpn.lateral.wgrad")
        for i in range(3):
            self.add_bwd(self.ew("dslb_colsum", self.dl[i], self.grad_view(f"{self.neck_prefix}lateral_convs.{i}.conv.bias"),
                                 B * cs[i][1] * cs[i][2], 256, 256), side=True, tag="colsum")
        # gradient w.r.t. the stage outputs C3, C4 (unmasked: more consumers follow) and C5 (masked: last consumer)
        self.gc = [self.buf(B, cs[i][1], cs[i][2], cs[i][3]) for i in range(3)]
        segs = []
        for i in range(3):
            # (fused plan only: C5 has no other consumer, so its ReLU mask rides on this epilogue; a standalone FPN must
            # return the plain gradient w.r.t. its input)
            kw = dict(relu_mask=cs[i][0]) if (i == 2 and self.parts == "all") else {}
            segs.append(self.lat[i].dseg(self.dl[i], self.gc[i], B, cs[i][1], cs[i][2], cs[i][1], cs[i][2], **kw))
        self.plan_bwd(segs, "fpn.lateral.dgrad")

    # ------------------------------------------------------------------------------------------ backbone backward
    def _build_backbone_bwd(self):
        if self.backbone == "rla":
            from . import engine_rla
            return engine_rla.build_backbone_bwd(self)
        B = self.B
        stage_last = {}
        for idx, blk in enumerate(self.blocks):
            stage_last[blk["li"]] = idx
        # M = masked gradient w.r.t. each trainable block's output; stage outputs C3..C5 start from the FPN's dgrad
        for idx in range(len(self.blocks) - 1, -1, -1):
            blk = self.blocks[idx]
            if not blk["trainable"]:
                break
            li, bi = blk["li"], blk["bi"]
            h, w, hin, win, planes = blk["h"], blk["w"], blk["hin"], blk["win"], blk["planes"]
            name = f"layer{li + 1}.{bi}"
            if idx == stage_last[li] and li == 2:
                self._emit_bucket((self.bb_prefix + "layer4.",))   # 60 MB of gradients are final once layer4 is done
            if idx == stage_last[li] and li == 1 and self.parts == "all":
                # layer3 (28 MB) is final before layer2's backward starts: only layer2's 5 MB travel after the backward's
                # end (N=8 timeline, profiles/r02_timeline_n8.txt: the last bucket's all-reduce is the exposed one)
                self._emit_bucket((self.bb_prefix + "layer3.",))
            if idx == stage_last[li]:
                if li == 3:
                    blk["M"] = self.gc[2]  # already masked by the lateral dgrad epilogue
                else:
                    # C3 / C4: lateral dgrad + the next stage's strided dgrads have accumulated; apply the ReLU mask
                    g = self.gc[li - 1]
                    self.add_bwd(self.ew("dslb_relu_family", g, blk["out"], g, g.numel(), 1))
                    blk["M"] = g
            M = blk["M"]
            blk["da2"] = self.buf(B, h, w, planes)
            blk["da1"] = self.buf(B, h, w, planes)
            self.plan_bwd([blk["c3"].dseg(M, blk["da2"], B, h, w, h, w, relu_mask=blk["a2"])], name + ".conv3.dgrad")
            self.plan_bwd([blk["c2"].dseg(blk["da2"], blk["da1"], B, h, w, h, w, relu_mask=blk["a1"])],
                          name + ".conv2.dgrad")
            wsegs = [blk["c3"].wseg(blk["a2"], M, B, h, w), blk["c2"].wseg(blk["a1"], blk["da2"], B, h, w),
                     blk["c1"].wseg(blk["xin"], blk["da1"], B, hin, win)]
            if bi == 0:
                wsegs.append(blk["ds"].wseg(blk["xin"], M, B, hin, win))
            self.plan_wgrad(wsegs, name + ".wgrad")
            if not blk["dgrad_in"]:
                continue
            prev = self.blocks[idx - 1]
            if bi > 0:
                # M_prev = (dgrad_conv1(da1) + M) * [xin > 0]   (identity shortcut fused as the residual operand)
                prev["M"] = self.buf(B, hin, win, blk["cin"])
                self.plan_bwd([blk["c1"].dseg(blk["da1"], prev["M"], B, h, w, hin, win, residual=M,
                                              relu_mask=blk["xin"])], name + ".conv1.dgrad")
            else:
                # first block of layer3 / layer4: both strided 1x1 dgrads scatter-accumulate into the stage-input
                # gradient that the FPN lateral dgrad has already initialised
                g = self.gc[li - 2]
                self.plan_bwd([blk["ds"].dseg(M, g, B, h, w, hin, win, residual=g)], name + ".downsample.dgrad")
                self.plan_bwd([blk["c1"].dseg(blk["da1"], g, B, h, w, hin, win, residual=g)], name + ".conv1.dgrad")
    def _emit_bucket(self, prefixes, head=False):
        """Close a gradient bucket: unpack the packed weight gradients of the convs named by `prefixes` (one launch,
        after every side-stream op issued so far) and record the flat-gradient range that is final from here on, so a
        data-parallel trainer can start its all-reduce while the rest of the backward still runs."""
        convs = [c for c in self.convs if c.trainable and c.wname.startswith(prefixes) and c.wname not in self._unpacked]
        descs = [c.unpack_desc() for c in convs]
        names = [c.wname for c in convs]
        if head:
            descs.append(dict(dw=self.rc_dw, g=self.grad_view("bbox_head.conv_reg.weight"), O=4, I=256, R=3, S=3, rows=5,
                              row_off=0))
            descs.append(dict(dw=self.rc_dw, g=self.grad_view("bbox_head.conv_centerness.weight"), O=1, I=256, R=3, S=3,
                              rows=5, row_off=4))
        self._unpacked.update(names)
        # trainable BatchNorms folded into those convs (RLA_ResNet): dgamma from the packed weight gradients, one launch
        bnd = [(n, d) for n, d in getattr(self, "bn_grad_descs", []) if n.startswith(prefixes) and n not in self._unpacked]
        if bnd:
            from .engine_rla import BnGradPlan, bn_grad_desc
            self._unpacked.update(n for n, _ in bnd)
            bplan = BnGradPlan([bn_grad_desc(self, n, convs) for n, convs in bnd])
            self.bn_grad_plans = getattr(self, "bn_grad_plans", []) + [bplan]
            self.add_bwd(bplan.run, side=True, tag="bn_grads")
        plan = TablePlan(descs, "unpack", "unpack_wgrads")
        self.unpack_plans = getattr(self, "unpack_plans", []) + [plan]
        # on the side stream, in order behind the weight / bias gradient launches it reads: the main (dgrad) stream never
        # stalls at a bucket boundary; backward() joins the two streams at the end of every op range
        self.add_bwd(plan.run, side=True, tag="unpack")
        if head:
            o_reg = self.store.offsets["bbox_head.conv_reg.bias"][0]
            o_ctr = self.store.offsets["bbox_head.conv_centerness.bias"][0]
            self.head_bias_idx = torch.tensor([o_reg, o_reg + 1, o_reg + 2, o_reg + 3, o_ctr], dtype=torch.long,
                                              device=self.dev)

            def finish_head_grads():
                s = L.cur_stream()
                L.check(L.lib.dslb_scatter_f32(L.ptr(self.grad), L.ptr(self.head_bias_idx), L.ptr(self.rc_db), 5, s),
                        "head bias grads")
                L.check(L.lib.dslb_scatter_f32(L.ptr(self.grad), L.ptr(self.scale_idx), L.ptr(self.dscale),
                                               len(self.psize), s), "scale grads")

            self.add_bwd(finish_head_grads, side=True, tag="finish_head")
        # last side op of the bucket: everything in its gradient range is final here -> `bucket_hook(k, lo, hi)` (a
        # single-GPU trainer takes the bucket's share of the gradient norm off the critical path)
        self.add_bwd(lambda k=len(self.bwd_buckets): self._bucket_done(k), side=True, tag="bucket_done")
        offs = [self.store.offsets[n] for n in names]
        lo = min(o for o, _ in offs)
        hi = max(o + n for o, n in offs)
        if head:
            hi = self.store.n_train          # head / neck biases (region B) and the Scale parameters travel with it
        self.bwd_buckets.append((len(self.bwd_ops), lo, hi))

    def _bucket_done(self, k):
        hook = getattr(self, "bucket_hook", None)
        if hook is not None:
            _, lo, hi = self.bwd_buckets[k]
            hook(k, lo, hi)

    def _build_finish_bwd(self):
        # the remaining packed wgrads -> their OIHW gradient views (standalone head: everything, in one launch)
        if self.parts == "head":
            self._emit_bucket(("bbox_head.",), head=True)
            self.bwd_buckets[-1] = (len(self.bwd_ops), 0, self.store.n_train)
            return
        if self.parts == "neck":
            self._emit_bucket(("lateral_convs.", "fpn_convs."))
            self.bwd_buckets[-1] = (len(self.bwd_ops), 0, self.store.n_train)
            return
        self._emit_bucket((self.bb_prefix,))
        if self.parts == "backbone":
            return
        # the buckets must tile the trainable range exactly: [0, layer4) | [layer4, neck) | [neck, n_train)
        rng = sorted((lo, hi) for _, lo, hi in self.bwd_buckets)
        assert rng[0][0] == 0 and rng[-1][1] == self.store.n_train and all(a[1] == b[0] for a, b in zip(rng, rng[1:])), rng

    # ------------------------------------------------------------------------------------------ run
    def _build_pack_plans(self):
        """Two multi-tensor launches: `pack_all` refreshes every derived operand, `pack_train` only those whose master
        weights an optimizer step can change (the student never needs more after construction)."""
        all_d, train_d = [], []
        for c in self.convs:
            ds = c.pack_descs()
            all_d += ds
            if c.trainable:
                train_d += ds
        for tr, d in self.extra_pack_descs:
            all_d.append(d)
            if tr:
                train_d.append(d)
        self.pack_all = TablePlan(all_d, "pack", "pack_all") if all_d else None
        self.pack_train = TablePlan(train_d, "pack", "pack_train") if train_d else None

    def repack(self, everything=True):
        """Refresh the derived bf16 operands from the fp32 master parameters (after an optimizer / EMA step)."""
        if everything or self.pack_train is None:
            if self.pack_all is not None:
                self.pack_all.run()
        else:
            self.pack_train.run()
        for f in self.repack_ops:
            f()

    def forward(self, start=0, end=None):
        for op in self.fwd_ops[start:end]:
            op()

    def forward_head(self):
        """FCOSHead only, on whatever is in the FPN output buffers self.p[l] (NHWC bf16)."""
        for op in self.fwd_ops[self.head_op_start:]:
            op()

    def backward(self, side_stream=None, start=0, end=None, join=True):
        """Run backward ops [start, end) (default: all). With `side_stream` the weight-gradient launches (off the dgrad
        critical path) go to that stream, ordered by events, so they fill the SMs the small deep-layer dgrad grids
        leave idle. A partial range ends with the main stream joined to the side stream (bucket boundaries of
        `bwd_buckets`: everything in the bucket's gradient range is final when the call returns). join=False (ranges
        of ONE stream-ordered sequence / ONE captured graph only): no join — the call returns the side stream's last
        event, the bucket is final once that event AND the main stream's position have been reached, and the dgrad chain
        runs on exactly as it does in an un-bucketed backward; the last range of the step must join."""
        end = len(self.bwd_ops) if end is None else end
        if side_stream is None:
            for op in self.bwd_ops[start:end]:
                op()
            return
        main = torch.cuda.current_stream()
        if not hasattr(self, "_bwd_events"):
            self._bwd_events = [(torch.cuda.Event(), torch.cuda.Event()) if m[0] else None for m in self.bwd_meta]
        # every earlier range ended joined to the side stream, so no event of it needs (or, under graph capture, may) be
        # waited on again
        if start == 0 or not getattr(self, "_bwd_open", False):
            self._bwd_done, self._bwd_last = {}, None
        self._bwd_open = not join   # an un-joined range leaves its side events live for the next range of the step
        done = self._bwd_done
        for i in range(start, end):
            op, (side, tag, wait), evs = self.bwd_ops[i], self.bwd_meta[i], self._bwd_events[i]
            if wait is not None:
                ev = self._bwd_last if wait == "__all__" else done.get(wait)
                if ev is not None:
                    main.wait_event(ev)
            if side:
                ev_main, ev_side = evs
                ev_main.record(main)
                side_stream.wait_event(ev_main)
                with torch.cuda.stream(side_stream):
                    op()
                    ev_side.record(side_stream)
                done[tag] = self._bwd_last = ev_side
            else:
                op()
        if not join:
            return self._bwd_last
        if self._bwd_last is not None:
            main.wait_event(self._bwd_last)   # join: the range's gradients are final (and a captured graph re-joins its fork)
        return None

    def losses(self):
        """dict of the reference's loss names -> 0-dim fp32 tensors (device)."""
        s = self.loss_f32
        out = dict(loss_cls=s[0], loss_bbox=s[1], loss_centerness=s[2])
        if self.si_weight != 0.0:
            out["loss_sisoft"] = s[3]
        return out
