"""RLA_ResNet backbone (the backbone of the shipped DSL configs, configs/fcos_semi/RLA_*.py:3-13) on the C-ABI kernels:
forward and backward execution plans for FCOSNet(backbone="rla").

Reference (paths relative to the reference root):
  RLA_Bottleneck.forward        mmdet/models/backbones/resnet_rla.py:105-137
  RLA_ResNet._forward_impl      mmdet/models/backbones/resnet_rla.py:289-327
  RLA_ResNet._freeze_stages     mmdet/models/backbones/resnet_rla.py:344-377   (what is trainable)

What one block computes (note `y = out` at :127 ALIASES the tensor the in-place `out += identity` / ReLU then rewrite, so
the reference's `y` IS the block output):
    a1  = relu(bn1(conv1(cat(x, h))))                    1x1, stride 1
    a2  = relu(bn2(conv2(a1)))                           3x3, stride s   (PyTorch style: the stride sits on conv2)
    out = relu(bn3(conv3(a2)) + (downsample(x) | x))
    h'  = recurrent_conv(tanh(bn_k(avgpool_s(h) + conv_out(out))))      conv_out / recurrent_conv shared per stage

B200 mapping. conv1 on the concatenation is two accumulating tensor-core launches — conv(h, W[:, C:]) into a1, then
conv(x, W[:, :C]) with a1 as the residual operand — so x keeps its plain [N*H*W][C] layout for every other consumer
(downsample, identity, FPN laterals, weight gradients). The 32-channel state lives in 64-channel rows (upper half zero),
which makes conv_out / recurrent_conv / the h half of conv1 regular 64-wide tensor-core tiles. BatchNorm layers are
folded into the conv operands; their affine parameters are TRAINABLE in stages 2-4 (statistics frozen), so the backward
also produces dgamma / dbeta (dslb_bn_grad_plan_*, dslb_colsum) and the operand refresh re-folds them after every step.
"""
import ctypes as C

import torch

from . import _lib as L
from .engine import BF16, ConvW, ceil_to, conv_out
from .params import RESNET_BLOCKS, RLA_CHANNEL

HLD = 64  # channel stride of every recurrent-state tensor (RLA_CHANNEL real channels + zero padding)


class SliceConvW(ConvW):
    """ConvW over an input-channel slice [i0, i0 + i_n) of an OIHW master weight, with the packed operand padded to
    i_pad input / o_pad output channels (zeros). Same interface as ConvW for pack / unpack / segment building."""

    def __init__(self, net, wname, i0=0, i_n=None, i_pad=None, o_pad=None, bn=None, bn_shift=True, stride=1, pad=0,
                 need_dgrad=False, trainable=False):
        st = net.store
        self.net = net
        self.wname = wname
        self.w = st[wname]
        self.O, self.I_total, self.R, self.S = self.w.shape
        self.i0 = i0
        self.I = self.I_total - i0 if i_n is None else i_n
        assert self.R * self.S == 1 or (i0 == 0 and self.I == self.I_total), "slices are for 1x1 convs"
        self.Ip = i_pad or ceil_to(self.I, 64)
        self.cout_pad = o_pad or ceil_to(self.O, 16)
        self.Oc = o_pad or self.O      # output channels the conv launch computes (padding rows are zero weights)
        self.stride, self.pad = stride, pad
        dev = st.device
        RS = self.R * self.S
        self.w_slice = self.w.view(self.O, -1)[:, i0 * RS:]          # data_ptr = first element of the slice
        self.wp = net.mem.zeros(RS, self.cout_pad, self.Ip, dtype=BF16)
        self.bn = bn
        self.scale = None
        self.shift = None
        if bn is not None:
            self.scale = torch.empty(self.O, dtype=torch.float32, device=dev)
            self._shift_buf = torch.empty(self.O, dtype=torch.float32, device=dev)
            self.shift = self._shift_buf if bn_shift else None      # the h half of conv1 adds no shift (the x half does)
        self.bias_name = None
        self.need_dgrad = need_dgrad
        self.trainable = trainable
        self.dy_ld = ceil_to(self.O, 64)
        if need_dgrad:
            self.wpT = net.mem.zeros(RS, self.Ip, self.dy_ld, dtype=BF16)
        self.dw = None
        if trainable:
            net.want_arena(self, "dw", RS * self.O * self.Ip, (RS, self.O, self.Ip))

    def pack_descs(self):
        st = self.net.store
        bn = {}
        if self.bn is not None:
            bn = dict(bn_gamma=st[self.bn + ".weight"], bn_beta=st[self.bn + ".bias"],
                      bn_mean=st[self.bn + ".running_mean"], bn_var=st[self.bn + ".running_var"], bn_eps=1e-5,
                      scale_out=self.scale, shift_out=self._shift_buf)
        out = [dict(w=self.w_slice, out=self.wp, O=self.O, I=self.I, R=self.R, S=self.S, rows_pad=self.cout_pad,
                    cols_pad=self.Ip, mode=0, fill_padding=1, w_ld=self.I_total, **bn)]
        if self.need_dgrad:
            out.append(dict(w=self.w_slice, out=self.wpT, O=self.O, I=self.I, R=self.R, S=self.S, rows_pad=self.Ip,
                            cols_pad=self.dy_ld, mode=1, fill_padding=1, w_ld=self.I_total, **bn))
        return out

    def unpack_desc(self):
        st = self.net.store
        RS = self.R * self.S
        g = self.net.grad_view(self.wname)[self.i0 * RS:]
        d = dict(dw=self.dw, g=g, O=self.O, I=self.I, R=self.R, S=self.S, rows=self.O, row_off=0, dw_ld=self.Ip,
                 g_ld=self.I_total)
        if self.bn is not None:
            d.update(bn_gamma=st[self.bn + ".weight"], bn_var=st[self.bn + ".running_var"], bn_eps=1e-5)
        return d

    def fseg(self, x, y, N, H, W, **kw):
        d = dict(x=x, w=self.wp, y=y, N=N, H=H, W=W, Cin=self.Ip, Cout=self.Oc, cout_pad=self.cout_pad, R=self.R,
                 S=self.S, stride=self.stride, pad=self.pad, ldc=self.Oc, shift=self.shift)
        d.update(kw)
        return d

    def dseg(self, dy, dx, N, Ho, Wo, Hin, Win, **kw):
        assert self.stride == 1
        d = dict(x=dy, w=self.wpT, y=dx, N=N, H=Ho, W=Wo, Cin=self.dy_ld, Cout=self.Ip, cout_pad=self.Ip, R=self.R,
                 S=self.S, stride=1, pad=self.R - 1 - self.pad, ldc=self.Ip)
        d.update(kw)
        return d

    def wseg(self, x, dy, N, H, W):
        return dict(x=x, dy=dy, dw=self.dw, N=N, H=H, W=W, Cin=self.Ip, Cout=self.O, ldy=self.dy_ld, dw_rows=self.O,
                    R=self.R, S=self.S, stride=self.stride, pad=self.pad)

    def bn_piece(self):
        """(dw, w slice, I, dw_ld, w_ld, rows) of this operand for dslb_bn_grad_desc_t."""
        return self.dw, self.w_slice, self.I, self.Ip, self.I_total, self.O


def _piece(c):
    if isinstance(c, SliceConvW):
        return c.bn_piece()
    return c.dw, c.w, c.I, c.I, c.I, c.O


class BnGradPlan:
    """dslb_bn_grad_plan_* wrapper: dgamma of every listed BatchNorm in one launch."""

    def __init__(self, descs):
        self.keep = []
        arr = (L.BnGradDesc * len(descs))()
        for a, d in zip(arr, descs):
            for k, v in d.items():
                if isinstance(v, torch.Tensor):
                    self.keep.append(v)
                    setattr(a, k, v.data_ptr())
                elif v is not None:
                    setattr(a, k, v)
        self.plan = C.c_void_p()
        self._lib = L.lib   # the library that owns the plan handle
        L.check(L.lib.dslb_bn_grad_plan_create(arr, len(descs), C.byref(self.plan)), "bn_grad_plan")

    def run(self):
        L.check(L.lib.dslb_bn_grad_plan_run(self.plan, L.cur_stream()), "bn_grads")

    def __del__(self):
        try:
            if self.plan:
                self._lib.dslb_bn_grad_plan_destroy(self.plan)
        except Exception:
            pass


def bn_grad_desc(net, bn, convs):
    """Descriptor of one trainable BatchNorm folded into `convs` (one ConvW, or the two halves of an RLA conv1)."""
    st = net.store
    assert all(c.dw is not None for c in convs), "weight-gradient accumulators are not allocated yet"
    d = dict(mean=st[bn + ".running_mean"], var=st[bn + ".running_var"], dbeta=net.grad_view(bn + ".bias"),
             dgamma=net.grad_view(bn + ".weight"), O=convs[0].O, R=convs[0].R, S=convs[0].S, bn_eps=1e-5)
    for k, c in enumerate(convs):
        dw, w, I, dw_ld, w_ld, rows = _piece(c)
        d.update({f"dw{k}": dw, f"w{k}": w, f"I{k}": I, f"dw_ld{k}": dw_ld, f"w_ld{k}": w_ld, f"rows{k}": rows})
    return d


def _real_flops(net, plan, over, bwd):
    """The plans count 2*MACs on the channel-padded operands; take the zero-padding part (`over`) out again so that the
    reported FLOPs stay algorithmic (32-channel state, not its 64-channel storage)."""
    plan.flops -= over
    if bwd:
        net.flops_bwd -= over
    else:
        net.flops_fwd -= over


def stage_prefixes(net, li):
    p = net.bb_prefix
    return (f"{p}stages.{li}.", f"{p}stage_bns.{li}.", f"{p}conv_outs.{li}.", f"{p}recurrent_convs.{li}.")


# ------------------------------------------------------------------------------------------------------ forward
def build_backbone(net):
    """Forward plan of RLA_ResNet._forward_impl (resnet_rla.py:289-327); fills net.blocks / net.stage_out."""
    B, H, W = net.B, net.H, net.W
    st = net.store
    pre = net.bb_prefix
    net.img = net.buf(B, 3, H, W, dtype=torch.float32)
    H2, W2 = conv_out(H, 7, 2, 3), conv_out(W, 7, 2, 3)
    H4, W4 = conv_out(H2, 3, 2, 1), conv_out(W2, 3, 2, 1)
    net.stem_out = net.buf(B, H2, W2, 64)
    net.x0 = net.buf(B, H4, W4, 64)
    net.img4 = net.buf(B, H, W, 4)
    net.add_fwd(net.ew("dslb_stem_conv", net.img, st[pre + "conv1.weight"], st[pre + "bn1.weight"], st[pre + "bn1.bias"],
                       st[pre + "bn1.running_mean"], st[pre + "bn1.running_var"], 1e-5, net.img4, net.stem_out, B, H,
                       W))
    net.flops_fwd += 2.0 * B * H2 * W2 * 64 * 147
    net.add_fwd(net.ew("dslb_maxpool3x3s2", net.stem_out, net.x0, B, H2, W2, 64))

    net.blocks = []
    net.stage_out = []
    net.bn_grad_descs = []     # (name of the BatchNorm, convs) of every trainable folded BatchNorm
    layers = RESNET_BLOCKS[net.depth]
    x, h, w, inpl = net.x0, H4, W4, 64
    hstate = net.buf(B, H4, W4, HLD)       # h0 = zeros (resnet_rla.py:296-300); never written
    for li, nb in enumerate(layers):
        planes = 64 * 2 ** li
        trainable = net.train and li >= 1          # frozen_stages = 1 (configs/fcos_semi/RLA_*.py:8)
        co = SliceConvW(net, f"{pre}conv_outs.{li}.weight", o_pad=HLD, need_dgrad=trainable, trainable=trainable)
        rc = SliceConvW(net, f"{pre}recurrent_convs.{li}.weight", i_pad=HLD, o_pad=HLD, pad=1, need_dgrad=trainable,
                        trainable=trainable)
        net.convs += [co, rc]
        for bi in range(nb):
            s = 2 if (bi == 0 and li > 0) else 1
            assert s == 1 or (h % 2 == 0 and w % 2 == 0), "AvgPool2d(2,2) of the state and the stride-2 conv2 must agree"
            ho, wo = conv_out(h, 3, s, 1), conv_out(w, 3, s, 1)
            p = f"{pre}stages.{li}.{bi}"
            last = li == len(layers) - 1 and bi == nb - 1     # its state update feeds nothing (outs = x only, :313)
            dgrad_in = trainable and not (li == 1 and bi == 0)
            blk = dict(li=li, bi=bi, stride=s, hin=h, win=w, h=ho, w=wo, cin=inpl, planes=planes, xin=x, hin_state=hstate,
                       trainable=trainable, dgrad_in=dgrad_in, last=last, co=co, rc=rc,
                       sbn=f"{pre}stage_bns.{li}.{bi}")
            c1x = SliceConvW(net, p + ".conv1.weight", i0=0, i_n=inpl, bn=p + ".bn1", need_dgrad=dgrad_in,
                             trainable=trainable)
            c1h = SliceConvW(net, p + ".conv1.weight", i0=inpl, i_n=RLA_CHANNEL, i_pad=HLD, bn=p + ".bn1",
                             bn_shift=False, need_dgrad=dgrad_in, trainable=trainable)
            net.convs += [c1x, c1h]
            blk["c1x"], blk["c1h"] = c1x, c1h
            blk["c2"] = net.conv(p + ".conv2.weight", bn=p + ".bn2", stride=s, pad=1, need_dgrad=trainable,
                                 trainable=trainable)
            blk["c3"] = net.conv(p + ".conv3.weight", bn=p + ".bn3", need_dgrad=trainable, trainable=trainable)
            blk["a1"] = net.buf(B, h, w, planes)
            blk["a2"] = net.buf(B, ho, wo, planes)
            blk["out"] = net.buf(B, ho, wo, planes * 4)
            segs = [c1h.fseg(hstate, blk["a1"], B, h, w)]
            if bi == 0:
                blk["ds"] = net.conv(p + ".downsample.0.weight", bn=p + ".downsample.1", stride=s, need_dgrad=dgrad_in,
                                     trainable=trainable)
                blk["idn"] = net.buf(B, ho, wo, planes * 4)
                segs.append(blk["ds"].fseg(x, blk["idn"], B, h, w))
            _real_flops(net, net.plan_fwd(segs, p + ".conv1.h"), 2.0 * B * h * w * planes * (HLD - RLA_CHANNEL), False)
            net.plan_fwd([c1x.fseg(x, blk["a1"], B, h, w, residual=blk["a1"], relu_nch=planes)], p + ".conv1.x")
            net.plan_fwd([blk["c2"].fseg(blk["a1"], blk["a2"], B, h, w, relu_nch=planes)], p + ".conv2")
            net.plan_fwd([blk["c3"].fseg(blk["a2"], blk["out"], B, ho, wo, residual=blk.get("idn", x),
                                         relu_nch=planes * 4)], p + ".conv3")
            if trainable:
                # (BatchNorm name, convs it is folded into): descriptors are built at bucket time, once the weight-
                # gradient accumulators exist (FCOSNet._alloc_arena)
                net.bn_grad_descs += [(p + ".bn1", [c1x, c1h]), (p + ".bn2", [blk["c2"]]), (p + ".bn3", [blk["c3"]])]
                if bi == 0:
                    net.bn_grad_descs.append((p + ".downsample.1", [blk["ds"]]))
            if not last:
                # RLA module update (resnet_rla.py:306-311)
                blk["yo"] = net.buf(B, ho, wo, HLD)
                blk["hb"] = net.buf(B, ho, wo, HLD)
                blk["hout"] = net.buf(B, ho, wo, HLD)
                sb = blk["sbn"]
                _real_flops(net, net.plan_fwd([co.fseg(blk["out"], blk["yo"], B, ho, wo)], p + ".conv_out"),
                            2.0 * B * ho * wo * planes * 4 * (HLD - RLA_CHANNEL), False)
                net.add_fwd(net.ew("dslb_rla_state_fwd", hstate, blk["yo"], st[sb + ".weight"], st[sb + ".bias"],
                                   st[sb + ".running_mean"], st[sb + ".running_var"], 1e-5, blk["hb"], B, ho, wo,
                                   int(s == 2)))
                _real_flops(net, net.plan_fwd([rc.fseg(blk["hb"], blk["hout"], B, ho, wo)], p + ".recurrent_conv"),
                            2.0 * B * ho * wo * 9 * (HLD * HLD - RLA_CHANNEL * RLA_CHANNEL), False)
                hstate = blk["hout"]
            net.blocks.append(blk)
            x, h, w, inpl = blk["out"], ho, wo, planes * 4
        net.stage_out.append((x, h, w, inpl))


# ------------------------------------------------------------------------------------------------------ backward
def build_backbone_bwd(net):
    """Backward plan (torch autograd of resnet_rla.py:105-137, 303-311 in the reference). On entry net.gc[0..2] hold
    the gradients w.r.t. C3..C5 from the FPN laterals (C5 already masked by its ReLU; masking is idempotent)."""
    B = net.B
    st = net.store
    stage_last = {}
    for idx, blk in enumerate(net.blocks):
        stage_last[blk["li"]] = idx
    for idx in range(len(net.blocks) - 1, -1, -1):
        blk = net.blocks[idx]
        if not blk["trainable"]:
            break
        li, bi, s = blk["li"], blk["bi"], blk["stride"]
        h, w, hin, win, planes, cin = blk["h"], blk["w"], blk["hin"], blk["win"], blk["planes"], blk["cin"]
        c1x, c1h, c2, c3, co, rc = blk["c1x"], blk["c1h"], blk["c2"], blk["c3"], blk["co"], blk["rc"]
        name = f"stages.{li}.{bi}"
        gv = net.grad_view
        if idx == stage_last[li] and li == 2:
            net._emit_bucket(stage_prefixes(net, 3))     # stage 4's gradients are final once its blocks are done
        # G: gradient w.r.t. the block output from every consumer but this block's own conv_out
        G = net.gc[li - 1] if idx == stage_last[li] else blk["G"]
        if blk["last"]:
            M = G                                         # C5: masked by the lateral dgrad epilogue / the standalone seed
        else:
            # ---- recurrent-state path: h' = rc(hb), hb = tanh(bn(pool(h) + conv_out(out)))
            dh_out = blk["dh_out"]                        # gradient w.r.t. h' (written by the next block)
            blk["d_hb"] = net.buf(B, h, w, HLD)
            blk["d_pre"] = net.buf(B, h, w, HLD)
            blk["dh_pool"] = net.buf(B, hin, win, HLD) if s == 2 else None
            _real_flops(net, net.plan_bwd([rc.dseg(dh_out, blk["d_hb"], B, h, w, h, w)], name + ".recurrent_conv.dgrad"),
                        2.0 * B * h * w * 9 * (HLD * HLD - RLA_CHANNEL * RLA_CHANNEL), True)
            sb = blk["sbn"]
            net.add_bwd(net.ew("dslb_rla_state_bwd", blk["d_hb"], blk["hb"], blk["hin_state"], blk["yo"],
                               st[sb + ".weight"], st[sb + ".running_mean"], st[sb + ".running_var"], 1e-5,
                               blk["d_pre"], blk["dh_pool"], gv(sb + ".weight"), gv(sb + ".bias"), B, h, w, int(s == 2)))
            _real_flops(net, net.plan_wgrad([rc.wseg(blk["hb"], dh_out, B, h, w),
                                             co.wseg(blk["out"], blk["d_pre"], B, h, w)], name + ".state.wgrad"),
                        2.0 * B * h * w * 9 * RLA_CHANNEL * (HLD - RLA_CHANNEL), True)
            # M = (G + conv_out^T(d_pre)) * [out > 0], in place
            _real_flops(net, net.plan_bwd([co.dseg(blk["d_pre"], G, B, h, w, h, w, residual=G, relu_mask=blk["out"])],
                                          name + ".conv_out.dgrad"),
                        2.0 * B * h * w * planes * 4 * (HLD - RLA_CHANNEL), True)
            M = G
        blk["M"] = M
        blk["da2"] = net.buf(B, h, w, planes)
        blk["da1"] = net.buf(B, hin, win, planes)
        net.plan_bwd([c3.dseg(M, blk["da2"], B, h, w, h, w, relu_mask=blk["a2"])], name + ".conv3.dgrad")
        if s == 2:
            blk["up"] = net.buf(B, hin, win, planes)
            net.add_bwd(net.ew("dslb_zero_upsample2", blk["da2"], blk["up"], B, h, w, hin, win, planes))
            net.plan_bwd([c2.dseg_upsampled(blk["up"], blk["da1"], B, hin, win, relu_mask=blk["a1"])],
                         name + ".conv2.dgrad")
        else:
            net.plan_bwd([c2.dseg(blk["da2"], blk["da1"], B, h, w, h, w, relu_mask=blk["a1"])], name + ".conv2.dgrad")
        wsegs = [c3.wseg(blk["a2"], M, B, h, w), c2.wseg(blk["a1"], blk["da2"], B, hin, win),
                 c1x.wseg(blk["xin"], blk["da1"], B, hin, win), c1h.wseg(blk["hin_state"], blk["da1"], B, hin, win)]
        if bi == 0:
            wsegs.append(blk["ds"].wseg(blk["xin"], M, B, hin, win))
        _real_flops(net, net.plan_wgrad(wsegs, name + ".wgrad"), 2.0 * B * hin * win * planes * (HLD - RLA_CHANNEL), True)
        # dbeta of the folded BatchNorms = column sums of the gradients w.r.t. their outputs
        p = f"{net.bb_prefix}{name}"
        net.add_bwd(net.ew("dslb_colsum", blk["da1"], gv(p + ".bn1.bias"), B * hin * win, planes, planes), side=True,
                    tag="colsum")
        net.add_bwd(net.ew("dslb_colsum", blk["da2"], gv(p + ".bn2.bias"), B * h * w, planes, planes), side=True,
                    tag="colsum")
        net.add_bwd(net.ew("dslb_colsum", M, gv(p + ".bn3.bias"), B * h * w, planes * 4, planes * 4), side=True,
                    tag="colsum")
        if bi == 0:
            net.add_bwd(lambda d=gv(p + ".downsample.1.bias"), s_=gv(p + ".bn3.bias"): d.copy_(s_), side=True,
                        tag="colsum")
        if not blk["dgrad_in"]:
            continue
        prev = net.blocks[idx - 1]
        # gradient w.r.t. the incoming state: conv1's h half + the state update's own path
        prev["dh_out"] = net.buf(B, hin, win, HLD)
        res = None if blk["last"] else (blk["dh_pool"] if s == 2 else blk["d_pre"])
        kw = dict(residual=res) if res is not None else {}
        _real_flops(net, net.plan_bwd([c1h.dseg(blk["da1"], prev["dh_out"], B, hin, win, hin, win, **kw)],
                                      name + ".conv1.h.dgrad"), 2.0 * B * hin * win * planes * (HLD - RLA_CHANNEL), True)
        if bi > 0:
            # unmasked: the previous block adds its conv_out path and applies its ReLU mask
            prev["G"] = net.buf(B, hin, win, cin)
            net.plan_bwd([c1x.dseg(blk["da1"], prev["G"], B, hin, win, hin, win, residual=M)], name + ".conv1.x.dgrad")
        else:
            g = net.gc[li - 2]
            net.plan_bwd([blk["ds"].dseg(M, g, B, h, w, hin, win, residual=g)], name + ".downsample.dgrad")
            net.plan_bwd([c1x.dseg(blk["da1"], g, B, hin, win, hin, win, residual=g)], name + ".conv1.x.dgrad")
