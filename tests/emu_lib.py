"""TEST INFRASTRUCTURE ONLY — a CPU stand-in for the subset of libdslb.so's C ABI (include/dslb.h) that a
FCOSNet(parts="backbone") plan calls, so the HOST-side plan logic of dsl_b200/engine.py / engine_rla.py (which buffer
feeds which launch, masks, residuals, gradient routing, operand packing with folded BatchNorm, buckets) can be executed
and checked against the oracle without a GPU. Each function implements the contract written in include/dslb.h with torch
fp32 math and bf16 rounding where the kernels round. It is never imported by the product; the CUDA kernels themselves are
checked on the GPU (`-m gpu` tests). The loss entry points are emulated THROUGH the oracle (autograd over its FCOSHead.loss
restatement), so what a full-detector run on the emulator checks is the plan around them, not the loss arithmetic.

Usage:  with emu_lib.installed(): net = FCOSNet(..., device="cpu"); net.forward(); ...; net.backward()
"""
import contextlib
import ctypes as C

import torch
import torch.nn.functional as F

BF16 = torch.bfloat16
GN_STAT_STRIDE = 32   # DSLB_GN_STAT_STRIDE


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, C.c_void_p):
        return p.value or 0
    return int(p)


def view(p, n, dtype):
    """Writable torch view of n elements at host address p."""
    a = _addr(p)
    assert a != 0
    nbytes = n * torch.empty((), dtype=dtype).element_size()
    buf = (C.c_char * nbytes).from_address(a)
    return torch.frombuffer(buf, dtype=dtype, count=n)


def _rows(p, nrows, ld, dtype):
    return view(p, nrows * ld, dtype).view(nrows, ld)


class _Plan:
    def __init__(self, kind, items):
        self.kind, self.items = kind, items


class EmuLib:
    """Attribute access returns the emulated entry point; anything not emulated raises."""

    def __init__(self):
        self.plans = {}
        self.next_handle = 1
        self.calls = []

    def __getattr__(self, name):
        raise AttributeError(f"emu_lib: {name} is not emulated")

    def _new(self, out_ref, plan):
        h = self.next_handle
        self.next_handle += 1
        self.plans[h] = plan
        out_ref._obj.value = h
        return 0

    def _get(self, h):
        return self.plans[_addr(h)]

    def dslb_last_error(self):
        return b"emulated"

    # ------------------------------------------------------------------------------------------ conv plans
    def dslb_conv_plan_create(self, segs, n, out):
        fields = [f for f, _ in type(segs[0])._fields_]
        items = [{f: getattr(segs[i], f) for f in fields} for i in range(n)]
        for s in items:
            assert s["Cin"] % 64 == 0 and s["cout_pad"] % 16 == 0 and s["cout_pad"] >= s["Cout"], s
            assert s["ldc"] >= s["Cout"]
        return self._new(out, _Plan("conv", items))

    def dslb_conv_plan_flops(self, h):
        return 0.0

    def dslb_conv_plan_destroy(self, h):
        self.plans.pop(_addr(h), None)

    def dslb_conv_plan_run(self, h, stream):
        for s in self._get(h).items:
            self._conv_seg(s)
        return 0

    def _conv_seg(self, s):
        N, H, W, Cin, Cout, R, S = s["N"], s["H"], s["W"], s["Cin"], s["Cout"], s["R"], s["S"]
        st, pad, ldc = s["stride"], s["pad"], s["ldc"]
        x = view(s["x"], N * H * W * Cin, BF16).view(N, H, W, Cin).permute(0, 3, 1, 2).float()
        w = view(s["w"], R * S * s["cout_pad"] * Cin, BF16).view(R, S, s["cout_pad"], Cin).permute(2, 3, 0, 1).float()
        acc = F.conv2d(x, w[:Cout].contiguous(), stride=st, padding=pad)
        Ho, Wo = acc.shape[2], acc.shape[3]
        v = acc.permute(0, 2, 3, 1).reshape(-1, Cout)
        if s["scale"]:
            v = v * view(s["scale"], Cout, torch.float32)
        if s["shift"]:
            v = v + view(s["shift"], Cout, torch.float32)
        npix = N * Ho * Wo
        ydt = torch.float32 if s["out_fp32"] else BF16
        if s["scatter2"]:
            Hs, Ws = s["Hs"], s["Ws"]
            ymap = view(s["y"], N * Hs * Ws * ldc, ydt).view(N, Hs, Ws, ldc)
            ysel = ymap[:, 0:2 * Ho:2, 0:2 * Wo:2, :Cout]
            if s["residual"]:
                r = view(s["residual"], N * Hs * Ws * ldc, BF16).view(N, Hs, Ws, ldc)[:, 0:2 * Ho:2, 0:2 * Wo:2, :Cout]
                v = v + r.reshape(-1, Cout).float()
            assert not s["relu_mask"] and not s["relu_nch"]
            ysel.copy_(v.view(N, Ho, Wo, Cout).to(ydt))
            return
        if s["residual"]:
            v = v + _rows(s["residual"], npix, ldc, BF16)[:, :Cout].float()
        if s["relu_nch"]:
            k = min(s["relu_nch"], Cout)
            v = torch.cat([v[:, :k].clamp_min(0), v[:, k:]], dim=1)
        if s["relu_mask"]:
            m = _rows(s["relu_mask"], npix, ldc, BF16)[:, :Cout].float()
            v = torch.where(m > 0, v, torch.zeros_like(v))
        vq = v.to(ydt)
        if s["gn_stats"]:   # (sum, sumsq) per (image, group) of the bf16-rounded output, fp64
            cpg = s["gn_cpg"]
            G = Cout // cpg
            st = view(s["gn_stats"], N * G * GN_STAT_STRIDE, torch.float64).view(N, G, GN_STAT_STRIDE)
            vg = vq.double().view(N, Ho * Wo, G, cpg)
            st[:, :, 0] += vg.sum(dim=(1, 3))
            st[:, :, 1] += (vg * vg).sum(dim=(1, 3))
        if s["gnb_x"]:   # group sums of the GroupNorm backward this output (dz) flows into, see include/dslb.h
            assert not s["gn_stats"] and not s["out_fp32"] and s["cout_pad"] == Cout
            cpg = s["gn_cpg"]
            G = Cout // cpg
            HW = Ho * Wo
            xg = view(s["gnb_x"], npix * Cout, BF16).view(N, HW, G, cpg).float()
            mr = view(s["gnb_mr"], N * G * 4, torch.float32).view(N, G, 4)
            gam = view(s["gnb_gamma"], Cout, torch.float32).view(1, 1, G, cpg)
            bet = view(s["gnb_beta"], Cout, torch.float32).view(1, 1, G, cpg)
            xh = (xg - mr[:, :, 0].view(N, 1, G, 1)) * mr[:, :, 1].view(N, 1, G, 1)
            dz = vq.float().view(N, HW, G, cpg)
            gdy = torch.where(xh * gam + bet > 0, dz, torch.zeros_like(dz)) * gam
            sums = view(s["gnb_sums"], N * G * GN_STAT_STRIDE, torch.float64).view(N, G, GN_STAT_STRIDE)
            sums[:, :, 0] += gdy.double().sum(dim=(1, 3))
            sums[:, :, 1] += (gdy * xh).double().sum(dim=(1, 3))
        _rows(s["y"], npix, ldc, ydt)[:, :Cout] = vq

    # ------------------------------------------------------------------------------------------ wgrad plans
    def dslb_wgrad_plan_create(self, segs, n, out):
        fields = [f for f, _ in type(segs[0])._fields_]
        items = [{f: getattr(segs[i], f) for f in fields} for i in range(n)]
        for s in items:
            assert s["Cin"] % 64 == 0 and s["ldy"] % 64 == 0 and s["ldy"] >= s["Cout"] and s["dw_rows"] >= s["Cout"], s
        return self._new(out, _Plan("wgrad", items))

    def dslb_wgrad_plan_flops(self, h):
        return 0.0

    def dslb_wgrad_plan_destroy(self, h):
        self.plans.pop(_addr(h), None)

    def dslb_wgrad_plan_run(self, h, stream):
        for s in self._get(h).items:
            N, H, W, Cin, Cout, R, S = s["N"], s["H"], s["W"], s["Cin"], s["Cout"], s["R"], s["S"]
            st, pad = s["stride"], s["pad"]
            Ho, Wo = (H + 2 * pad - R) // st + 1, (W + 2 * pad - S) // st + 1
            x = view(s["x"], N * H * W * Cin, BF16).view(N, H, W, Cin).permute(0, 3, 1, 2).float()
            dy_full = _rows(s["dy"], N * Ho * Wo, s["ldy"], BF16).float()
            assert float(dy_full[:, Cout:].abs().max() if s["ldy"] > Cout else 0.0) == 0.0, "dy padding must be zero"
            dy = dy_full[:, :Cout].reshape(N, Ho, Wo, Cout).permute(0, 3, 1, 2)
            g = torch.nn.grad.conv2d_weight(x, (Cout, Cin, R, S), dy.contiguous(), stride=st, padding=pad)
            dw = view(s["dw"], R * S * s["dw_rows"] * Cin, torch.float32).view(R * S, s["dw_rows"], Cin)
            dw[:, :Cout] += g.permute(2, 3, 0, 1).reshape(R * S, Cout, Cin)
        return 0

    # ------------------------------------------------------------------------------------------ pack / unpack
    def dslb_pack_plan_create(self, descs, n, out):
        fields = [f for f, _ in type(descs[0])._fields_]
        return self._new(out, _Plan("pack", [{f: getattr(descs[i], f) for f in fields} for i in range(n)]))

    def dslb_unpack_plan_create(self, descs, n, out):
        fields = [f for f, _ in type(descs[0])._fields_]
        return self._new(out, _Plan("unpack", [{f: getattr(descs[i], f) for f in fields} for i in range(n)]))

    def dslb_table_plan_destroy(self, h):
        self.plans.pop(_addr(h), None)

    @staticmethod
    def _w_slice(p, O, I, RS, ld):
        flat = view(p, (O - 1) * ld * RS + I * RS, torch.float32)
        return torch.as_strided(flat, (O, I, RS), (ld * RS, RS, 1))

    def dslb_table_plan_run(self, h, stream):
        plan = self._get(h)
        for d in plan.items:
            O, I, R, S = d["O"], d["I"], d["R"], d["S"]
            RS = R * S
            sc = torch.ones(O)
            if d["bn_gamma"]:
                sc = view(d["bn_gamma"], O, torch.float32) / torch.sqrt(view(d["bn_var"], O, torch.float32) + d["bn_eps"])
            if plan.kind == "pack":
                ld = d["w_ld"] or I
                w = self._w_slice(d["w"], O, I, RS, ld) * sc.view(O, 1, 1)          # [O][I][RS]
                rp, cp = d["rows_pad"], d["cols_pad"]
                out = view(d["out"], RS * rp * cp, BF16).view(RS, rp, cp)
                assert d["mode"] in (0, 1)
                if d["mode"] == 0:
                    blk = w.permute(2, 0, 1)                                          # [tap][o][i]
                    if d["bn_gamma"]:
                        view(d["scale_out"], O, torch.float32).copy_(sc)
                        view(d["shift_out"], O, torch.float32).copy_(
                            view(d["bn_beta"], O, torch.float32) - view(d["bn_mean"], O, torch.float32) * sc)
                else:
                    blk = w.permute(2, 1, 0).flip(0)                                  # [RS-1-tap][i][o]
                r0, c0 = d["row_off"], d["col_off"]
                if d["fill_padding"]:
                    assert r0 == 0 and c0 == 0
                    out.zero_()
                out[:, r0:r0 + blk.shape[1], c0:c0 + blk.shape[2]] = blk.to(BF16)
            else:
                ldd, ldg = d["dw_ld"] or I, d["g_ld"] or I
                dw = view(d["dw"], RS * d["rows"] * ldd, torch.float32).view(RS, d["rows"], ldd)
                g = self._w_slice(d["g"], O, I, RS, ldg)
                g.copy_((dw[:, d["row_off"]:d["row_off"] + O, :I] * sc.view(1, O, 1)).permute(1, 2, 0))
        return 0

    # ------------------------------------------------------------------------------------------ glue kernels
    def dslb_stem_conv(self, img, w, g, b, mean, var, eps, ws, out, N, H, W, stream):
        x = view(img, N * 3 * H * W, torch.float32).view(N, 3, H, W)
        wt = view(w, 64 * 147, torch.float32).view(64, 3, 7, 7)
        sc = view(g, 64, torch.float32) / torch.sqrt(view(var, 64, torch.float32) + eps)
        sh = view(b, 64, torch.float32) - view(mean, 64, torch.float32) * sc
        y = F.conv2d(x.to(BF16).float(), (wt * sc.view(64, 1, 1, 1)).to(BF16).float(), stride=2, padding=3)
        y = F.relu(y + sh.view(1, 64, 1, 1))
        Ho, Wo = y.shape[2], y.shape[3]
        view(out, N * Ho * Wo * 64, BF16).view(N, Ho, Wo, 64).copy_(y.permute(0, 2, 3, 1).to(BF16))
        return 0

    def dslb_maxpool3x3s2(self, x, y, N, H, W, Cc, stream):
        xi = view(x, N * H * W * Cc, BF16).view(N, H, W, Cc).permute(0, 3, 1, 2).float()
        yo = F.max_pool2d(xi, 3, 2, 1)
        view(y, yo.numel(), BF16).view(N, yo.shape[2], yo.shape[3], Cc).copy_(yo.permute(0, 2, 3, 1).to(BF16))
        return 0

    def dslb_relu_family(self, x, m, y, n, mode, stream):
        xv = view(x, n, BF16).float()
        if mode == 0:
            r = xv.clamp_min(0)
        elif mode == 1:
            r = torch.where(view(m, n, BF16).float() > 0, xv, torch.zeros_like(xv))
        else:
            r = xv + view(m, n, BF16).float()
        view(y, n, BF16).copy_(r.to(BF16))
        return 0

    def dslb_zero(self, p, nbytes, stream):
        if nbytes:
            view(p, int(nbytes), torch.uint8).zero_()
        return 0

    def dslb_scatter_f32(self, dst, idx, src, n, stream):
        i = view(idx, n, torch.int64)
        s_ = view(src, n, torch.float32)
        for k in range(n):
            view(_addr(dst) + 4 * int(i[k]), 1, torch.float32)[0] = s_[k]
        return 0

    def dslb_f64_to_f32(self, src, dst, n, stream):
        view(dst, n, torch.float32).copy_(view(src, n, torch.float64).to(torch.float32))
        return 0

    def dslb_colsum(self, x, out, npix, ld, Cc, stream):
        view(out, Cc, torch.float32).add_(_rows(x, npix, ld, BF16)[:, :Cc].float().sum(0))
        return 0

    def dslb_zero_upsample2(self, x, y, N, h, w, H, W, Cc, stream):
        yo = view(y, N * H * W * Cc, BF16).view(N, H, W, Cc)
        yo.zero_()
        yo[:, 0:2 * h:2, 0:2 * w:2] = view(x, N * h * w * Cc, BF16).view(N, h, w, Cc)
        return 0

    # ------------------------------------------------------------------------------------------ FPN glue
    def dslb_upsample_add(self, dst, src, N, H, W, h, w, Cc, stream):
        d = view(dst, N * H * W * Cc, BF16).view(N, H, W, Cc)
        sv = view(src, N * h * w * Cc, BF16).view(N, h, w, Cc).permute(0, 3, 1, 2).float()
        up = F.interpolate(sv, size=(H, W), mode="nearest").permute(0, 2, 3, 1)
        d.copy_((d.float() + up).to(BF16))
        return 0

    def dslb_upsample_add_bwd(self, dsrc, ddst, N, H, W, h, w, Cc, stream):
        g = view(ddst, N * H * W * Cc, BF16).view(N, H, W, Cc).permute(0, 3, 1, 2).float()
        z = torch.zeros(N, Cc, h, w, requires_grad=True)
        F.interpolate(z, size=(H, W), mode="nearest").backward(g)
        d = view(dsrc, N * h * w * Cc, BF16).view(N, h, w, Cc)
        d.copy_((d.float() + z.grad.permute(0, 2, 3, 1)).to(BF16))
        return 0

    # ------------------------------------------------------------------------------------------ GroupNorm
    def dslb_gn_bwd_blocks(self, segs, n):
        return 1

    def dslb_gn_bwd_plan(self, segs, n, host):
        return 0

    @staticmethod
    def _gn_items(segs, n):
        fields = [f for f, _ in type(segs[0])._fields_]
        return [{f: getattr(segs[i], f) for f in fields} for i in range(n)]

    @staticmethod
    def _gn_norm(d, Cc, groups, eps):
        N, HW = d["N"], d["HW"]
        cpg = Cc // groups
        x = view(d["x"], N * HW * Cc, BF16).view(N, HW, groups, cpg).float()
        st = view(d["stats"], N * groups * GN_STAT_STRIDE, torch.float64).view(N, groups, GN_STAT_STRIDE)
        cnt = float(HW * cpg)
        mean = st[:, :, 0] / cnt
        var = (st[:, :, 1] / cnt - mean * mean).clamp_min(0)
        rstd = 1.0 / torch.sqrt(var + eps)
        mean, rstd = mean.float().view(N, 1, groups, 1), rstd.float().view(N, 1, groups, 1)
        gamma = view(d["gamma"], Cc, torch.float32).view(1, 1, groups, cpg)
        beta = view(d["beta"], Cc, torch.float32).view(1, 1, groups, cpg)
        xhat = (x - mean) * rstd
        return x, xhat, gamma, beta, mean, rstd, cpg

    def dslb_gn_apply_relu_tab(self, segs, n, Cc, groups, eps, tab, nb, stream):
        for d in self._gn_items(segs, n):
            x, xhat, gamma, beta, mean, rstd, cpg = self._gn_norm(d, Cc, groups, eps)
            y = (xhat * gamma + beta).clamp_min(0)
            view(d["y"], y.numel(), BF16).copy_(y.reshape(-1).to(BF16))
            if d["mr"]:
                mr = view(d["mr"], d["N"] * groups * 4, torch.float32).view(d["N"], groups, 4)
                mr[:, :, 0] = mean.view(d["N"], groups)
                mr[:, :, 1] = rstd.view(d["N"], groups)
        return 0

    def dslb_gn_apply_relu(self, segs, n, Cc, groups, eps, stream):
        return self.dslb_gn_apply_relu_tab(segs, n, Cc, groups, eps, None, 0, stream)

    def dslb_gn_bwd(self, segs, n, Cc, groups, eps, tab, nb, stream):
        for d in self._gn_items(segs, n):
            N, HW = d["N"], d["HW"]
            x, xhat, gamma, beta, mean, rstd, cpg = self._gn_norm(d, Cc, groups, eps)
            dz = view(d["dz"], N * HW * Cc, BF16).view(N, HW, groups, cpg).float()
            dy = torch.where(xhat * gamma + beta > 0, dz, torch.zeros_like(dz))
            red = view(d["red"], N * Cc * 2, torch.float64).view(N, Cc, 2)
            red[:, :, 0] += dy.sum(dim=1).reshape(N, Cc).double()
            red[:, :, 1] += (dy * xhat).sum(dim=1).reshape(N, Cc).double()
            gdy = gamma * dy
            if d["gsums"]:   # the producing conv's epilogue left the group sums: no reduction here
                gs = view(d["gsums"], N * groups * GN_STAT_STRIDE, torch.float64).view(N, groups, GN_STAT_STRIDE)
                m1 = (gs[:, :, 0] / float(HW * cpg)).float().view(N, 1, groups, 1)
                m2 = (gs[:, :, 1] / float(HW * cpg)).float().view(N, 1, groups, 1)
            else:
                m1 = gdy.mean(dim=(1, 3), keepdim=True)
                m2 = (gdy * xhat).mean(dim=(1, 3), keepdim=True)
            dx = (rstd * (gdy - m1 - xhat * m2)).to(BF16)
            view(d["y"], dx.numel(), BF16).copy_(dx.reshape(-1))
            if d["dbias"]:
                view(d["dbias"], Cc, torch.float32).add_(dx.float().sum(dim=(0, 1)).reshape(Cc))
        return 0

    def dslb_gn_bwd_params(self, red, dgamma, dbeta, N, Cc, stream):
        r = view(red, N * Cc * 2, torch.float64).view(N, Cc, 2)
        view(dgamma, Cc, torch.float32).add_(r[:, :, 1].sum(0).float())
        view(dbeta, Cc, torch.float32).add_(r[:, :, 0].sum(0).float())
        return 0

    # ------------------------------------------------------------------------------------------ head / loss
    def dslb_fcos_regctr_affine(self, scales, stride, reg_bias, ctr_bias, level_mult, rc_scale, rc_shift, scale_vals, nl,
                                stream):
        sc = view(scales, (nl - 1) * stride + 1, torch.float32)[::stride]
        lm = view(level_mult, nl, torch.float32)
        bias5 = torch.cat([view(reg_bias, 4, torch.float32), view(ctr_bias, 1, torch.float32)])
        rs = view(rc_scale, nl * 8, torch.float32).view(nl, 8)
        rh = view(rc_shift, nl * 8, torch.float32).view(nl, 8)
        rs.fill_(1.0)
        rs[:, :4] = (sc * lm).view(nl, 1)
        rh.zero_()
        rh[:, :5] = bias5.view(1, 5) * rs[:, :5]
        view(scale_vals, nl, torch.float32).copy_(sc)
        return 0

    @staticmethod
    def _levels(levels, nl):
        fields = [f for f, _ in type(levels[0])._fields_]
        return [{f: getattr(levels[i], f) for f in fields} for i in range(nl)]

    @staticmethod
    def _box_lists(boxes, off, B, labels=None):
        o = view(off, B + 1, torch.int32).tolist()
        bx = view(boxes, max(o[-1], 1) * 4, torch.float32).view(-1, 4)
        out = [bx[o[i]:o[i + 1]].clone() for i in range(B)]
        if labels is None:
            return out
        lb = view(labels, max(o[-1], 1), torch.int64)
        return out, [lb[o[i]:o[i + 1]].clone() for i in range(B)]

    def dslb_fcos_targets(self, levels, nl, B, Cn, gt_boxes, gt_labels, gt_off, ig_boxes, ig_off, center_sampling,
                          norm_on_bbox, lw, n_labeled, labels, bbox_targets, weights, ctr_targets, counts, stream):
        """Targets through the oracle's restatement of get_targets / the ignore and unlabeled weights (the CUDA kernel is
        checked bit-exact against the same functions on the GPU)."""
        from oracle import fcos_oracle as O
        lv = self._levels(levels, nl)
        gts, gls = self._box_lists(gt_boxes, gt_off, B, gt_labels)
        igs = self._box_lists(ig_boxes, ig_off, B) if _addr(ig_boxes) else None
        strides = tuple(l["stride"] for l in lv)
        rr = tuple((l["rr_lo"], l["rr_hi"]) for l in lv)
        radius = lv[0]["cs_radius"] / lv[0]["stride"]
        self.loss_ctx = dict(gts=gts, gls=gls, igs=igs, strides=strides, rr=rr, radius=radius, Cn=Cn,
                             center_sampling=bool(center_sampling), norm_on_bbox=bool(norm_on_bbox))
        zeros = lambda c: [torch.zeros(B, c, l["h"], l["w"]) for l in lv]  # noqa: E731
        out = O.fcos_loss(zeros(Cn), zeros(4), zeros(1), gts, gls, igs, strides=strides, regress_ranges=rr,
                          num_classes=Cn, center_sampling=bool(center_sampling), radius=radius,
                          norm_on_bbox=bool(norm_on_bbox), loss_weight=lw, return_aux=True)
        aux = out["_aux"]
        P = aux["labels"].numel()
        view(labels, P, torch.int64).copy_(aux["labels"])
        view(bbox_targets, P * 4, torch.float32).view(P, 4).copy_(aux["bbox_targets"])
        view(weights, P, torch.float32).copy_(aux["weight"])
        ct = view(ctr_targets, P, torch.float32)
        ct.zero_()
        ct[aux["pos_inds"]] = aux["centerness_targets"]
        c = view(counts, 2, torch.float64)
        c[0] += aux["num_pos_local"]
        c[1] += aux["ctr_sum_local"]
        return 0

    def dslb_fcos_norm(self, counts, world, norm, stream):
        c = view(counts, 2, torch.float64)
        nv = view(norm, 2, torch.float32)
        nv[0] = max(float(c[0]) / world, 1.0)
        nv[1] = max(float(c[1]) / world, 1e-6)
        return 0

    def dslb_fcos_loss(self, levels, nl, B, Cn, labels, bbox_targets, weights, ctr_targets, norm, alpha, gamma, lw,
                       n_labeled, si_weight, level_scales, loss_sums, dscale, stream):
        """Losses and head-output gradients by autograd over the oracle's FCOSHead.loss restatement, then the chain rule
        the header documents (d/d conv_reg through relu(scale * x), d/d scale)."""
        from oracle import fcos_oracle as O
        lv, ctx = self._levels(levels, nl), self.loss_ctx
        cls, box, ctr = [], [], []
        for l in lv:
            n = B * l["h"] * l["w"]
            c = _rows(l["cls"], n, l["ld_cls"], torch.float32)[:, :Cn].reshape(B, l["h"], l["w"], Cn)
            r = _rows(l["regctr"], n, 8, torch.float32).reshape(B, l["h"], l["w"], 8)
            cls.append(c.permute(0, 3, 1, 2).clone().requires_grad_(True))
            box.append(r[..., :4].permute(0, 3, 1, 2).clone().requires_grad_(True))
            ctr.append(r[..., 4:5].permute(0, 3, 1, 2).clone().requires_grad_(True))
        nv = view(norm, 2, torch.float32)
        out = O.fcos_loss(cls, box, ctr, ctx["gts"], ctx["gls"], ctx["igs"], strides=ctx["strides"],
                          regress_ranges=ctx["rr"], num_classes=Cn, center_sampling=ctx["center_sampling"],
                          radius=ctx["radius"], norm_on_bbox=ctx["norm_on_bbox"], loss_weight=lw,
                          soft_weight=float(si_weight), soft_warm_up=-1, world_num_pos=float(nv[0]),
                          world_ctr_sum=float(nv[1]))
        sum(out.values()).backward()
        ls = view(loss_sums, 4, torch.float64)
        for i, k in enumerate(("loss_cls", "loss_bbox", "loss_centerness", "loss_sisoft")):
            if k in out:
                ls[i] += float(out[k].detach())
        sc = view(level_scales, nl, torch.float32) if _addr(level_scales) else None
        for i, l in enumerate(lv):
            n = B * l["h"] * l["w"]
            scale = float(sc[i]) if sc is not None else l["scale"]
            gcls = (cls[i].grad if cls[i].grad is not None else torch.zeros_like(cls[i])).permute(0, 2, 3, 1).reshape(n, Cn)
            gbox = (box[i].grad if box[i].grad is not None else torch.zeros_like(box[i])).permute(0, 2, 3, 1).reshape(n, 4)
            gctr = (ctr[i].grad if ctr[i].grad is not None else torch.zeros_like(ctr[i])).permute(0, 2, 3, 1).reshape(n, 1)
            bx = box[i].detach().permute(0, 2, 3, 1).reshape(n, 4)
            if l["dcls_bf16"]:
                _rows(l["dcls_bf16"], n, l["ld_dcls"], BF16)[:, :Cn] = gcls.to(BF16)
            if l["dcls_f32"]:
                _rows(l["dcls_f32"], n, Cn, torch.float32).copy_(gcls)
            live = (bx > 0).float()
            if l["dregctr_bf16"]:
                d = _rows(l["dregctr_bf16"], n, l["ld_dreg"], BF16)
                d[:, :4] = (gbox * scale * live).to(BF16)
                d[:, 4:5] = gctr.to(BF16)
                d[:, 5:8] = 0
            if l["dregctr_f32"]:
                d = _rows(l["dregctr_f32"], n, 8, torch.float32)
                d[:, :4] = gbox
                d[:, 4:5] = gctr
            if _addr(dscale):
                view(dscale, nl, torch.float32)[i] += float((gbox * live * bx / scale).sum())
        return 0

    # ------------------------------------------------------------------------------------------ RLA
    @staticmethod
    def _rla_pre(h_old, y_out, N, Ho, Wo, pool):
        y = view(y_out, N * Ho * Wo * 64, BF16).view(N, Ho, Wo, 64)[..., :32].float()
        if pool:
            h = view(h_old, N * 4 * Ho * Wo * 64, BF16).view(N, 2 * Ho, 2 * Wo, 64)[..., :32].float()
            h = h.view(N, Ho, 2, Wo, 2, 32).sum(dim=(2, 4)) * 0.25
        else:
            h = view(h_old, N * Ho * Wo * 64, BF16).view(N, Ho, Wo, 64)[..., :32].float()
        return h + y

    def dslb_rla_state_fwd(self, h_old, y_out, g, b, mean, var, eps, hb, N, Ho, Wo, pool, stream):
        f32 = lambda p: view(p, 32, torch.float32)  # noqa: E731
        pre = self._rla_pre(h_old, y_out, N, Ho, Wo, pool)
        sc = f32(g) / torch.sqrt(f32(var) + eps)
        out = torch.tanh(pre * sc + (f32(b) - f32(mean) * sc))
        view(hb, N * Ho * Wo * 64, BF16).view(N, Ho, Wo, 64)[..., :32] = out.to(BF16)
        return 0

    def dslb_rla_state_bwd(self, d_hb, hb, h_old, y_out, g, mean, var, eps, d_pre, dh_old, dgamma, dbeta, N, Ho, Wo,
                           pool, stream):
        f32 = lambda p: view(p, 32, torch.float32)  # noqa: E731
        pre = self._rla_pre(h_old, y_out, N, Ho, Wo, pool)
        t = view(hb, N * Ho * Wo * 64, BF16).view(N, Ho, Wo, 64)[..., :32].float()
        dh = view(d_hb, N * Ho * Wo * 64, BF16).view(N, Ho, Wo, 64)[..., :32].float()
        rstd = 1.0 / torch.sqrt(f32(var) + eps)
        gg = dh * (1 - t * t)
        if _addr(dgamma):
            f32(dbeta).add_(gg.sum(dim=(0, 1, 2)))
            f32(dgamma).add_((gg * (pre - f32(mean)) * rstd).sum(dim=(0, 1, 2)))
        o = (gg * f32(g) * rstd).to(BF16)
        view(d_pre, N * Ho * Wo * 64, BF16).view(N, Ho, Wo, 64)[..., :32] = o
        if pool:
            q = (o.float() * 0.25).to(BF16)
            up = q.view(N, Ho, 1, Wo, 1, 32).expand(N, Ho, 2, Wo, 2, 32).reshape(N, 2 * Ho, 2 * Wo, 32)
            view(dh_old, N * 4 * Ho * Wo * 64, BF16).view(N, 2 * Ho, 2 * Wo, 64)[..., :32] = up
        return 0

    def dslb_bn_grad_plan_create(self, descs, n, out):
        fields = [f for f, _ in type(descs[0])._fields_]
        return self._new(out, _Plan("bn", [{f: getattr(descs[i], f) for f in fields} for i in range(n)]))

    def dslb_bn_grad_plan_destroy(self, h):
        self.plans.pop(_addr(h), None)

    def dslb_bn_grad_plan_run(self, h, stream):
        for d in self._get(h).items:
            O, RS = d["O"], d["R"] * d["S"]
            dot = torch.zeros(O)
            for k in (0, 1):
                if not d[f"dw{k}"]:
                    continue
                I, ldd, ldw, rows = d[f"I{k}"], d[f"dw_ld{k}"] or d[f"I{k}"], d[f"w_ld{k}"] or d[f"I{k}"], d[f"rows{k}"]
                dw = view(d[f"dw{k}"], RS * rows * ldd, torch.float32).view(RS, rows, ldd)[:, :O, :I]
                w = self._w_slice(d[f"w{k}"], O, I, RS, ldw)
                dot += (dw.permute(1, 2, 0) * w).sum(dim=(1, 2))
            mean, var = view(d["mean"], O, torch.float32), view(d["var"], O, torch.float32)
            view(d["dgamma"], O, torch.float32).copy_(
                (dot - mean * view(d["dbeta"], O, torch.float32)) / torch.sqrt(var + d["bn_eps"]))
        return 0


@contextlib.contextmanager
def installed():
    """Swap dsl_b200._lib.lib (and the stream lookup) for the emulator inside the block."""
    from dsl_b200 import _lib as L
    real_lib, real_stream = L.lib, L.cur_stream
    emu = EmuLib()
    L.lib = emu
    L.cur_stream = lambda: None
    try:
        yield emu
    finally:
        L.lib, L.cur_stream = real_lib, real_stream
