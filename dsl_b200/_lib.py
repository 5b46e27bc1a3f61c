"""ctypes binding of libdslb.so (C ABI in include/dslb.h).

The library is the product: there is NO Python/torch fallback for any op it exports. If the shared object is
missing the import of this module raises, loudly, with the build command.
"""
import ctypes as C
import os

import torch  # noqa: F401  (loads libcudart.so.12 before libdslb.so so both share one runtime)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DSLB_LIB") or os.path.join(_HERE, "libdslb.so")  # DSLB_LIB: A/B builds of the kernels

MAX_SEGS = 10


class DslbError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise DslbError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C dsl_b200/csrc`). dsl_b200 has no CPU/eager fallback.")
    return C.CDLL(LIB_PATH)


lib = _load()
ABI_VERSION = 103      # what include/dslb.h declares; an older build lacks entry points this package binds
lib.dslb_version.restype = C.c_int
if lib.dslb_version() < ABI_VERSION:
    raise DslbError(f"{LIB_PATH} is version {lib.dslb_version()}, include/dslb.h is {ABI_VERSION}: rebuild it "
                    "(`python -c 'import __graft_entry__ as g; g.build()'`)")


class ConvSeg(C.Structure):
    """dslb_conv_seg_t"""
    _fields_ = [
        ("x", C.c_void_p), ("w", C.c_void_p), ("y", C.c_void_p), ("residual", C.c_void_p),
        ("relu_mask", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p), ("gn_stats", C.c_void_p),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
        ("Cout", C.c_int32), ("cout_pad", C.c_int32),
        ("R", C.c_int32), ("S", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
        ("ldc", C.c_int32), ("out_fp32", C.c_int32), ("relu_nch", C.c_int32), ("gn_cpg", C.c_int32),
        ("scatter2", C.c_int32), ("Hs", C.c_int32), ("Ws", C.c_int32),
        ("gnb_x", C.c_void_p), ("gnb_mr", C.c_void_p), ("gnb_gamma", C.c_void_p), ("gnb_beta", C.c_void_p),
        ("gnb_sums", C.c_void_p),
    ]


class WgradSeg(C.Structure):
    """dslb_wgrad_seg_t"""
    _fields_ = [
        ("x", C.c_void_p), ("dy", C.c_void_p), ("dw", C.c_void_p),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
        ("Cout", C.c_int32), ("ldy", C.c_int32), ("dw_rows", C.c_int32),
        ("R", C.c_int32), ("S", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
    ]


class GnSeg(C.Structure):
    """dslb_gn_seg_t"""
    _fields_ = [
        ("x", C.c_void_p), ("y", C.c_void_p), ("dz", C.c_void_p), ("stats", C.c_void_p),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("red", C.c_void_p), ("dbias", C.c_void_p),
        ("mr", C.c_void_p),
        ("N", C.c_int32), ("HW", C.c_int32),
        ("gsums", C.c_void_p),
    ]


class PackDesc(C.Structure):
    """dslb_pack_desc_t"""
    _fields_ = [
        ("w", C.c_void_p), ("out", C.c_void_p), ("bn_gamma", C.c_void_p), ("bn_beta", C.c_void_p),
        ("bn_mean", C.c_void_p), ("bn_var", C.c_void_p), ("scale_out", C.c_void_p), ("shift_out", C.c_void_p),
        ("O", C.c_int32), ("I", C.c_int32), ("R", C.c_int32), ("S", C.c_int32),
        ("rows_pad", C.c_int32), ("cols_pad", C.c_int32), ("row_off", C.c_int32), ("col_off", C.c_int32),
        ("mode", C.c_int32), ("fill_padding", C.c_int32), ("bn_eps", C.c_float), ("w_ld", C.c_int32),
    ]


class UnpackDesc(C.Structure):
    """dslb_unpack_desc_t"""
    _fields_ = [
        ("dw", C.c_void_p), ("g", C.c_void_p), ("bn_gamma", C.c_void_p), ("bn_var", C.c_void_p),
        ("O", C.c_int32), ("I", C.c_int32), ("R", C.c_int32), ("S", C.c_int32),
        ("rows", C.c_int32), ("row_off", C.c_int32), ("bn_eps", C.c_float), ("dw_ld", C.c_int32),
        ("g_ld", C.c_int32),
    ]


class BnGradDesc(C.Structure):
    """dslb_bn_grad_desc_t"""
    _fields_ = [
        ("dw0", C.c_void_p), ("w0", C.c_void_p), ("dw1", C.c_void_p), ("w1", C.c_void_p),
        ("mean", C.c_void_p), ("var", C.c_void_p), ("dbeta", C.c_void_p), ("dgamma", C.c_void_p),
        ("O", C.c_int32), ("R", C.c_int32), ("S", C.c_int32),
        ("I0", C.c_int32), ("dw_ld0", C.c_int32), ("w_ld0", C.c_int32), ("rows0", C.c_int32),
        ("I1", C.c_int32), ("dw_ld1", C.c_int32), ("w_ld1", C.c_int32), ("rows1", C.c_int32),
        ("bn_eps", C.c_float),
    ]


class FcosLevel(C.Structure):
    """dslb_fcos_level_t"""
    _fields_ = [
        ("cls", C.c_void_p), ("regctr", C.c_void_p), ("dcls_bf16", C.c_void_p), ("dcls_f32", C.c_void_p),
        ("dregctr_bf16", C.c_void_p), ("dregctr_f32", C.c_void_p),
        ("h", C.c_int32), ("w", C.c_int32), ("stride", C.c_int32),
        ("ld_cls", C.c_int32), ("ld_dcls", C.c_int32), ("ld_dreg", C.c_int32),
        ("rr_lo", C.c_float), ("rr_hi", C.c_float), ("scale", C.c_float), ("cs_radius", C.c_float),
    ]


lib.dslb_last_error.restype = C.c_char_p
lib.dslb_version.restype = C.c_int


def _proto(name, restype, *argtypes):
    fn = getattr(lib, name, None)
    if fn is None:
        return None
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


_proto("dslb_conv_plan_create", C.c_int, C.POINTER(ConvSeg), C.c_int, C.POINTER(C.c_void_p))
_proto("dslb_conv_plan_run", C.c_int, C.c_void_p, C.c_void_p)
_proto("dslb_conv_plan_destroy", None, C.c_void_p)
_proto("dslb_conv_plan_flops", C.c_double, C.c_void_p)
_proto("dslb_wgrad_plan_create", C.c_int, C.POINTER(WgradSeg), C.c_int, C.POINTER(C.c_void_p))
_proto("dslb_wgrad_plan_run", C.c_int, C.c_void_p, C.c_void_p)
_proto("dslb_wgrad_plan_destroy", None, C.c_void_p)
_proto("dslb_wgrad_plan_flops", C.c_double, C.c_void_p)

VP, I, LL, F, D = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double
_proto("dslb_nchw_to_nhwc_bf16", I, VP, VP, I, I, I, I, I, VP)
_proto("dslb_nhwc_to_nchw_f32", I, VP, VP, I, I, I, I, I, I, VP)
_proto("dslb_stem_conv", I, VP, VP, VP, VP, VP, VP, F, VP, VP, I, I, I, VP)
_proto("dslb_maxpool3x3s2", I, VP, VP, I, I, I, I, VP)
_proto("dslb_si_half_image", I, VP, VP, I, I, I, VP)
_proto("dslb_upsample_add", I, VP, VP, I, I, I, I, I, I, VP)
_proto("dslb_upsample_add_bwd", I, VP, VP, I, I, I, I, I, I, VP)
_proto("dslb_relu_family", I, VP, VP, VP, LL, I, VP)
_proto("dslb_gn_apply_relu", I, C.POINTER(GnSeg), I, I, I, F, VP)
_proto("dslb_gn_apply_relu_tab", I, C.POINTER(GnSeg), I, I, I, F, VP, I, VP)
_proto("dslb_gn_apply_relu_split", I, C.POINTER(GnSeg), I, I, I, F, VP)
_proto("dslb_bf16_to_split", I, VP, VP, LL, I, VP)
_proto("dslb_gn_bwd_blocks", I, C.POINTER(GnSeg), I)
_proto("dslb_gn_bwd_plan", I, C.POINTER(GnSeg), I, VP)
_proto("dslb_gn_bwd", I, C.POINTER(GnSeg), I, I, I, F, VP, I, VP)
_proto("dslb_gn_bwd_params", I, VP, VP, VP, I, I, VP)
_proto("dslb_pack_weight", I, VP, VP, I, I, I, I, I, I, VP, I, VP)
_proto("dslb_unpack_wgrad", I, VP, VP, I, I, I, I, I, VP, I, VP)
_proto("dslb_bn_fold", I, VP, VP, VP, VP, F, VP, VP, I, VP)
_proto("dslb_pack_plan_create", I, C.POINTER(PackDesc), I, C.POINTER(C.c_void_p))
_proto("dslb_unpack_plan_create", I, C.POINTER(UnpackDesc), I, C.POINTER(C.c_void_p))
_proto("dslb_table_plan_run", I, VP, VP)
_proto("dslb_table_plan_destroy", None, VP)
_proto("dslb_rla_state_fwd", I, VP, VP, VP, VP, VP, VP, F, VP, I, I, I, I, VP)
_proto("dslb_rla_state_bwd", I, VP, VP, VP, VP, VP, VP, VP, F, VP, VP, VP, VP, I, I, I, I, VP)
_proto("dslb_bn_grad_plan_create", I, C.POINTER(BnGradDesc), I, C.POINTER(C.c_void_p))
_proto("dslb_bn_grad_plan_run", I, VP, VP)
_proto("dslb_bn_grad_plan_destroy", None, VP)
_proto("dslb_fcos_regctr_affine", I, VP, I, VP, VP, VP, VP, VP, VP, I, VP)
_proto("dslb_zero_upsample2", I, VP, VP, I, I, I, I, I, I, VP)
_proto("dslb_colsum", I, VP, VP, LL, I, I, VP)
_proto("dslb_zero", I, VP, C.c_size_t, VP)
_proto("dslb_scatter_f32", I, VP, VP, VP, I, VP)
_proto("dslb_f64_to_f32", I, VP, VP, I, VP)
_proto("dslb_fcos_targets", I, C.POINTER(FcosLevel), I, I, I, VP, VP, VP, VP, VP, I, I, F, I, VP, VP, VP, VP, VP, VP)
_proto("dslb_fcos_norm", I, VP, F, VP, VP)
_proto("dslb_fcos_loss", I, C.POINTER(FcosLevel), I, I, I, VP, VP, VP, VP, VP, F, F, F, I, F, VP, VP, VP, VP)
_proto("dslb_ema_update", I, VP, VP, LL, F, F, VP)
_proto("dslb_sq_norm", I, VP, LL, VP, VP)
_proto("dslb_clip_coef", I, VP, F, VP, VP)
_proto("dslb_sgd_step", I, VP, VP, VP, LL, VP, VP, F, F, F, I, VP)
_proto("dslb_sgd_ema_step", I, VP, VP, VP, LL, VP, VP, F, F, F, I, VP, F, F, VP)
_proto("dslb_fcos_point_scores", I, VP, VP, VP, LL, I, I, VP)
_proto("dslb_fcos_topk_points", I, VP, VP, VP, VP, I, I, VP)
_proto("dslb_fcos_decode_gate", I, VP, VP, VP, I, I, I, I, I, I, I, VP, VP, F, I, VP, VP, VP, VP, VP, I, VP, VP)

lib.dslb_nms_workspace_bytes.restype = C.c_size_t
lib.dslb_nms_workspace_bytes.argtypes = [I, I]
_proto("dslb_multiclass_nms", I, VP, VP, VP, VP, VP, I, I, I, F, I, VP, C.c_size_t, VP, VP, VP, VP)
_proto("dslb_pseudo_labels", I, VP, VP, VP, VP, VP, I, I, I, D, F, D, I, VP, VP, VP, VP, VP, VP)
_proto("dslb_pseudo_labels_stats", I, VP, VP, VP, VP, VP, I, I, I, D, F, D, I, VP, VP, VP, VP, VP, VP, VP, VP, VP)
_proto("dslb_pseudo_labels_saved", I, VP, VP, VP, VP, VP, I, I, I, D, F, D, I, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP)
_proto("dslb_sigmoid_focal_loss", I, VP, VP, VP, LL, I, F, F, VP, VP, VP, VP)
_proto("dslb_giou_loss", I, VP, VP, VP, LL, F, VP, VP, VP, VP)
_proto("dslb_bce_with_logits", I, VP, VP, VP, LL, VP, VP, VP, VP)
lib.dslb_view_boxes_workspace_bytes.restype = C.c_size_t
lib.dslb_view_boxes_workspace_bytes.argtypes = [I]
_proto("dslb_view_boxes", I, VP, VP, VP, VP, I, I, I, VP, C.c_size_t, VP, VP, VP, VP)
_proto("dslb_append_scaled_boxes", I, VP, VP, VP, I, F, I, VP)
_proto("dslb_pad_batch", I, VP, VP, VP, I, I, I, I, VP)
_proto("dslb_view_images", I, VP, VP, I, VP, VP, I, VP, I, I, VP)
_proto("dslb_adathres_finalize", I, VP, VP, I, D, D, D, D, D, D, VP, VP, VP, VP)

GN_STAT_STRIDE = 32


launch_count = 0  # C-ABI calls issued so far (each one launches at least one kernel of libdslb.so)


def reset_launch_count():
    global launch_count
    launch_count = 0


def check(rc, what=""):
    global launch_count
    launch_count += 1
    if rc != 0:
        msg = lib.dslb_last_error()
        raise DslbError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def zero(t):
    """t.zero_() without a framework kernel: cudaMemsetAsync on the current stream (a memset node under graph capture)."""
    check(lib.dslb_zero(ptr(t), t.numel() * t.element_size(), cur_stream()), "zero")


def cur_stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
