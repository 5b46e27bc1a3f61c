"""ctypes binding of libdslb.so (C ABI in include/dslb.h).

The library is the product: there is NO Python/torch fallback for any op it exports. If the shared object is
missing the import of this module raises, loudly, with the build command.
"""
import ctypes as C
import os

import torch  # noqa: F401  (loads libcudart.so.12 before libdslb.so so both share one runtime)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdslb.so")

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
    ]


class WgradSeg(C.Structure):
    """dslb_wgrad_seg_t"""
    _fields_ = [
        ("x", C.c_void_p), ("dy", C.c_void_p), ("dw", C.c_void_p),
        ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
        ("Cout", C.c_int32), ("ldy", C.c_int32), ("dw_rows", C.c_int32),
        ("R", C.c_int32), ("S", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
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


def check(rc, what=""):
    if rc != 0:
        msg = lib.dslb_last_error()
        raise DslbError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def cur_stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
