"""ctypes binding of libcv_b200.so (include/cv_b200.h).  There is NO fallback: if the library is missing or
the device is not sm_100 the product path raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libcv_b200.so")
_lib = None


class CvError(RuntimeError):
    pass


class cv_nodes_caps(C.Structure):
    _fields_ = [("max_external", C.c_int32), ("max_contours", C.c_int32), ("max_points", C.c_int32),
                ("max_pairs", C.c_int32)]


# numpy mirrors of the C structs (include/cv_b200.h)
BOX_DTYPE = np.dtype([("xmin", "<i4"), ("ymin", "<i4"), ("xmax", "<i4"), ("ymax", "<i4"),
                      ("rxmin", "<i4"), ("rymin", "<i4"), ("rxmax", "<i4"), ("rymax", "<i4"),
                      ("flags", "<i4"), ("thresh", "<i4"), ("uid_group", "<i4"), ("reserved", "<i4")])
CONTOUR_DTYPE = np.dtype([("start_x", "<i4"), ("start_y", "<i4"), ("offset", "<i4"), ("nverts", "<i4"),
                          ("xmin", "<i4"), ("ymin", "<i4"), ("xmax", "<i4"), ("ymax", "<i4"),
                          ("a00", "<i8"), ("a01", "<i8"), ("new_id", "<i4"), ("ncomp", "<i4"),
                          ("has_source", "<i4"), ("centroid_y", "<i4")])
PAIR_DTYPE = np.dtype([("contour", "<i4"), ("box", "<i4"), ("px", "<i4"), ("py", "<i4")])
RESULT_DTYPE = np.dtype([("n_external", "<i4"), ("n_contours", "<i4"), ("n_points", "<i4"), ("n_pairs", "<i4"),
                         ("n_nodes", "<i4"), ("ground", "<i4"), ("inverted", "<i4"), ("status", "<i4")])
assert BOX_DTYPE.itemsize == 48 and CONTOUR_DTYPE.itemsize == 64 and PAIR_DTYPE.itemsize == 16
assert RESULT_DTYPE.itemsize == 32

CV_BOX_ZERO_IN_MASK, CV_BOX_IS_COMPONENT, CV_BOX_IS_SOURCE, CV_BOX_IS_TERMINAL = 1, 2, 4, 8
DEFAULT_CAPS = dict(max_external=32768, max_contours=2560, max_points=262144, max_pairs=8192)


class cv_gemm_epilogue(C.Structure):
    """include/cv_b200.h cv_gemm_epilogue (field order and types must match)."""
    _fields_ = [("bias", C.c_void_p), ("act", C.c_int), ("res_before_act", C.c_int), ("residual", C.c_void_p),
                ("ld_res", C.c_longlong), ("res_row_mod", C.c_longlong), ("out_f32", C.c_void_p), ("ld_f32", C.c_longlong),
                ("out_16", C.c_void_p), ("ld_16", C.c_longlong), ("map_mode", C.c_int), ("ws", C.c_int), ("nwx", C.c_int),
                ("nwy", C.c_int), ("H", C.c_int), ("W", C.c_int), ("cout", C.c_int), ("pool_cols", C.c_int),
                ("pool_out", C.c_void_p), ("ld_pool", C.c_longlong), ("operand_fp16", C.c_int)]


def _declare(lib):
    vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
    lib.cv_last_error.restype = C.c_char_p
    lib.cv_version.restype = C.c_char_p
    lib.cv_device_is_sm100.argtypes = [i32]
    lib.cv_last_launch_count.restype = i32
    lib.cv_nodes_resized_width.argtypes = [i32, i32]
    lib.cv_nodes_workspace_bytes.argtypes = [i32, i32, i32, C.POINTER(cv_nodes_caps)]
    lib.cv_nodes_workspace_bytes.restype = sz
    lib.cv_nodes_analyze.argtypes = [vp, i32, i32, i32, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp,
                                     C.POINTER(cv_nodes_caps), vp, sz, vp]
    lib.cv_nodes_pack.argtypes = [vp, vp, vp, vp, i32, C.POINTER(cv_nodes_caps), vp, vp, C.c_longlong, vp]
    lib.cv_terminals_workspace_bytes.argtypes = [i32, i32, i32, C.POINTER(cv_nodes_caps)]
    lib.cv_terminals_workspace_bytes.restype = sz
    lib.cv_terminals_analyze.argtypes = [vp, i32, i32, i32, vp, vp, i32, i32, vp, vp, vp, vp, vp, C.POINTER(cv_nodes_caps),
                                         vp, sz, vp]
    lib.cv_ccl_workspace_bytes.argtypes = [i32, i32, i32]
    lib.cv_ccl_workspace_bytes.restype = sz
    lib.cv_ccl_label.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, sz, vp]
    ll = C.c_longlong
    lib.cv_gemm_bf16.argtypes = [vp, ll, vp, ll, i32, i32, i32, vp, i32, vp, ll, vp, ll, vp, ll, vp]
    lib.cv_gemm_ex.argtypes = [vp, ll, vp, ll, i32, i32, i32, C.POINTER(cv_gemm_epilogue), vp]
    lib.cv_attn_set_trace.argtypes = [vp]
    lib.cv_mlp_fused.argtypes = [vp, i32, i32, vp, vp, C.c_float, vp, vp, vp, vp, i32, vp]
    lib.cv_attention_bf16.argtypes = [vp, ll, i32, i32, vp, ll, i32, i32, vp, ll, i32, i32, i32, i32, i32, i32, i32,
                                      i32, C.c_float, vp, ll, vp]
    lib.cv_profile_enable.argtypes = [i32]
    lib.cv_profile_get.argtypes = [i32, C.c_char_p, i32, C.POINTER(C.c_longlong), C.POINTER(C.c_double),
                                   C.POINTER(C.c_double)]
    for name, fn in _OPTIONAL_DECLS.items():
        if hasattr(lib, name):
            fn(getattr(lib, name))


_OPTIONAL_DECLS = {}


def load():
    """Return the loaded library or raise CvError (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CvError(f"{LIB_PATH} is missing: build it with `python -m circuitvision_b200.build` "
                      "(the B200 path has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    _declare(lib)
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        raise CvError(f"{what} failed with status {status}: {load().cv_last_error().decode()}")


def require_device(device_index: int = 0):
    import torch
    if not torch.cuda.is_available():
        raise CvError("circuitvision_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if not load().cv_device_is_sm100(device_index):
        raise CvError("circuitvision_b200 kernels are built for sm_100a (B200) only")


def profile_table():
    """[{name, launches, ms, work}] of every kernel timed since the last cv_profile_reset()."""
    lib = load()
    out = []
    for i in range(lib.cv_profile_count()):
        name = C.create_string_buffer(256)
        n, ms, w = C.c_longlong(0), C.c_double(0.0), C.c_double(0.0)
        check(lib.cv_profile_get(i, name, 256, C.byref(n), C.byref(ms), C.byref(w)), "cv_profile_get")
        out.append(dict(name=name.value.decode(), launches=n.value, ms=ms.value, work=w.value))
    return out
