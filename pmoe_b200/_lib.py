"""ctypes binding of libpmoe_b200.so (the C-ABI declared in include/pmoe_b200.h).

There is no fallback: if the library is missing it is built with nvcc; if that is impossible, or a
call returns an error, a RuntimeError is raised. Nothing in this package computes on the CPU.
"""
import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpmoe_b200.so")
_lock = threading.Lock()
_lib = None

MAX_SRC = 6
MAX_SEG = 64
ACT = {None: 0, "none": 0, "relu": 1, "elu": 2, "tanh": 3, "sigmoid": 4, "relu6": 5, "hswish": 6, "hsigmoid": 7}
F32, BF16 = 0, 1


class View4(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32),
                ("sn", C.c_int64), ("sh", C.c_int64), ("sw", C.c_int64)]


class Seg(C.Structure):
    _fields_ = [("src", C.c_int8), ("dh", C.c_int8), ("dw", C.c_int8), ("reserved", C.c_int8),
                ("c0", C.c_uint16), ("nchunks", C.c_uint16)]


class ConvTc(C.Structure):
    _fields_ = [("n_src", C.c_int32), ("src", View4 * MAX_SRC), ("n_seg", C.c_int32), ("seg", Seg * MAX_SEG),
                ("ck", C.c_int32), ("ktot", C.c_int32), ("cout_pad", C.c_int32), ("wpack", C.c_void_p),
                ("out", View4), ("scale", C.c_void_p), ("shift", C.c_void_p), ("act", C.c_int32),
                ("residual", View4), ("stat_sum", C.c_void_p), ("stat_sqsum", C.c_void_p),
                ("pool_sum", C.c_void_p), ("pool_stride", C.c_int32), ("n_out_extra", C.c_int32), ("out_cols", C.c_int32),
                ("out_extra", View4 * 3), ("pool2_out", View4), ("nchw_out", C.c_void_p), ("nchw_sn", C.c_int64),
                ("nchw_sc", C.c_int64), ("nchw_sh", C.c_int64), ("nchw_sw", C.c_int64), ("nchw_c", C.c_int32),
                ("wpack_img_stride", C.c_int64), ("shift_img_stride", C.c_int64)]


def lib():
    """Load (building first if needed) the CUDA library. Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(_LIB_PATH):
                from . import build as _build
                _build.build()
            if not os.path.exists(_LIB_PATH):
                raise RuntimeError("pmoe_b200: libpmoe_b200.so is missing and could not be built; there is no fallback path")
            l = C.CDLL(_LIB_PATH)
            l.pmoe_last_error.restype = C.c_char_p
            _declare(l)
            _lib = l
    return _lib


def _declare(l):
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    l.pmoe_version.restype = i32
    l.pmoe_device_check.restype = i32
    l.pmoe_conv_tc.argtypes = [C.POINTER(ConvTc), vp]
    l.pmoe_dbg_umma_view.argtypes = [vp, i32, vp, vp, i32, i32, i32, vp]
    l.pmoe_segloss_workspace_floats.argtypes = [i32, i32]
    l.pmoe_segloss_workspace_floats.restype = C.c_size_t
    from . import _sigs
    for name, argtypes in _sigs.SIGS.items():
        fn = getattr(l, name)
        fn.argtypes = argtypes
        fn.restype = i32


def check(rc, what=""):
    if rc != 0:
        msg = lib().pmoe_last_error().decode("utf-8", "replace")
        raise RuntimeError("pmoe_b200 %s failed (%d): %s" % (what, rc, msg))


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, what="tensor"):
    if not t.is_cuda:
        raise RuntimeError("pmoe_b200: %s must live on a CUDA device (this package has no CPU path)" % what)


def view4(t, c=None):
    """Describe a (N,H,W,C) torch tensor (any strides, channel stride 1) as a PmoeView4."""
    assert t.dim() == 4 and t.stride(3) == 1, (t.shape, t.stride())
    v = View4()
    v.ptr = t.data_ptr()
    v.n, v.h, v.w = t.shape[0], t.shape[1], t.shape[2]
    v.c = t.shape[3] if c is None else c
    v.sn, v.sh, v.sw = t.stride(0), t.stride(1), t.stride(2)
    return v


def null_view():
    return View4()


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())
