"""ctypes binding of libnaf_b200.so (the C ABI declared in include/naf_b200.h).

The library is built in-tree by `naf_b200/csrc/build.py`.  There is NO fallback: if the shared
object is missing or a call fails, the Python layer raises -- it never substitutes a PyTorch or
CPU implementation.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NAF_B200_LIB") or os.path.join(_HERE, "csrc", "libnaf_b200.so")
ABI_VERSION = 4

NAF_OK, NAF_ERR_BAD_SHAPE, NAF_ERR_UNSUPPORTED, NAF_ERR_WINDOW, NAF_ERR_ALIGNMENT, NAF_ERR_NULL, NAF_ERR_CUDA = range(7)
ALGO_AUTO, ALGO_GENERIC, ALGO_CELL_SIMT, ALGO_CELL_TC, ALGO_CELL_TCWS, ALGO_CELL_TMA, ALGO_UNION_TC = range(7)
# (3 was the non-pipelined tensor-core kernel, removed in ABI v3; the value stays reserved)
ALGO_NAMES = {ALGO_AUTO: "auto", ALGO_GENERIC: "generic", ALGO_CELL_SIMT: "cell_simt", ALGO_CELL_TCWS: "cell_tcws",
              ALGO_CELL_TMA: "cell_tma", ALGO_UNION_TC: "union_tc"}

_fp = C.c_void_p  # device pointers travel as integers


class KPoolParams(C.Structure):
    """Mirror of `naf_kpool_params` (include/naf_b200.h)."""

    _fields_ = [
        ("x", _fp), ("k_out", _fp), ("q_out", _fp),
        ("cos_y", _fp), ("sin_y", _fp), ("cos_x", _fp), ("sin_x", _fp),
        ("B", C.c_int32), ("D", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("h", C.c_int32), ("w", C.c_int32), ("rope_heads", C.c_int32),
        ("x_stride_b", C.c_int64), ("x_stride_y", C.c_int64), ("x_stride_x", C.c_int64),
        ("rep_y", C.c_int32), ("rep_x", C.c_int32),
    ]

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.rep_y = self.rep_x = 1


class XAttnParams(C.Structure):
    """Mirror of `naf_xattn_params` (include/naf_b200.h)."""

    _fields_ = [
        ("q", _fp), ("k", _fp), ("v", _fp), ("out", _fp), ("scores", _fp),
        ("row_tap", _fp), ("col_tap", _fp),
        ("cos_y", _fp), ("sin_y", _fp), ("cos_x", _fp), ("sin_x", _fp),
        ("B", C.c_int32), ("D", C.c_int32), ("C", C.c_int32), ("heads", C.c_int32),
        ("Ho", C.c_int32), ("Wo", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("K", C.c_int32),
        ("scale", C.c_float),
        ("q_stride_b", C.c_int64), ("q_stride_y", C.c_int64), ("q_stride_x", C.c_int64),
        ("algo", C.c_int32), ("rep_y", C.c_int32), ("rep_x", C.c_int32), ("out_dtype", C.c_int32),
        ("workspace", _fp), ("workspace_bytes", C.c_int64),
        ("q_dtype", C.c_int32), ("k_dtype", C.c_int32), ("v_dtype", C.c_int32), ("Kw", C.c_int32),
    ]

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.rep_y = self.rep_x = 1


class XAttnBwdParams(C.Structure):
    """Mirror of `naf_xattn_bwd_params` (include/naf_b200.h)."""

    _fields_ = [
        ("q", _fp), ("k", _fp), ("v", _fp), ("dout", _fp), ("dq", _fp), ("dk", _fp), ("dv", _fp),
        ("row_tap", _fp), ("col_tap", _fp),
        ("cos_y", _fp), ("sin_y", _fp), ("cos_x", _fp), ("sin_x", _fp),
        ("B", C.c_int32), ("D", C.c_int32), ("C", C.c_int32), ("heads", C.c_int32),
        ("Ho", C.c_int32), ("Wo", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("K", C.c_int32),
        ("scale", C.c_float),
        ("q_stride_b", C.c_int64), ("q_stride_y", C.c_int64), ("q_stride_x", C.c_int64),
        ("algo", C.c_int32), ("rep_y", C.c_int32), ("rep_x", C.c_int32),
        ("Kw", C.c_int32), ("reserved_", C.c_int32),
    ]

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.rep_y = self.rep_x = 1


class KPoolBwdParams(C.Structure):
    """Mirror of `naf_kpool_bwd_params` (include/naf_b200.h)."""

    _fields_ = [
        ("dq", _fp), ("dk", _fp), ("dx", _fp),
        ("cos_y", _fp), ("sin_y", _fp), ("cos_x", _fp), ("sin_x", _fp),
        ("B", C.c_int32), ("D", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("h", C.c_int32), ("w", C.c_int32), ("rope_heads", C.c_int32),
    ]


DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2

#: every symbol include/naf_b200.h declares: name -> (restype, argtypes)
EXPORTS = {
    "naf_abi_version": (C.c_int, []),
    "naf_last_error": (C.c_char_p, []),
    "naf_has_tensor_path": (C.c_int, []),
    "naf_launch_count": (C.c_ulonglong, [C.c_char_p]),
    "naf_pack_nhwc_f32": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_int64, C.c_int64, C.c_int64, C.c_int64, _fp]),
    "naf_pack_nhwc_slab_f32": (C.c_int, [_fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, _fp]),
    "naf_rope_kpool_f32": (C.c_int, [C.POINTER(KPoolParams), _fp]),
    "naf_xattn_fwd_f32": (C.c_int, [C.POINTER(XAttnParams), _fp]),
    "naf_xattn_select_algo": (C.c_int, [C.POINTER(XAttnParams)]),
    "naf_xattn_workspace_bytes": (C.c_size_t, [C.POINTER(XAttnParams)]),
    "naf_xattn_bwd_f32": (C.c_int, [C.POINTER(XAttnBwdParams), _fp]),
    "naf_rope_kpool_bwd_f32": (C.c_int, [C.POINTER(KPoolBwdParams), _fp]),
    "naf_concat_bias_nhwc_f32": (C.c_int, [_fp, _fp, C.c_int, _fp, _fp, C.c_int, _fp, C.c_int64, _fp]),
    "naf_gn_stats_f32": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int64, C.c_int, C.c_int, _fp]),
    "naf_gn_silu_apply_f32": (C.c_int, [_fp, _fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_float, C.c_int, _fp]),
    "naf_enc_stem_f32": (C.c_int, [_fp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _fp, _fp, _fp, _fp,
                                   C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "naf_enc_stem_tc_f32": (C.c_int, [_fp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _fp, _fp, _fp, _fp,
                                      C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "naf_enc_gn_coef_f32": (C.c_int, [_fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, _fp]),
    "naf_enc_conv_pack_f32": (C.c_int, [_fp, _fp, C.c_int, _fp]),
    "naf_enc_conv_f32": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_int, _fp, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_int, _fp]),
    "naf_enc_stem_ex": (C.c_int, [_fp, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _fp, _fp, _fp, _fp,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "naf_enc_conv_ex": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int64, C.c_int, _fp, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "naf_xattn_dump_taps_i32": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, _fp]),
}


class NafLibraryError(RuntimeError):
    """The native library is missing or returned a CUDA error."""


_lock = threading.Lock()
_lib = None


def load() -> C.CDLL:
    """Load libnaf_b200.so (once).  Raises NafLibraryError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise NafLibraryError(
                f"{LIB_PATH} not found: build it with `python -m naf_b200.csrc.build` "
                "(naf_b200 has no PyTorch/CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        ver = lib.naf_abi_version()
        if ver != ABI_VERSION:
            raise NafLibraryError(f"libnaf_b200.so has ABI {ver}, Python layer expects {ABI_VERSION}")
        _lib = lib
        return lib


def last_error() -> str:
    return load().naf_last_error().decode("utf-8", "replace")


def check(status: int, what: str) -> None:
    """Translate a naf_status into the exception the reference would have raised."""
    if status == NAF_OK:
        return
    msg = f"{what}: {last_error()}"
    if status == NAF_ERR_BAD_SHAPE and "divisible by num_heads" in msg:
        raise AssertionError(msg)  # reference: assert dim % num_heads == 0 (attentions.py:41)
    if status in (NAF_ERR_BAD_SHAPE, NAF_ERR_WINDOW, NAF_ERR_ALIGNMENT, NAF_ERR_NULL):
        raise ValueError(msg)
    if status == NAF_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise NafLibraryError(msg)
