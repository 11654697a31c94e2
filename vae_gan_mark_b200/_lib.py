"""ctypes binding of libvaegan_b200.so (the C ABI declared in include/vaegan_b200.h).

There is no fallback: if the shared library is missing or fails to load, importing the ops raises.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libvaegan_b200.so")
VG_MAX_TAPS = 16
VG_MAX_FPROP_TAPS = 96


class VgError(RuntimeError):
    pass


class VgConvFprop(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_n", C.c_int), ("x_h", C.c_int), ("x_w", C.c_int), ("x_ld", C.c_int),
        ("x_stride", C.c_int), ("m_n", C.c_int), ("m_h", C.c_int), ("m_w", C.c_int),
        ("cin", C.c_int), ("num_taps", C.c_int), ("taps", (C.c_int * 4) * VG_MAX_FPROP_TAPS),
        ("use_wk", C.c_int), ("wk", C.c_int * VG_MAX_FPROP_TAPS),
        ("w", C.c_void_p), ("w_ld", C.c_int), ("n_gemm", C.c_int),
        ("out", C.c_void_p), ("out_kind", C.c_int),
        ("out_h", C.c_int), ("out_w", C.c_int), ("out_ld", C.c_int), ("out_coff", C.c_int),
        ("su_h", C.c_int), ("su_w", C.c_int), ("sub_h0", C.c_int), ("sub_w0", C.c_int), ("cout_per_sub", C.c_int),
        ("bias", C.c_void_p), ("act", C.c_int), ("ksplit", C.c_int), ("force_bn", C.c_int),
        ("b_mn_major", C.c_int), ("w_rows", C.c_int),
        ("num_groups", C.c_int), ("group_ntaps", C.c_int * 4), ("group_sub", (C.c_int * 2) * 4),
        ("halo_mode", C.c_int), ("stats", C.c_void_p),
    ]


class VgConvWgrad(C.Structure):
    _fields_ = [
        ("g", C.c_void_p), ("g_ld", C.c_int), ("g_coff", C.c_int), ("cout", C.c_int),
        ("x", C.c_void_p), ("x_n", C.c_int), ("x_h", C.c_int), ("x_w", C.c_int), ("x_ld", C.c_int),
        ("x_stride", C.c_int), ("m_n", C.c_int), ("m_h", C.c_int), ("m_w", C.c_int),
        ("cin", C.c_int), ("num_taps", C.c_int), ("taps", (C.c_int * 4) * VG_MAX_TAPS),
        ("num_combos", C.c_int), ("combo_g", C.c_int * 8), ("combo_x", C.c_int * 8),
        ("dw", C.c_void_p), ("dw_ld", C.c_int), ("ksplit", C.c_int), ("force_bn", C.c_int),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_longlong),
    ]


class VgNormApply(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_ld", C.c_int), ("x_coff", C.c_int),
        ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("c", C.c_int),
        ("mean_rstd", C.c_void_p), ("per_sample", C.c_int), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("act", C.c_int), ("y", C.c_void_p), ("y_ld", C.c_int), ("y_coff", C.c_int),
        ("pool", C.c_void_p), ("p_ld", C.c_int), ("p_coff", C.c_int), ("dtype", C.c_int),
    ]


class VgNormBackward(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("x_ld", C.c_int), ("x_coff", C.c_int),
        ("dy", C.c_void_p), ("dy_ld", C.c_int), ("dy_coff", C.c_int),
        ("dpool", C.c_void_p), ("dp_ld", C.c_int), ("dp_coff", C.c_int),
        ("n", C.c_int), ("h", C.c_int), ("w", C.c_int), ("c", C.c_int),
        ("mean_rstd", C.c_void_p), ("per_sample", C.c_int), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("act", C.c_int), ("sums", C.c_void_p),
        ("dx", C.c_void_p), ("dx_ld", C.c_int), ("dx_coff", C.c_int),
        ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("accumulate", C.c_int), ("dtype", C.c_int),
        ("virt_h", C.c_int),
    ]


class VgWarpJob(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("src_h", C.c_int), ("src_w", C.c_int), ("channels", C.c_int), ("src_row_bytes", C.c_longlong),
        ("minv", C.c_double * 9), ("out_h", C.c_int), ("out_w", C.c_int),
        ("dst_u8", C.c_void_p), ("dst_chw", C.c_void_p), ("transparent", C.c_int),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load the shared library once; raise loudly if it is not there (no CPU / eager fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise VgError(f"{LIB_PATH} is missing: build it with `python -m vae_gan_mark_b200.build` "
                          "(or __graft_entry__.build()); this package has no fallback path")
        _lib = C.CDLL(LIB_PATH)
        _lib.vg_last_error.restype = C.c_char_p
        _lib.vg_version.restype = C.c_int
        _lib.vg_launch_count.restype = C.c_ulonglong
        _lib.vg_conv_wgrad_workspace.restype = C.c_longlong
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise VgError(f"{what}: {lib().vg_last_error().decode(errors='replace')} (code {rc})")


def call(name: str, *args) -> None:
    """Call an int-returning C-ABI function and raise VgError on a non-zero return."""
    fn = getattr(lib(), name)
    check(fn(*args), name)
