"""ctypes binding of ``libseedvc_b200.so`` (C ABI in ``include/seedvc_b200.h``).

There is no fallback: if the library is missing or a call fails, an exception is
raised.  The reference's precedent for this boundary is its pybind op
``anti_alias_activation_cuda.forward`` (reference:
modules/bigvgan/alias_free_activation/cuda/anti_alias_activation.cpp:19-22).
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libseedvc_b200.so")

SVC_BF16, SVC_F32, SVC_F16 = 0, 1, 2
BACKEND_AUTO, BACKEND_SIMT = 0, 1
ACT_NONE, ACT_SILU, ACT_SWIGLU_PAIR, ACT_TANH_SIG_PAIR, ACT_ROPE = 0, 1, 2, 3, 4
MAX_SEG = 16

c_ll, c_int, c_float, c_vp = C.c_longlong, C.c_int, C.c_float, C.c_void_p


class GemmDesc(C.Structure):
    _fields_ = [
        ("dtype", c_int), ("B", c_int), ("T", c_int), ("N", c_int), ("n_seg", c_int),
        ("a_ptr", c_vp * MAX_SEG), ("a_bstride", c_ll * MAX_SEG), ("a_rstride", c_ll * MAX_SEG),
        ("a_rows", c_int * MAX_SEG), ("a_shift", c_int * MAX_SEG),
        ("w_ptr", c_vp * MAX_SEG), ("w_rstride", c_ll * MAX_SEG), ("K", c_int * MAX_SEG),
        ("bias", c_vp), ("rowbias", c_vp), ("rowbias_bstride", c_ll),
        ("act", c_int), ("rope_tab", c_vp), ("rope_cols", c_int), ("rope_pos0", c_int),
        ("q_cols", c_int), ("q_scale", c_float),
        ("gate", c_vp), ("gate_bstride", c_ll),
        ("res", c_vp), ("res_bstride", c_ll), ("res_rstride", c_ll),
        ("alpha", c_float), ("accumulate", c_int),
        ("out_f32", c_vp), ("of_bstride", c_ll), ("of_rstride", c_ll),
        ("out_op", c_vp), ("oo_bstride", c_ll), ("oo_rstride", c_ll),
        ("rope_tab_t", c_vp), ("rope_ld", c_int), ("out_op_dtype_p1", c_int),
    ]


# name -> argtypes; every symbol include/seedvc_b200.h declares
SIGNATURES = {
    "svc_gemm": [C.POINTER(GemmDesc), c_int, c_vp],
    "svc_attention": [c_vp, c_vp, c_vp, c_ll, c_ll, c_vp, c_ll, c_ll, c_int, c_int, c_int, c_vp,
                      c_int, c_int, c_vp],
    "svc_norm_mod": [c_vp, c_ll, c_ll, c_vp, c_vp, c_vp, c_float, c_int, c_vp, c_ll, c_ll, c_int,
                     c_int, c_int, c_int, c_vp],
    "svc_norm_mod_copy": [c_vp, c_ll, c_ll, c_vp, c_vp, c_vp, c_float, c_int, c_vp, c_vp, c_ll, c_ll, c_int,
                          c_int, c_int, c_int, c_vp],
    "svc_snake_aa": [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    "svc_snake_conv_post": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int,
                            c_int, c_vp],
    "svc_cfg_euler": [c_vp, c_vp, c_int, c_float, c_float, c_float, c_float, c_int, c_int, c_int,
                      c_int, c_vp, c_vp, c_int, c_vp],
    "svc_bct_to_btc": [c_vp, c_vp, c_ll, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_vp],
    "svc_btc_to_bct": [c_vp, c_vp, c_int, c_int, c_int, c_vp],
    "svc_cast": [c_vp, c_vp, c_ll, c_int, c_vp],
    "svc_reflect_halo": [c_vp, c_ll, c_ll, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp],
    "svc_timestep_embedding": [c_vp, c_vp, c_vp, c_int, c_int, c_vp],
    "svc_set_rows": [c_vp, c_ll, c_vp, c_ll, c_int, c_int, c_vp],
    "svc_interp_rows": [c_vp, c_ll, c_ll, c_vp, c_vp, c_vp, c_vp, c_ll, c_vp, c_vp, c_ll, c_ll, c_int, c_int,
                        c_int, c_int, c_vp],
    "svc_groupnorm1_mish": [c_vp, c_ll, c_ll, c_vp, c_vp, c_float, c_vp, c_vp, c_ll, c_ll, c_int, c_int, c_int,
                            c_int, c_int, c_vp],
    "svc_mask_rows": [c_vp, c_ll, c_ll, c_vp, c_int, c_int, c_int, c_vp],
    "svc_reflect_pad1d": [c_vp, c_ll, c_int, c_int, c_int, c_vp, c_ll, c_ll, c_vp],
    "svc_stft_mag": [c_vp, c_ll, c_ll, c_int, c_float, c_vp, c_ll, c_vp],
    "svc_log_clamp": [c_vp, c_ll, c_float, c_vp],
    "svc_sola_stitch": [c_vp, c_ll, c_int, c_vp, c_ll, c_vp, c_vp, c_vp, c_ll, c_vp, c_int, c_int, c_int, c_int,
                        c_vp],
    "svc_crossfade_stitch": [c_vp, c_ll, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_ll, c_vp],
}

_lib = None


class SvcError(RuntimeError):
    pass


def load_library(path: str | None = None):
    """Load the shared library and bind every entry point (raises if absent)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("SEEDVC_B200_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise SvcError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). seedvc_b200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(p)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = c_int
    lib.svc_last_error.restype = C.c_char_p
    lib.svc_last_error.argtypes = []
    lib.svc_version.restype = c_int
    lib.svc_version.argtypes = []
    if path is None:
        _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load_library().svc_last_error().decode(errors="replace")
        raise SvcError(f"{what} failed (code {rc}): {msg}")
