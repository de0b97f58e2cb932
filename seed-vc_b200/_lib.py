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
        ("row_ss_out", c_vp), ("row_ss_in", c_vp), ("rs_inv_dim", c_float), ("rs_eps", c_float),
    ]


SS_SLOTS = 8     # SVC_SS_SLOTS


MAX_LAYERS, MAX_WN_LAYERS, MAX_BRANCH = 32, 16, 3


class DitWeights(C.Structure):
    """svc_dit_weights (include/seedvc_b200.h)."""
    _fields_ = [
        ("version", c_int), ("D", c_int), ("H", c_int), ("L", c_int), ("C", c_int), ("I", c_int),
        ("time_as_token", c_int), ("style_as_token", c_int), ("uvit", c_int), ("long_skip", c_int),
        ("head", c_int), ("Dw", c_int), ("wn_layers", c_int), ("wn_kernel", c_int),
        ("op_dtype", c_int), ("stream_dtype", c_int),
        ("wqkv", c_vp * MAX_LAYERS), ("wo", c_vp * MAX_LAYERS), ("w13", c_vp * MAX_LAYERS), ("w2", c_vp * MAX_LAYERS),
        ("g_attn", c_vp * MAX_LAYERS), ("g_ffn", c_vp * MAX_LAYERS),
        ("skip_w", c_vp * MAX_LAYERS), ("skip_b", c_vp * MAX_LAYERS),
        ("ada_attn", c_int * MAX_LAYERS), ("ada_ffn", c_int * MAX_LAYERS),
        ("g_final", c_vp), ("ada_final", c_int),
        ("merge_wx", c_vp), ("merge_w_rstride", c_ll),
        ("lskip_w", c_vp), ("lskip_b", c_vp),
        ("mlp0_w", c_vp), ("mlp0_b", c_vp), ("mlp2_w", c_vp), ("mlp2_b", c_vp),
        ("conv1_w", c_vp), ("conv1_b", c_vp), ("resp_w", c_vp), ("conv2_w", c_vp), ("conv2_b", c_vp),
        ("fl_w", c_vp), ("fl_b", c_vp),
        ("wn_in_w", c_vp * MAX_WN_LAYERS), ("wn_rs_w", c_vp * MAX_WN_LAYERS), ("wn_rs_b", c_vp * MAX_WN_LAYERS),
        ("wn_skip_w", c_vp), ("wn_skip_b", c_vp), ("ada_fl", c_int),
        ("rope_tab", c_vp), ("rope_tab_t", c_vp), ("rope_ld", c_int),
    ]


class DitState(C.Structure):
    """svc_dit_state (include/seedvc_b200.h)."""
    _fields_ = [
        ("B", c_int), ("T", c_int), ("n_branch", c_int), ("n_steps", c_int), ("n_ada", c_int),
        ("ada", c_vp), ("t1", c_vp), ("wn_g", c_vp),
        ("const_kind", c_int * MAX_BRANCH), ("const_ptr", c_vp * MAX_BRANCH),
        ("style_tok", c_vp), ("style_tok_null", c_vp), ("branch_style", c_int * MAX_BRANCH),
        ("kv_len", c_vp), ("wn_lens", c_vp),
        ("h", c_vp), ("xn", c_vp), ("xn_f", c_vp), ("qkv", c_vp), ("att", c_vp), ("ff", c_vp), ("h_op", c_vp),
        ("skips", c_vp * (MAX_LAYERS // 2)), ("v", c_vp), ("x_res", c_vp), ("y", c_vp),
        ("xw", c_vp), ("xw_op", c_vp), ("acts", c_vp), ("wn_out", c_vp), ("ln", c_vp),
    ]


MAX_STAGES, MAX_TAPS = 8, 16


class ConvPlan(C.Structure):
    _fields_ = [("f", c_int), ("n_taps", c_int), ("shifts", c_int * MAX_TAPS), ("w", c_vp), ("b", c_vp), ("k", c_int)]


class AmpPair(C.Structure):
    _fields_ = [("a1", c_vp), ("inv_b1", c_vp), ("a2", c_vp), ("inv_b2", c_vp), ("c1", ConvPlan), ("c2", ConvPlan)]


class BigVGANStage(C.Structure):
    _fields_ = [("u", c_int), ("O", c_int), ("n_delta", c_int), ("deltas", c_int * MAX_TAPS), ("up_w", c_vp),
                ("up_b", c_vp), ("pairs", (AmpPair * 3) * 3)]


class BigVGANWeights(C.Structure):
    """svc_bigvgan_weights (include/seedvc_b200.h)."""
    _fields_ = [("n_mels", c_int), ("c0", c_int), ("n_stages", c_int), ("n_kernels", c_int), ("n_dil", c_int),
                ("op_dtype", c_int), ("precise", c_int), ("pre_w", c_vp), ("pre_b", c_vp),
                ("stages", BigVGANStage * MAX_STAGES), ("post_a", c_vp), ("post_inv_b", c_vp), ("post_w", c_vp),
                ("post_b", c_vp), ("post_k", c_int), ("use_tanh", c_int)]


# name -> argtypes; every symbol include/seedvc_b200.h declares
SIGNATURES = {
    "svc_gemm": [C.POINTER(GemmDesc), c_int, c_vp],
    "svc_attention": [c_vp, c_vp, c_vp, c_ll, c_ll, c_vp, c_ll, c_ll, c_int, c_int, c_int, c_vp,
                      c_int, c_int, c_vp],
    "svc_norm_mod": [c_vp, c_ll, c_ll, c_vp, c_vp, c_vp, c_float, c_int, c_vp, c_ll, c_ll, c_int,
                     c_int, c_int, c_int, c_vp],
    "svc_norm_mod_copy": [c_vp, c_ll, c_ll, c_vp, c_vp, c_vp, c_float, c_int, c_vp, c_vp, c_ll, c_ll, c_int,
                          c_int, c_int, c_int, c_int, c_vp],
    "svc_snake_aa": [c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    "svc_conv_post": [c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp],
    "svc_snake_conv_post": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int,
                            c_int, c_vp],
    "svc_cfg_euler": [c_vp, c_vp, c_int, c_float, c_float, c_float, c_float, c_int, c_int, c_int,
                      c_int, c_vp, c_vp, c_int, c_vp],
    "svc_bct_to_btc": [c_vp, c_vp, c_ll, c_ll, c_int, c_int, c_int, c_int, c_int, c_int, c_vp],
    "svc_btc_to_bct": [c_vp, c_vp, c_int, c_int, c_int, c_vp],
    "svc_cast": [c_vp, c_vp, c_ll, c_int, c_vp],
    "svc_scale_cols": [c_vp, c_ll, c_vp, c_vp, c_ll, c_vp, c_int, c_int, c_int, c_int, c_vp],
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
    "svc_unary": [c_vp, c_ll, c_ll, c_vp, c_ll, c_ll, c_int, c_int, c_int, c_int, c_float, c_vp, c_int, c_int, c_vp],
    "svc_hift_source": [c_vp, c_ll, c_vp, c_vp, c_vp, c_float, c_vp, c_vp, c_ll, c_int, c_int, c_int, c_int, c_float,
                        c_float, c_float, c_float, c_vp],
    "svc_hift_stft": [c_vp, c_ll, c_vp, c_ll, c_ll, c_int, c_int, c_int, c_int, c_int, c_vp],
    "svc_hift_istft": [c_vp, c_ll, c_ll, c_vp, c_ll, c_int, c_int, c_float, c_float, c_int, c_vp],
    "svc_bigvgan_workspace_bytes": [C.POINTER(BigVGANWeights), c_int, c_int],
    "svc_bigvgan_forward": [C.POINTER(BigVGANWeights), c_vp, c_vp, c_vp, c_int, c_int, c_vp],
    "svc_dit_step": [C.POINTER(DitWeights), C.POINTER(DitState), c_int, c_vp, c_vp],
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
        fn.restype = c_ll if name == "svc_bigvgan_workspace_bytes" else c_int
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
