"""ctypes binding of libdualdiff_sm100.so (C ABI declared in include/dualdiff_b200.h).

There is no CPU fallback: if the shared library is missing, importing :mod:`dualdiff_b200.ops`
raises; if a call fails the wrapper raises ``RuntimeError`` with ``dd_last_error()``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdualdiff_sm100.so")


class DDError(RuntimeError):
    pass


_vp = C.c_void_p
_ll = C.c_longlong
_i = C.c_int
_f = C.c_float


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", _vp), ("a2", _vp), ("w", _vp), ("out", _vp), ("bias", _vp), ("rowvec", _vp),
        ("res1", _vp), ("res2", _vp),
        ("M", _i), ("N", _i), ("K", _i), ("k1", _i), ("taps", _i), ("conv_h", _i), ("conv_w", _i),
        ("a_ld", _ll), ("a2_ld", _ll), ("w_ld", _ll), ("out_ld", _ll), ("res1_ld", _ll), ("res2_ld", _ll),
        ("rowvec_ld", _i), ("rows_per_img", _i), ("out_f32", _i), ("geglu", _i), ("force_bn", _i), ("act", _i), ("no_tma_epilogue", _i), ("one_cta", _i),
        ("stream_k", _i), ("workspace", _vp), ("workspace_bytes", _ll),
    ]


class GroupNormArgs(C.Structure):
    _fields_ = [
        ("x1", _vp), ("x2", _vp), ("out", _vp), ("stats", _vp), ("gamma", _vp), ("beta", _vp),
        ("x1_ld", _ll), ("x2_ld", _ll), ("out_ld", _ll),
        ("n_img", _i), ("h", _i), ("w", _i), ("c1", _i), ("c2", _i), ("groups", _i),
        ("eps", _f), ("silu", _i), ("padded_out", _i), ("two_pass", _i),
    ]


class LayerNormArgs(C.Structure):
    _fields_ = [("x", _vp), ("out", _vp), ("gamma", _vp), ("beta", _vp), ("x_ld", _ll), ("out_ld", _ll),
                ("rows", _i), ("c", _i), ("eps", _f)]


class AttentionArgs(C.Structure):
    _fields_ = [
        ("q", _vp), ("k", _vp), ("v", _vp), ("out", _vp), ("kv_map", _vp),
        ("q_ld", _ll), ("k_ld", _ll), ("v_ld", _ll), ("out_ld", _ll),
        ("q_cols", _i), ("k_cols", _i), ("v_cols", _i),
        ("q_col0", _i), ("k_col0", _i), ("v_col0", _i),
        ("q_head_stride", _i), ("k_head_stride", _i), ("v_head_stride", _i),
        ("n_img", _i), ("n_kv_img", _i), ("heads", _i), ("head_dim", _i), ("lq", _i), ("lk", _i), ("n_src", _i),
        ("scale", _f), ("variant", _i), ("v_ones", _i), ("concat", _i),
    ]


class TemporalAttentionArgs(C.Structure):
    _fields_ = [
        ("q", _vp), ("k", _vp), ("v", _vp), ("out", _vp),
        ("q_ld", _ll), ("k_ld", _ll), ("v_ld", _ll), ("out_ld", _ll),
        ("q_col0", _i), ("k_col0", _i), ("v_col0", _i),
        ("q_head_stride", _i), ("k_head_stride", _i), ("v_head_stride", _i),
        ("n_outer", _i), ("n_view", _i), ("tokens", _i), ("heads", _i), ("head_dim", _i),
        ("frames_q", _i), ("frames_kv", _i), ("frames_per_rank", _i),
        ("kv_rank_stride", _ll), ("scale", _f),
        ("frames_q_per_rank", _i), ("q_rank_stride", _ll),
    ]


class SeqAttentionArgs(C.Structure):
    _fields_ = [
        ("q", _vp), ("k", _vp), ("v", _vp), ("out", _vp),
        ("q_ld", _ll), ("k_ld", _ll), ("v_ld", _ll), ("out_ld", _ll),
        ("q_col0", _i), ("k_col0", _i), ("v_col0", _i),
        ("q_head_stride", _i), ("k_head_stride", _i), ("v_head_stride", _i),
        ("n_seq", _i), ("seq_len", _i), ("heads", _i), ("head_dim", _i), ("causal", _i), ("scale", _f),
    ]


class ToPaddedArgs(C.Structure):
    _fields_ = [("src", _vp), ("out", _vp), ("stride_outer", _ll), ("stride_view", _ll), ("stride_c", _ll),
                ("stride_h", _ll), ("n_outer", _i), ("n_view", _i), ("c", _i), ("h", _i), ("w", _i), ("cp", _i),
                ("src_f32", _i)]


class LinearF32Args(C.Structure):
    _fields_ = [("x", _vp), ("w", _vp), ("b", _vp), ("y", _vp), ("y16", _vp), ("x_ld", _ll), ("y_ld", _ll),
                ("y16_ld", _ll), ("M", _i), ("N", _i), ("K", _i), ("act", _i)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DDError(
                f"{LIB_PATH} not found - build it with dualdiff_b200/csrc/build.sh "
                "(or __graft_entry__.build()); there is no CPU fallback")
        _lib = C.CDLL(LIB_PATH)
        _lib.dd_version.restype = _i
        _lib.dd_last_error.restype = C.c_char_p
        _lib.dd_launch_count.restype = _ll
        _lib.dd_groupnorm_scratch_floats.restype = _ll
        for name in EXPORTS:
            if name in ("dd_version", "dd_last_error", "dd_launch_count", "dd_groupnorm_scratch_floats"):
                continue
            fn = getattr(_lib, name)
            fn.restype = _i
    return _lib


# every symbol include/dualdiff_b200.h declares (tests/test_abi.py checks the list against the header)
EXPORTS = [
    "dd_version", "dd_last_error", "dd_launch_count", "dd_gemm", "dd_groupnorm", "dd_groupnorm_scratch_floats", "dd_layernorm", "dd_attention", "dd_temporal_attention", "dd_ors_project",
    "dd_nchw_to_padded", "dd_im2col_s2", "dd_upsample_pad", "dd_pad_rows", "dd_linear_f32",
    "dd_timestep_embedding", "dd_fourier_embed", "dd_box_features", "dd_silu_to_bf16", "dd_add_bf16",
    "dd_nchw_to_rows", "dd_rows_to_nchw", "dd_cfg_sched_step", "dd_softmax_rows",
    "dd_clip_embed", "dd_seq_attention", "dd_quick_gelu", "dd_nchw_patches",
]


def check(rc, what):
    if rc != 0:
        msg = lib().dd_last_error().decode("utf-8", "replace")
        raise DDError(f"{what} failed (rc={rc}): {msg}")
