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
        ("rowvec_ld", _i), ("rows_per_img", _i), ("out_f32", _i), ("geglu", _i), ("force_bn", _i),
    ]


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
        for name in EXPORTS:
            if name in ("dd_version", "dd_last_error", "dd_launch_count"):
                continue
            fn = getattr(_lib, name)
            fn.restype = _i
    return _lib


# every symbol include/dualdiff_b200.h declares (tests/test_abi.py checks the list against the header)
EXPORTS = [
    "dd_version", "dd_last_error", "dd_launch_count", "dd_gemm",
]


def check(rc, what):
    if rc != 0:
        msg = lib().dd_last_error().decode("utf-8", "replace")
        raise DDError(f"{what} failed (rc={rc}): {msg}")
