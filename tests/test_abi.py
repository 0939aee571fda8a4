"""The C-ABI library loads on a CPU-only box and exports every symbol include/dualdiff_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dualdiff_b200.h")).read()
    return sorted(set(re.findall(r"DD_API\s+[\w\s\*]+?\b(dd_\w+)\s*\(", src)))


def test_header_symbols_exported():
    from dualdiff_b200 import _lib
    names = _declared()
    assert "dd_gemm" in names and "dd_version" in names
    so = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(so, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.EXPORTS) == names
    so.dd_version.restype = ctypes.c_int
    assert so.dd_version() == 100


def test_no_cpu_fallback():
    """ops refuse CPU tensors instead of silently computing on the host"""
    import pytest
    import torch
    from dualdiff_b200 import ops, _lib
    a = torch.zeros(128, 64, dtype=torch.bfloat16)
    with pytest.raises(_lib.DDError):
        ops.gemm(a, a)
