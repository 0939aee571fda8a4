"""tile-width sweep for the GEMM shapes that dominate the step (force_bn testing hook)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiff_b200 import ops
from microbench import timeit, mk
n = 96
cases = [("N320_K320", 134400, 320, 320, 1, True), ("N1088_K320", 134400, 1088, 320, 1, False), ("N640_K640", 33600, 640, 640, 1, True),
         ("N1280_K1280", 8736, 1280, 1280, 1, True), ("conv N320_K2880", 134400, 320, 320, 9, True),
         ("conv N1280_K11520 L2", 8736, 1280, 1280, 9, True), ("conv N1280_K11520 L3", 2688, 1280, 1280, 9, True),
         ("N320_K1280", 134400, 320, 1280, 1, True)]
geo = {134400: (28, 50), 33600: (14, 25), 8736: (7, 13), 2688: (4, 7)}
for label, rows, N, K, taps, res in cases:
    H, W = geo[rows]
    w = mk(N, K * taps); bias = torch.zeros(N, device="cuda")
    r1 = mk(rows, N) if res else None
    a = mk(ops.padded_rows(n, H, W), K) if taps == 9 else mk(rows, K)
    line = f"{label:24s}"
    for bn in (128, 160, 192, 256):
        if taps == 9:
            fn = lambda: ops.gemm(a, w, bias=bias, taps=9, conv_hw=(H, W), n_img=n, res1=r1, force_bn=bn)
        else:
            fn = lambda: ops.gemm(a, w, bias=bias, res1=r1, force_bn=bn)
        ms = timeit(fn)
        line += f"  bn{bn}: {ms:6.3f} ms {2.0 * rows * N * K * taps / ms / 1e9:6.0f}"
    print(line)
