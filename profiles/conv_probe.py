"""3x3 implicit-GEMM conv launches of the step (96 images) against one build of the library, rotating operands:
    python profiles/conv_probe.py profiles/ab/lib_<tag>.so"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dualdiff_b200._lib as L
L.LIB_PATH = os.path.abspath(sys.argv[1])
from dualdiff_b200 import ops  # noqa: E402
tag = os.path.basename(sys.argv[1])
n = 96
# (label, H, W, c_in, c_out, residual)
CASES = [("L0 320->320 +tvec", 28, 50, 320, 320, False), ("L0 320->320 +tvec+res", 28, 50, 320, 320, True), ("L0 640->320", 28, 50, 640, 320, False),
         ("L0 960->320", 28, 50, 960, 320, False), ("L1 640->640", 14, 25, 640, 640, True), ("L1 1280->640", 14, 25, 1280, 640, False),
         ("L1 320->640", 14, 25, 320, 640, False), ("L2 1280->1280", 7, 13, 1280, 1280, True), ("L3 1280->1280", 4, 7, 1280, 1280, True)]
SETS = 4
mk = lambda *s, sc=0.5: (torch.randn(*s, device="cuda") * sc).to(torch.bfloat16)
for label, H, W, ci, co, res in CASES:
    A = [mk(ops.padded_rows(n, H, W), ci) for _ in range(SETS)]
    Wt = mk(co, 9 * ci, sc=0.02)
    R = [mk(n * H * W, co) for _ in range(SETS)] if res else [None] * SETS
    bias = torch.randn(co, device="cuda"); rv = torch.randn(n, co, device="cuda")
    run = lambda i: ops.gemm(A[i % SETS], Wt, bias=bias, rowvec=rv, rows_per_img=H * W, res1=R[i % SETS], taps=9, conv_hw=(H, W), n_img=n)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 16
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"{tag:18s} conv {label:22s} {us:8.1f} us  {2.0 * n * H * W * co * 9 * ci / us / 1e6:7.0f} TFLOP/s", flush=True)
    del A, R
