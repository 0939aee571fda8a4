"""tile-width sweep (force_bn testing hook) over the TMA-epilogue shapes of levels 1-2, rotating operands:
    python profiles/gemm_bn_sweep.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dualdiff_b200 import ops
CASES = [("N640_K640+r", 33600, 640, 640, True), ("N1920_K640", 33600, 1920, 640, False), ("N640_K2560+r", 33600, 640, 2560, True),
         ("N640_K320", 33600, 640, 320, False), ("N1280_K1280+r", 8736, 1280, 1280, True), ("N3840_K1280", 8736, 3840, 1280, False),
         ("N1280_K5120+r", 8736, 1280, 5120, True), ("N320_K1280+r", 134400, 320, 1280, True), ("N1280_K1280+r L3", 2688, 1280, 1280, True)]
SETS = 6
mk = lambda *s, sc=0.5: (torch.randn(*s, device="cuda") * sc).to(torch.bfloat16)
for label, M, N, K, res in CASES:
    A = [mk(M, K) for _ in range(SETS)]; W = mk(N, K, sc=0.05); bias = torch.randn(N, device="cuda")
    R = [mk(M, N) for _ in range(SETS)] if res else [None] * SETS
    outs = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(SETS)]
    line = f"{label:18s}"
    for bn in (0, 128, 192, 256):
        run = lambda i: ops.gemm(A[i % SETS], W, out=outs[i % SETS], bias=bias, res1=R[i % SETS], force_bn=bn)
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(24):
            run(i)
        e1.record(); torch.cuda.synchronize()
        line += f"  bn{bn if bn else 'auto':>4}: {e0.elapsed_time(e1) / 24 * 1e3:7.1f} us"
    print(line, flush=True)
