"""Differential timing of the plain-GEMM (TMA epilogue) shapes of the step against one build of the library:
    python profiles/gemm_probe.py profiles/ab/lib_<tag>.so [check]
Each shape is timed with rotating operands (8 sets -> every launch reads cold data, as inside the step) and, with `check`,
compared with a torch fp32 matmul on sampled rows (the DD_PROBE builds are wrong by construction: no check there)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dualdiff_b200._lib as L
L.LIB_PATH = os.path.abspath(sys.argv[1])
from dualdiff_b200 import ops  # noqa: E402
CHECK = len(sys.argv) > 2 and sys.argv[2] == "check"
tag = os.path.basename(sys.argv[1])

# (label, M, N, K, residual, geglu, force_bn)
CASES = [("N320_K320+r", 134400, 320, 320, True, False, 0), ("N320_K320", 134400, 320, 320, False, False, 0),
         ("N320_K320+r bn128", 134400, 320, 320, True, False, 128), ("N320_K320+r bn256", 134400, 320, 320, True, False, 256),
         ("N320_K40", 134400, 320, 40, False, False, 0), ("N384_K320", 134400, 384, 320, False, False, 0),
         ("N1152_K320", 134400, 1152, 320, False, False, 0), ("N2560_K320 geglu", 134400, 2560, 320, False, True, 0),
         ("N320_K1280+r", 134400, 320, 1280, True, False, 0),
         ("N640_K640+r", 33600, 640, 640, True, False, 0), ("N1920_K640", 33600, 1920, 640, False, False, 0),
         ("N5120_K640 geglu", 33600, 5120, 640, False, True, 0), ("N640_K2560+r", 33600, 640, 2560, True, False, 0),
         ("N1280_K1280+r", 8736, 1280, 1280, True, False, 0), ("N3840_K1280", 8736, 3840, 1280, False, False, 0)]
SETS = 6


def mk(*shape, s=0.5):
    return (torch.randn(*shape, device="cuda") * s).to(torch.bfloat16)


for label, M, N, K, res, geglu, bn in CASES:
    A = [mk(M, K) for _ in range(SETS)]
    W = mk(N, K, s=0.05)
    R = [mk(M, N) for _ in range(SETS)] if res else [None] * SETS
    bias = torch.randn(N, device="cuda")
    outs = [torch.empty(M, N // 2 if geglu else N, device="cuda", dtype=torch.bfloat16) for _ in range(SETS)]
    run = lambda i: ops.gemm(A[i % SETS], W, out=outs[i % SETS], bias=bias, res1=R[i % SETS], geglu=geglu, force_bn=bn)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 24
    e0.record()
    for i in range(reps):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    msg = ""
    if CHECK:
        idx = torch.cat([torch.arange(0, 300, device="cuda"), torch.randint(0, M, (2000,), device="cuda"),
                         torch.arange(M - 300, M, device="cuda")])
        run(0)
        ref = A[0][idx].float() @ W.float().t() + bias
        if geglu:
            t = ref.view(ref.shape[0], N // 256, 2, 128)     # per 256-column tile: [value 128 | gate 128] (packing.py)
            ref = (t[:, :, 0] * torch.nn.functional.gelu(t[:, :, 1])).reshape(ref.shape[0], N // 2)
        if res:
            ref = ref + R[0][idx].float()
        got = outs[0][idx].float()
        err = ((got - ref).norm() / ref.norm()).item()
        msg = f"  rel-L2 {err:.2e}" + ("" if err < 8e-3 else "  MISMATCH")
    byts = (M * K + N * K + M * (N // 2 if geglu else N) * (2 if res else 1)) * 2
    print(f"{tag:18s} {label:20s} {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.0f} TFLOP/s  {byts / us / 1e3:6.0f} GB/s{msg}", flush=True)
    del A, R, outs
