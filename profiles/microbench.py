"""Kernel microbenchmarks on the shapes of the DualDiff step (B=8 scenes -> n=96 images).  CUDA-event timing,
3 warm-ups, L2-sized inputs.  Usage: python profiles/microbench.py [gemm|attn|norm|all] [--ncu]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dualdiff_b200 import ops, packing  # noqa: E402

N_IMG = int(os.environ.get("N_IMG", "96"))
REPS = int(os.environ.get("REPS", "5"))


def timeit(fn, reps=REPS, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def mk(*shape):
    return (torch.randn(*shape, device="cuda") * 0.5).to(torch.bfloat16)


def bench_gemm():
    n = N_IMG
    print(f"--- gemm_tcgen05 (n_img={n}) ---")
    shapes = [  # (label, rows(HW), N, K, taps, geglu)
        ("L0 conv3x3 320->320", (28, 50), 320, 320, 9, 0), ("L0 conv3x3 640->320", (28, 50), 320, 640, 9, 0),
        ("L0 conv3x3 960->320", (28, 50), 320, 960, 9, 0),
        ("L1 conv3x3 640->640", (14, 25), 640, 640, 9, 0), ("L1 conv3x3 1280->640", (14, 25), 640, 1280, 9, 0),
        ("L2 conv3x3 1280->1280", (7, 13), 1280, 1280, 9, 0), ("L2 conv3x3 2560->1280", (7, 13), 1280, 2560, 9, 0),
        ("L3 conv3x3 1280->1280", (4, 7), 1280, 1280, 9, 0),
        ("L0 qkv 320->1088", (28, 50), 1088, 320, 1, 0), ("L0 out 320->320", (28, 50), 320, 320, 1, 0),
        ("L0 geglu 320->2560", (28, 50), 2560, 320, 1, 1), ("L0 ff2 1280->320", (28, 50), 320, 1280, 1, 0),
        ("L1 qkv 640->1920", (14, 25), 1920, 640, 1, 0), ("L1 geglu 640->5120", (14, 25), 5120, 640, 1, 1),
        ("L1 ff2 2560->640", (14, 25), 640, 2560, 1, 0),
        ("L2 qkv 1280->3840", (7, 13), 3840, 1280, 1, 0), ("L2 geglu 1280->10240", (7, 13), 10240, 1280, 1, 1),
        ("L2 ff2 5120->1280", (7, 13), 1280, 5120, 1, 0),
    ]
    for label, (H, W), N, K, taps, geglu in shapes:
        rows = n * H * W
        w = mk(N, K * taps)
        bias = torch.zeros(N, device="cuda")
        if taps == 9:
            a = mk(ops.padded_rows(n, H, W), K)
            fn = lambda: ops.gemm(a, w, bias=bias, taps=9, conv_hw=(H, W), n_img=n)
        else:
            a = mk(rows, K)
            fn = lambda: ops.gemm(a, w, bias=bias, geglu=bool(geglu))
        ms = timeit(fn)
        fl = 2.0 * rows * N * K * taps
        print(f"{label:28s} M={rows:7d} N={N:5d} K={K * taps:5d}  {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s")


def bench_attn():
    n = N_IMG
    print(f"--- attn_tcgen05 (n_img={n}) ---")
    kv_map = torch.tensor([[(i // 6) * 6 + (i % 6 + 5) % 6, (i // 6) * 6 + (i % 6 + 1) % 6] for i in range(n)],
                          dtype=torch.int32, device="cuda")
    for label, d, L, lk, nsrc in [("L0 self", 40, 1400, 1400, 1), ("L0 xview", 40, 1400, 1400, 2), ("L0 text", 40, 1400, 110, 1),
                                  ("L1 self", 80, 350, 350, 1), ("L1 xview", 80, 350, 350, 2), ("L1 text", 80, 350, 110, 1),
                                  ("L2 self", 160, 91, 91, 1), ("L2 xview", 160, 91, 91, 2), ("L3 self", 160, 28, 28, 1)]:
        dp = 48 if d == 40 else d
        if lk == L:
            qkv = mk(n * L, 16 * dp + 8 * d)
            fn = lambda: ops.attention(qkv, qkv, qkv, n_img=n, lq=L, lk=L, heads=8, head_dim=d, q_col0=0, k_col0=8 * dp,
                                       v_col0=16 * dp, kv_map=kv_map if nsrc == 2 else None, n_src=nsrc)
        else:
            q = mk(n * L, 8 * dp); kv = mk(n * lk, 8 * dp + 8 * d)
            fn = lambda: ops.attention(q, kv, kv, n_img=n, lq=L, lk=lk, heads=8, head_dim=d, k_col0=0, v_col0=8 * dp)
        ms = timeit(fn)
        fl = 4.0 * n * L * lk * 8 * d * nsrc
        print(f"{label:10s} d={d:3d} Lq={L:5d} Lk={lk:5d} src={nsrc}  {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s")


def bench_norm():
    n = N_IMG
    print(f"--- groupnorm / layernorm (n_img={n}) ---")
    for (H, W), c1, c2 in [((28, 50), 320, 0), ((28, 50), 640, 320), ((14, 25), 640, 0), ((14, 25), 1280, 640),
                           ((7, 13), 1280, 0), ((7, 13), 1280, 1280)]:
        C = c1 + c2
        x1 = mk(n * H * W, c1); x2 = mk(n * H * W, c2) if c2 else None
        g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
        ms = timeit(lambda: ops.groupnorm(x1, g, b, n_img=n, hw=(H, W), x2=x2, padded_out=True))
        by = 2.0 * (2 * n * H * W * C + ops.padded_rows(n, H, W) * C)
        print(f"groupnorm+silu {H}x{W} C={C:5d}  {ms:8.3f} ms  {by / ms / 1e6:8.1f} GB/s (2 reads + 1 write)")
    for rows, C in [(n * 1400, 320), (n * 350, 640), (n * 91, 1280)]:
        x = mk(rows, C); g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
        ms = timeit(lambda: ops.layernorm(x, g, b))
        print(f"layernorm rows={rows} C={C:5d}  {ms:8.3f} ms  {4.0 * rows * C / ms / 1e6:8.1f} GB/s")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("gemm", "all"):
        bench_gemm()
    if what in ("attn", "all"):
        bench_attn()
    if what in ("norm", "all"):
        bench_norm()
