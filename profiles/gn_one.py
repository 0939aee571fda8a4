"""GroupNorm+SiLU on the shapes of the step (96 images, padded output): the single-pass cluster kernel against the two-kernel form.
    python profiles/gn_one.py             # table
    GN_C=320 GN_H=28 GN_W=50 TWO_PASS=0 python profiles/gn_one.py   # one configuration (for ncu)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiff_b200 import ops
n = 96
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(C1, C2, H, W, two_pass, cold=True):
    x1 = (torch.randn(n * H * W, C1, device="cuda")).to(torch.bfloat16)
    x2 = (torch.randn(n * H * W, C2, device="cuda")).to(torch.bfloat16) if C2 else None
    C = C1 + C2
    g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    out = torch.empty((n * (H + 1) * (W + 1), C), device="cuda", dtype=torch.bfloat16)
    ts = []
    for i in range(11):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.groupnorm(x1, g, b, n_img=n, hw=(H, W), x2=x2, silu=True, padded_out=True, out=out, two_pass=two_pass)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    byt = n * H * W * C * 2 + n * (H + 1) * (W + 1) * C * 2
    return t, byt


if os.environ.get("GN_C"):
    shapes = [(int(os.environ["GN_C"]), 0, int(os.environ.get("GN_H", "28")), int(os.environ.get("GN_W", "50")))]
    modes = [bool(int(os.environ.get("TWO_PASS", "0")))]
else:
    shapes = [(320, 0, 28, 50), (640, 0, 14, 25), (1280, 0, 7, 13), (1280, 0, 4, 7), (1280, 1280, 4, 7), (1280, 1280, 7, 13),
              (1280, 640, 7, 13), (1280, 640, 14, 25), (640, 640, 14, 25), (640, 320, 14, 25), (640, 320, 28, 50), (320, 320, 28, 50)]
    modes = [True, False]
for (C1, C2, H, W) in shapes:
    for tp in modes:
        t, byt = timed(C1, C2, H, W, tp)
        print(f"GN C={C1}+{C2} {H}x{W} {'two-pass ' if tp else 'one-pass '}: {t * 1e3:7.1f} us  {byt / t / 1e6:7.0f} GB/s (1 read + 1 write, cold L2)", flush=True)
