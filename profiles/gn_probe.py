"""GroupNorm+SiLU launches of the step (96 images, padded output, L2 flushed) against one build of the library:
    python profiles/gn_probe.py <lib.so>"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dualdiff_b200._lib as L
L.LIB_PATH = os.path.abspath(sys.argv[1])
from dualdiff_b200 import ops  # noqa: E402
tag = os.path.basename(sys.argv[1]); n = 96
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
line = f"{tag:14s}"
for (C1, C2, H, W) in [(320, 0, 28, 50), (640, 0, 14, 25), (1280, 0, 7, 13), (1280, 0, 4, 7), (1280, 1280, 4, 7), (1280, 1280, 7, 13), (640, 320, 28, 50)]:
    x1 = torch.randn(n * H * W, C1, device="cuda").to(torch.bfloat16)
    x2 = torch.randn(n * H * W, C2, device="cuda").to(torch.bfloat16) if C2 else None
    C = C1 + C2
    g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    out = torch.empty((n * (H + 1) * (W + 1), C), device="cuda", dtype=torch.bfloat16)
    ts = []
    for i in range(13):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.groupnorm(x1, g, b, n_img=n, hw=(H, W), x2=x2, silu=True, padded_out=True, out=out); e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    line += f"  C{C}_HW{H * W}: {sorted(ts)[len(ts) // 2] * 1e3:6.1f}"
print(line + "  (us)", flush=True)
