"""LayerNorm on the token rows of the three transformer levels (96 images): time and bandwidth per launch.
    python profiles/ln_one.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiff_b200 import ops
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for rows, C in ((96 * 1400, 320), (96 * 350, 640), (96 * 91, 1280)):
    x = torch.randn(rows, C, device="cuda").to(torch.bfloat16)
    g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    out = torch.empty_like(x)
    ts = []
    for i in range(11):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.layernorm(x, g, b, out=out); e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"LayerNorm rows={rows} C={C}: {t * 1e3:7.1f} us  {rows * C * 4 / t / 1e6:7.0f} GB/s (1 read + 1 write, cold L2)", flush=True)
