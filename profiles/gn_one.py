"""GroupNorm+SiLU launches for ncu: GN_C channels, GN_H x GN_W pixels, 96 images, padded output"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiff_b200 import ops
C, H, W = int(os.environ.get("GN_C", "320")), int(os.environ.get("GN_H", "28")), int(os.environ.get("GN_W", "50"))
n = 96
x = (torch.randn(n * H * W, C, device="cuda")).to(torch.bfloat16)
g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(13):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.groupnorm(x, g, b, n_img=n, hw=(H, W), silu=True, padded_out=True)
    e1.record()
    torch.cuda.synchronize()
    if i >= 3:
        ts.append(e0.elapsed_time(e1))
t = sorted(ts)[len(ts) // 2]
byt = n * H * W * C * 2 + n * (H + 1) * (W + 1) * C * 2
print(f"GN C={C} {H}x{W} fused={os.environ.get('DD_GN_FUSED', '1')}: {t * 1e3:7.1f} us  {byt / t / 1e6:7.0f} GB/s (1 read + 1 write)")
