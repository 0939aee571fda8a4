"""one implicit-GEMM 3x3 conv launch for ncu: LEVEL=0 (N=320,K=2880, M=134400) or 2 (N=1280,K=11520, M=8736)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiff_b200 import ops
n = 96
lvl = int(os.environ.get("LEVEL", "2"))
H, W, C = {0: (28, 50, 320), 1: (14, 25, 640), 2: (7, 13, 1280)}[lvl]
a = (torch.randn(ops.padded_rows(n, H, W), C, device="cuda") * 0.5).to(torch.bfloat16)
w = (torch.randn(C, 9 * C, device="cuda") * 0.02).to(torch.bfloat16)
b = torch.zeros(C, device="cuda")
r = (torch.randn(n * H * W, C, device="cuda") * 0.5).to(torch.bfloat16) if os.environ.get("GRES", "1") == "1" else None
for _ in range(3):
    ops.gemm(a, w, bias=b, taps=9, conv_hw=(H, W), n_img=n, res1=r)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(12):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.gemm(a, w, bias=b, taps=9, conv_hw=(H, W), n_img=n, res1=r)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = sorted(ts)[len(ts) // 2]
print(f"conv level {lvl}: {t * 1e3:8.1f} us  {2.0 * n * H * W * C * C * 9 / t / 1e9:8.1f} TFLOP/s")
