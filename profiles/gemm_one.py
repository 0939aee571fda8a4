"""one small-K GEMM launch with residual (M=134400, N=320, K=320) for ncu"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiff_b200 import ops
M, N, K = 134400, int(os.environ.get("GN", "320")), int(os.environ.get("GK", "320"))
a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
r = (torch.randn(M, N, device="cuda") * 0.5).to(torch.bfloat16)
b = torch.zeros(N, device="cuda")
use_res = os.environ.get('GRES', '1') == '1'
for _ in range(3):
    ops.gemm(a, w, bias=b, res1=r if use_res else None)
torch.cuda.synchronize()
ts = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(12):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gemm(a, w, bias=b, res1=r if use_res else None); e1.record()
    torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
t = sorted(ts)[6]
print(f"gemm M={M} N={N} K={K} res={use_res}: {t * 1e3:.1f} us  {2.0 * M * N * K / t / 1e9:.0f} TFLOP/s  {(M * K + M * N * (2 if use_res else 1)) * 2 / t / 1e6:.0f} GB/s algorithmic")
