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
r = (torch.randn(n * H * W, C, device="cuda") * 0.5).to(torch.bfloat16)
for _ in range(3):
    ops.gemm(a, w, bias=b, taps=9, conv_hw=(H, W), n_img=n, res1=r)
torch.cuda.synchronize()
