"""one L0 self-attention launch (d=40, T=1400) for ncu"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiff_b200 import ops
n = int(os.environ.get("N_IMG", "24"))
L, d, dp = 1400, 40, 48
qkv = (torch.randn(n * L, 16 * dp + 8 * d, device="cuda") * 0.5).to(torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, qkv, qkv, n_img=n, lq=L, lk=L, heads=8, head_dim=d, q_col0=0, k_col0=8 * dp, v_col0=16 * dp)
torch.cuda.synchronize()
ts = []
for i in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.attention(qkv, qkv, qkv, n_img=n, lq=L, lk=L, heads=8, head_dim=d, q_col0=0, k_col0=8 * dp, v_col0=16 * dp)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
t = sorted(ts)[len(ts) // 2]
print(f"attention d={d} L={L} n_img={n} poly={os.environ.get('DD_ATTN_POLY', '0')}: {t * 1e3:8.1f} us  {4.0 * n * L * L * 8 * d / t / 1e9:7.1f} TFLOP/s")
