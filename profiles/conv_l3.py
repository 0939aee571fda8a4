"""L3 (4x7-pixel) 3x3 conv, 96 images, 1280 -> 1280: classic vs stream-K (CUDA events, 20 reps)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiff_b200 import ops, packing
n, H, W, ci, co = 96, 4, 7, int(os.environ.get("CI", "1280")), 1280
x = (torch.randn(n, H, W, ci, device="cuda") * 0.5).to(torch.bfloat16)
w = (torch.randn(co, ci, 3, 3, device="cuda") * 0.01).to(torch.bfloat16)
a, wp = packing.to_padded(x), packing.pack_conv3x3(w)
b = torch.zeros(co, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for sk in (-1, 0, 1):
    ts = []
    for i in range(23):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(a, wp, bias=b, taps=9, conv_hw=(H, W), n_img=n, stream_k=sk)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"stream_k={sk:2d}: {t * 1e3:8.1f} us  {2.0 * n * H * W * co * ci * 9 / t / 1e9:8.1f} TFLOP/s")
