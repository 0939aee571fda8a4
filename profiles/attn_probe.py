"""short-K/V-stream attention launches (text Lk = 106 at every level, self / cross-view at levels 1-3) against one build of
the library, L2 flushed before every timed launch:  python profiles/attn_probe.py profiles/ab/lib_<tag>.so"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dualdiff_b200._lib as L
L.LIB_PATH = os.path.abspath(sys.argv[1])
from dualdiff_b200 import ops  # noqa: E402
tag = os.path.basename(sys.argv[1])
n = 96
kv_map = torch.tensor([[(i // 6) * 6 + (i + 5) % 6, (i // 6) * 6 + (i + 1) % 6] for i in range(n)], dtype=torch.int32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=9):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for d, Lq in ((40, 1400), (80, 350), (160, 91), (160, 28)):
    C = 8 * d
    dp = 48 if d == 40 else d                      # head_dim 40 is stored padded to 48 columns per head
    Cp = 8 * dp
    qkv = (torch.randn(n * Lq, 3 * Cp, device="cuda") * 0.5).to(torch.bfloat16)
    txt = (torch.randn(n * 106, 2 * Cp, device="cuda") * 0.5).to(torch.bfloat16)
    if d == 40:                                    # ones column of the padded V heads (softmax denominator)
        qkv.view(n * Lq, 3, 8, dp)[:, 2, :, 40] = 1.0
        txt.view(n * 106, 2, 8, dp)[:, 1, :, 40] = 1.0
    kw = dict(n_img=n, heads=8, head_dim=d)
    if d == 40:
        kw.update(v_ones=True)
    kinds = [("text", lambda: ops.attention(qkv, txt, txt, lq=Lq, lk=106, q_col0=0, k_col0=0, v_col0=Cp, q_cols=Cp, **kw), 106, 1)]
    if d != 40:
        kinds += [("self", lambda: ops.attention(qkv, qkv, qkv, lq=Lq, lk=Lq, q_col0=0, k_col0=Cp, v_col0=2 * Cp, **kw), Lq, 1),
                  ("xview", lambda: ops.attention(qkv, qkv, qkv, lq=Lq, lk=Lq, q_col0=0, k_col0=Cp, v_col0=2 * Cp, kv_map=kv_map, n_src=2, **kw), Lq, 2)]
    for kind, fn, lk, ns in kinds:
        t = timed(fn)
        print(f"{tag:18s} attention d={d:3d} {kind:5s} Lq={Lq:4d} Lk={lk:4d}: {t * 1e3:7.1f} us  {4.0 * n * Lq * lk * ns * C / t / 1e9:7.1f} TFLOP/s", flush=True)
