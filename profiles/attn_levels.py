"""level 1-3 attention launches (head_dim 80 / 160: 350 / 91 / 28 tokens) timed with one CTA per item (variant 1) against
the persistent item loop (variant 0): self, cross-view (two sources) and text (Lk = 106).
    python profiles/attn_levels.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiff_b200 import ops
n = int(os.environ.get("N_IMG", "96"))
kv_map = torch.tensor([[(i // 6) * 6 + (i + 5) % 6, (i // 6) * 6 + (i + 1) % 6] for i in range(n)], dtype=torch.int32, device="cuda")


def make(d, L):
    C = 8 * d
    qkv = (torch.randn(n * L, 3 * C, device="cuda") * 0.5).to(torch.bfloat16)
    txt = (torch.randn(n * 106, 2 * C, device="cuda") * 0.5).to(torch.bfloat16)
    return qkv, txt


def run(kind, d, L, qkv, txt, variant):
    C = 8 * d
    if kind == "self":
        return ops.attention(qkv, qkv, qkv, n_img=n, lq=L, lk=L, heads=8, head_dim=d, q_col0=0, k_col0=C, v_col0=2 * C, variant=variant)
    if kind == "xview":
        return ops.attention(qkv, qkv, qkv, n_img=n, lq=L, lk=L, heads=8, head_dim=d, q_col0=0, k_col0=C, v_col0=2 * C,
                             kv_map=kv_map, n_src=2, variant=variant)
    return ops.attention(qkv, txt, txt, n_img=n, lq=L, lk=106, heads=8, head_dim=d, q_col0=0, k_col0=0, v_col0=C, q_cols=C,
                         variant=variant)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for d, L in ((80, 350), (160, 91), (160, 28)):
    qkv, txt = make(d, L)
    for kind in ("self", "xview", "text"):
        lk, ns = (106, 1) if kind == "text" else (L, 2 if kind == "xview" else 1)
        ref = None
        for v in (1, 0):
            t = timed(lambda: run(kind, d, L, qkv, txt, v))
            out = run(kind, d, L, qkv, txt, v).float()
            ref = out if ref is None else ref
            err = ((out - ref).abs().max() / ref.abs().max()).item()
            print(f"attention d={d} {kind:5s} Lq={L} Lk={lk} n_img={n} {'one CTA per item' if v == 1 else 'persistent      '}: "
                  f"{t * 1e3:7.1f} us  {4.0 * n * L * lk * ns * 8 * d / t / 1e9:7.1f} TFLOP/s  max-rel diff {err:.2e}", flush=True)
