"""level-0 attention launches (d = 40) timed per kernel variant: self (T x T), cross-view (two sources) and text (Lk = 106).
    python profiles/attn_one.py            # table over variants 1 (one-tile), 2 (two-tile, all MUFU), 0 (two-tile + FMA-pipe share)
    VARIANT=0 ONLY=self python profiles/attn_one.py   # one configuration (for ncu)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiff_b200 import ops
n = int(os.environ.get("N_IMG", "96"))
L = int(os.environ.get("TOKENS", "1400"))
d, dp = 40, 48
ones = os.environ.get("V_ONES", "0") == "1"      # the shipped layout: V heads padded to 48 columns, column 40 = 1.0
dv = dp if ones else d
qkv = (torch.randn(n * L, 16 * dp + 8 * dv, device="cuda") * 0.5).to(torch.bfloat16)
txt = (torch.randn(n * 106, 8 * dp + 8 * dv, device="cuda") * 0.5).to(torch.bfloat16)
if ones:
    for t, rows, k0 in ((qkv, n * L, 16), (txt, n * 106, 8)):
        v = t.view(rows, -1, dp)
        v[:, k0:, d:] = 0
        v[:, k0:, d] = 1.0
kv_map = torch.tensor([[(i // 6) * 6 + (i + 5) % 6, (i // 6) * 6 + (i + 1) % 6] for i in range(n)], dtype=torch.int32, device="cuda")


def run(kind, variant):
    if kind == "self":
        return ops.attention(qkv, qkv, qkv, n_img=n, lq=L, lk=L, heads=8, head_dim=d, q_col0=0, k_col0=8 * dp, v_col0=16 * dp,
                             variant=variant, v_ones=ones)
    if kind == "xview":
        return ops.attention(qkv, qkv, qkv, n_img=n, lq=L, lk=L, heads=8, head_dim=d, q_col0=0, k_col0=8 * dp, v_col0=16 * dp,
                             kv_map=kv_map, n_src=2, variant=variant, v_ones=ones)
    return ops.attention(qkv, txt, txt, n_img=n, lq=L, lk=106, heads=8, head_dim=d, q_col0=0, k_col0=0, v_col0=8 * dp,
                         q_cols=8 * dp, variant=variant, v_ones=ones)


def timed(kind, variant, reps=8):
    for _ in range(3):
        run(kind, variant)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(kind, variant); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


kinds = [os.environ["ONLY"]] if os.environ.get("ONLY") else ["self", "xview", "text"]
variants = [int(v) for v in os.environ["VARIANT"].split(",")] if os.environ.get("VARIANT") else [1, 2, 0]
ref = {}
for kind in kinds:
    lk, ns = (106, 1) if kind == "text" else (L, 2 if kind == "xview" else 1)
    for v in variants:
        t = timed(kind, v)
        out = run(kind, v).float()
        if kind not in ref:
            ref[kind] = out
        err = ((out - ref[kind]).abs().max() / ref[kind].abs().max()).item()
        print(f"attention d={d} {kind:5s} Lq={L} Lk={lk} n_img={n} variant={v}: {t * 1e3:8.1f} us  "
              f"{4.0 * n * L * lk * ns * 8 * d / t / 1e9:7.1f} TFLOP/s  max-rel diff vs first variant {err:.2e}", flush=True)
