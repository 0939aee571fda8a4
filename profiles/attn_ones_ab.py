"""level-0 attention (d = 40, 1400 tokens, 96 images) with the softmax denominator taken from the column of ones in the padded V
heads (v_ones) against the kernel's own packed row sum, same process, same inputs, interleaved repetitions.
    python profiles/attn_ones_ab.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualdiff_b200 import ops
n, L, d, dp = 96, 1400, 40, 48
qkv = (torch.randn(n * L, 24 * dp, device="cuda") * 0.5).to(torch.bfloat16)
qkv.view(n * L, 24, dp)[:, 16:, d:] = 0
qkv.view(n * L, 24, dp)[:, 16:, d] = 1.0                      # V heads: column 40 = 1
txt = (torch.randn(n * 106, 16 * dp, device="cuda") * 0.5).to(torch.bfloat16)
txt.view(n * 106, 16, dp)[:, 8:, d:] = 0
txt.view(n * 106, 16, dp)[:, 8:, d] = 1.0
kv_map = torch.tensor([[(i // 6) * 6 + (i + 5) % 6, (i // 6) * 6 + (i + 1) % 6] for i in range(n)], dtype=torch.int32, device="cuda")


def run(kind, ones, variant=0):
    kw = dict(n_img=n, lq=L, heads=8, head_dim=d, v_hs=dp, v_ones=ones, variant=variant)
    if kind == "self":
        return ops.attention(qkv, qkv, qkv, lk=L, q_col0=0, k_col0=8 * dp, v_col0=16 * dp, **kw)
    if kind == "xview":
        return ops.attention(qkv, qkv, qkv, lk=L, q_col0=0, k_col0=8 * dp, v_col0=16 * dp, kv_map=kv_map, n_src=2, **kw)
    return ops.attention(qkv, txt, txt, lk=106, q_col0=0, k_col0=0, v_col0=8 * dp, q_cols=8 * dp, **kw)


for kind in ("self", "xview", "text"):
    ts = {False: [], True: []}
    for ones in (False, True):
        for _ in range(3):
            run(kind, ones)
    torch.cuda.synchronize()
    for rep in range(12):
        for ones in (False, True):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(kind, ones); e1.record()
            torch.cuda.synchronize()
            ts[ones].append(e0.elapsed_time(e1))
    a, b = run(kind, False).float(), run(kind, True).float()
    err = ((a - b).abs().max() / a.abs().max()).item()
    m = {k: sorted(v)[len(v) // 2] * 1e3 for k, v in ts.items()}
    print(f"attention d=40 {kind:5s}: own row sum {m[False]:7.1f} us | denominator from the ones column {m[True]:7.1f} us "
          f"({100 * (m[False] - m[True]) / m[False]:+.1f} %)  max-rel diff {err:.2e}", flush=True)
    # with the denominator from the ones column: 25 % of the exponentials on the FMA pipe (default) against all on MUFU
    tv = {0: [], 2: []}
    for v in (0, 2):
        for _ in range(3):
            run(kind, True, v)
    torch.cuda.synchronize()
    for rep in range(12):
        for v in (0, 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(kind, True, v); e1.record()
            torch.cuda.synchronize()
            tv[v].append(e0.elapsed_time(e1))
    mv = {k: sorted(v)[len(v) // 2] * 1e3 for k, v in tv.items()}
    print(f"               ones column, FMA-pipe share 25 % {mv[0]:7.1f} us | all exponentials on MUFU {mv[2]:7.1f} us", flush=True)
