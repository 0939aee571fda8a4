"""Timeline of one CTA of attn_pp_kernel (SM clock at every hand-over between the softmax warpgroups and the issuing warps).
Builds a separate -DDD_ATTN_TRACE library (the shipped one carries no instrumentation) and prints, per key tile, how long each
role waited for the other.   python profiles/attn_trace.py [variant]"""
import ctypes as C, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
src = os.path.join(ROOT, "dualdiff_b200", "csrc")
out = "/tmp/libdualdiff_trace.so"
objs = []
PREBUILT = os.path.join(ROOT, "profiles", "ab", "lib_attn_trace.so")   # built on the CPU box: saves GPU-box minutes
if os.path.exists(PREBUILT):
    out = PREBUILT
for f in ([] if out == PREBUILT else sorted(os.listdir(src))):
    if f.startswith("dd_") and f.endswith(".cu"):
        o = f"/tmp/trace_{f[:-3]}.o"
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler",
                               "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-DDD_ATTN_TRACE", "-c", os.path.join(src, f), "-o", o])
        objs.append(o)
if out != PREBUILT:
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-shared", "-o", out] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"])
from dualdiff_b200 import _lib
_lib.LIB_PATH = out
import torch
from dualdiff_b200 import ops
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 2
TEXT = len(sys.argv) > 2 and sys.argv[2] == "text"     # text cross-attention: 106 keys (three 48-key tiles per item)
n, L, d, dp = 96, 1400, 40, 48
if TEXT:
    q = (torch.randn(n * L, 8 * dp, device="cuda") * 0.5).to(torch.bfloat16)
    kv = (torch.randn(n * 106, 16 * dp, device="cuda") * 0.5).to(torch.bfloat16)
    kv.view(n * 106, 2, 8, dp)[:, 1, :, 40] = 1.0
    run = lambda: ops.attention(q, kv, kv, n_img=n, lq=L, lk=106, heads=8, head_dim=d, q_col0=0, k_col0=0, v_col0=8 * dp, v_ones=True, variant=variant)
else:
    qkv = (torch.randn(n * L, 16 * dp + 8 * d, device="cuda") * 0.5).to(torch.bfloat16)
    run = lambda: ops.attention(qkv, qkv, qkv, n_img=n, lq=L, lk=L, heads=8, head_dim=d, q_col0=0, k_col0=8 * dp, v_col0=16 * dp, variant=variant)
lib = _lib.lib()
buf = (C.c_ulonglong * (2 * 10 * 4096))()
for _ in range(3):
    run()
torch.cuda.synchronize()
lib.dd_attn_trace_read(buf)          # clear
run()
torch.cuda.synchronize()
lib.dd_attn_trace_read(buf)
ev = [((buf[2 * i] >> 48) & 0xffff, (buf[2 * i] >> 32) & 0xffff, buf[2 * i] & 0xffffffff, buf[2 * i + 1]) for i in range(10 * 4096) if buf[2 * i + 1]]
ne = len(ev)
t0 = min(e[3] for e in ev)
names = {0: "wait S", 1: "got S", 2: "S in regs", 3: "exp done", 4: "PV(k-1) ok", 5: "P published", 6: "epilogue", 7: "last PV ok", 20: "O requested", 21: "O in regs", 22: "1/l", 23: "O stored", 24: "item done", 25: "advanced", 26: "row pointer", 10: "iter", 11: "s_free seen",
         12: "QK(k+1) issued", 13: "p_full seen", 14: "produce start", 15: "produce done", 16: "PV issued"}
print(f"{ne} events from one CTA; times in SM cycles")
if TEXT:   # raw hand-over sequence of one softmax warp: the per-item epilogue and the step to the next item
    seq = sorted([(t - t0, names.get(e, str(e))) for e, w, k, t in ev if w == 0])
    print("warp 0, first 70 events (cycle, event, +delta): " + " | ".join(f"{c} {nm} +{c - (seq[i - 1][0] if i else 0)}" for i, (c, nm) in enumerate(seq[:70])))
by = {}
for e, w, k, t in ev:
    by.setdefault((w, k), {})[e] = t - t0
for w in (0, 4, 8, 9):
    ks = sorted(k for (ww, k) in by if ww == w)
    print(f"--- warp {w} ({'softmax' if w < 8 else 'issuer'} tile {(w >> 2) if w < 8 else w - 8})")
    prev_end = None
    agg = {}
    for k in ks[:70]:
        d_ = by[(w, k)]
        line = f"k={k:3d} " + " ".join(f"{names[e].split()[0]}{e}@{d_[e]:7d}" for e in sorted(d_))
        if k < 40 or k % 10 == 0:
            print(line)
    # per-phase averages over the steady state
    import statistics
    def avg(e0, e1, shift=0):
        xs = [by[(w, k + shift)][e1] - by[(w, k)][e0] for k in ks[5:-5] if e0 in by[(w, k)] and (w, k + shift) in by and e1 in by[(w, k + shift)]]
        return statistics.mean(xs) if xs else float("nan")
    if w < 8:
        print(f"  avg cycles: wait for S {avg(0, 1):.0f} | TMEM load {avg(1, 2):.0f} | max+exp {avg(2, 3):.0f} | wait PV(k-1) {avg(3, 4):.0f} | "
              f"P store {avg(4, 5):.0f} | tile period {avg(0, 0, 1):.0f}")
    else:
        print(f"  avg cycles: iter start -> s_free seen {avg(10, 11):.0f} | QK issue (incl. kv_full wait) {avg(11, 12):.0f} | produce {avg(12, 15):.0f} | "
              f"wait p_full + PV issue {avg(15, 16):.0f} | loop {avg(16, 10, 1):.0f} | period {avg(10, 10, 1):.0f}")
