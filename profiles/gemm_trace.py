"""Timeline of one CTA of gemm_tcgen05_kernel (SM clock at the hand-overs between the TMA producer, the UMMA issuer and two
epilogue warps) for the short-K plain GEMMs.  Uses profiles/ab/lib_trace.so (a -DDD_GEMM_TRACE build of dd_gemm.cu; the
shipped library carries no instrumentation).   python profiles/gemm_trace.py"""
import ctypes as C, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dualdiff_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "profiles", "ab", "lib_trace.so")
import torch
from dualdiff_b200 import ops
lib = _lib.lib()
NW, NE = 10, 2048
buf = (C.c_ulonglong * (2 * NW * NE))()
names = {8: "transposed", 9: "chunk stored", 0: "wait tfull", 1: "tfull", 2: "staging free", 3: "tmem in regs", 4: "half0 done", 5: "half1 done", 6: "fenced", 7: "store issued",
         10: "wait tempty", 11: "tempty", 12: "first stage full", 13: "tile committed", 20: "tile start", 21: "first slot free", 22: "loads issued"}


def trace(label, M, N, K, res, conv=None):
    """conv = (n_img, H, W): 3x3 implicit GEMM over the padded layout with a per-image vector (time embedding) + residual"""
    w = (torch.randn(N, K * (9 if conv else 1), device="cuda") * 0.05).to(torch.bfloat16)
    b = torch.randn(N, device="cuda")
    if conv:
        n, H, W = conv
        a = (torch.randn(ops.padded_rows(n, H, W), K, device="cuda") * 0.5).to(torch.bfloat16)
        r = (torch.randn(n * H * W, N, device="cuda") * 0.5).to(torch.bfloat16) if res else None
        rv = torch.randn(n, N, device="cuda")
        run = lambda: ops.gemm(a, w, bias=b, res1=r, rowvec=rv, rows_per_img=H * W, taps=9, conv_hw=(H, W), n_img=n)
    else:
        a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
        r = (torch.randn(M, N, device="cuda") * 0.5).to(torch.bfloat16) if res else None
        run = lambda: ops.gemm(a, w, bias=b, res1=r)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        run()
    flush.zero_()
    torch.cuda.synchronize()
    lib.dd_gemm_trace_read(buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    print(f"launch: {e0.elapsed_time(e1) * 1e3:.1f} us (instrumented)")
    lib.dd_gemm_trace_read(buf)
    raw = [((buf[2 * i] >> 48) & 0xffff, (buf[2 * i] >> 32) & 0xffff, buf[2 * i] & 0xffffffff, buf[2 * i + 1]) for i in range(NW * NE) if buf[2 * i + 1]]
    ev = [e for e in raw if not (e[3] >> 62) & 1]
    vals = {}
    for e, wp, t, v in raw:
        if (v >> 62) & 1:
            vals.setdefault(e, []).append(v & ((1 << 62) - 1))
    for e, xs in sorted(vals.items()):
        print(f"   issuer, cycles per tile spent waiting on the {'activation' if e == 14 else 'weight'} ring: " + " ".join(str(x) for x in xs[:12]))
    t0 = min(e[3] for e in ev)
    print(f"=== {label}: {len(ev)} events of CTA 6 (leader of pair 3), SM cycles")
    for wp in sorted(set(e[1] for e in ev)):
        seq = [(e[0], e[2], e[3] - t0) for e in ev if e[1] == wp]
        seq.sort(key=lambda x: x[2])
        role = "producer" if wp == 0 else "issuer" if wp == 1 else f"epilogue q{wp & 3} half{(wp - 2) >> 2}"
        if wp not in (0, 1, 2, 6):
            continue
        print(f"--- warp {wp} ({role}); first 40 events: " + " ".join(f"{e}:{t}@{c}" for e, t, c in seq[:40]))
        # per-transition averages over the steady state
        d = {}
        for (e0, _, c0), (e1, _, c1) in zip(seq[20:-10], seq[21:-9]):
            d.setdefault((e0, e1), []).append(c1 - c0)
        for (e0, e1), xs in sorted(d.items()):
            print(f"   {names[e0]:>16s} -> {names[e1]:<16s} n={len(xs):3d} mean {statistics.mean(xs):7.0f} median {statistics.median(xs):7.0f}")
        starts = [c for e, _, c in seq if e in (0, 10, 20)]
        if len(starts) > 12:
            per = [b - a for a, b in zip(starts[5:-3], starts[6:-2])]
            print(f"   tile period mean {statistics.mean(per):.0f} cycles over {len(per)} tiles")


if len(sys.argv) > 1 and sys.argv[1] == "conv":
    trace("conv L0 96 x 28x50, 320 -> 320, + time vector", 0, 320, 320, False, conv=(96, 28, 50))
    trace("conv L0 96 x 28x50, 320 -> 320, + time vector + residual", 0, 320, 320, True, conv=(96, 28, 50))
    trace("conv L1 96 x 14x25, 640 -> 640", 0, 640, 640, False, conv=(96, 14, 25))
    trace("M33600 N640 K640 + residual", 33600, 640, 640, True)
    sys.exit(0)
trace("M134400 N320 K320 no residual", 134400, 320, 320, False)
trace("M134400 N320 K320 + residual", 134400, 320, 320, True)
trace("M134400 N320 K40", 134400, 320, 40, False)
trace("M134400 N1152 K320", 134400, 1152, 320, False)
