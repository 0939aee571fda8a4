"""Per-shape gap to the roofline for one 8-scene CFG step: reads a per-shape timing table (`r01_step_shapes.txt`, written by
bench.py's instrumented pass) and prints, for every GEMM / attention / norm shape, the measured time, the lower bound from
the measured peaks (tensor: sustained bf16 1354.8 TFLOP/s; HBM: 6545 GB/s; attention at head_dim 40 additionally the MUFU
bound of 16 ex2/clk/SM) and the time that would be saved at the bound -- sorted by that saving.
    python profiles/gap_table.py profiles/r01_step_shapes.txt > profiles/r01_gap_table.txt"""
import re
import sys

TENSOR, HBM, SMS, CLK = 1354.8e12, 6545.3e9, 148, 1.70e9     # CLK: typical SM clock under the power cap on this pool
N_IMG = 96


def main(path):
    rows = []
    for line in open(path):
        p = line.split()
        if len(p) < 6:
            continue
        fam = p[0]
        tag = p[1] if p[2] == "x" else ""
        i = p.index("x")
        cnt, ms = int(p[i + 1]), float(p[i + 2])
        lb, why = None, ""
        if fam == "gemm_tcgen05":
            m = re.match(r"M(\d+)_N(\d+)_K(\d+)", tag)
            M, N, K = (int(x) for x in m.groups())
            fl = 2.0 * M * N * K
            by = 2.0 * (M * K + N * K + M * N)                # operands + output; residual / bias reads not counted
            if K >= 2880 and K % 9 == 0:                      # 3x3 conv (K = 9 x Cin): the activation is read once, not 9 times
                by = 2.0 * (M * K / 9 + N * K + M * N)
            t_t, t_h = fl / TENSOR, by / HBM
            lb, why = max(t_t, t_h) * 1e3, "tensor" if t_t >= t_h else "hbm"
        elif fam == "attn_tcgen05":
            m = re.match(r"d(\d+)_Lq(\d+)_Lk(\d+)_s(\d+)", tag)
            d, lq, lk, s = (int(x) for x in m.groups())
            heads = 8
            fl = 4.0 * N_IMG * lq * lk * heads * d * s
            by = 2.0 * heads * d * N_IMG * (2 * lq + 2 * lk * s)
            ex = N_IMG * heads * lq * (-(-lk // 64) * 64) * s     # exponentials incl. tile padding (64-key granularity)
            t_t, t_h, t_x = fl / TENSOR, by / HBM, ex / (16.0 * SMS * CLK)
            lb = max(t_t, t_h, t_x) * 1e3
            why = {t_t: "tensor", t_h: "hbm", t_x: "mufu"}[max(t_t, t_h, t_x)]
        elif fam == "groupnorm_silu":
            m = re.match(r"C(\d+)_HW(\d+)", tag)
            C, HW = int(m.group(1)), int(m.group(2))
            lb, why = 2.0 * 3 * N_IMG * HW * C / HBM * 1e3, "hbm"
        elif fam == "layernorm":
            C = int(tag[1:])
            rows_ = {320: 134400, 640: 33600}.get(C, 8736)
            lb, why = 4.0 * rows_ * C / HBM * 1e3, "hbm (rows approximated)"
        if lb is None:
            continue
        rows.append((ms - lb * cnt, fam, tag, cnt, ms, lb * cnt, why))
    rows.sort(reverse=True)
    tot_ms = sum(r[4] for r in rows)
    tot_lb = sum(r[5] for r in rows)
    print(f"{'family':16s} {'shape':28s} {'x':>3s} {'measured ms':>12s} {'bound ms':>9s} {'gap ms':>7s}  bound")
    for gap, fam, tag, cnt, ms, lb, why in rows:
        if gap < 0.05:
            continue
        print(f"{fam:16s} {tag:28s} {cnt:3d} {ms:12.3f} {lb:9.3f} {gap:7.3f}  {why}")
    print(f"\nsum over the listed families: measured {tot_ms:.2f} ms, bound {tot_lb:.2f} ms "
          f"(per launch max(tensor, HBM[, MUFU]) at the measured peaks; launch gaps and tails not modelled)")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "profiles/r01_step_shapes.txt")
