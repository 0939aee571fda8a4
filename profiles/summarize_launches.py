"""summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
per kernel family launches / time / share / DRAM traffic -> <stem>_summary.txt and <stem>_traffic.json"""
import collections, csv, json, sys

src = sys.argv[1]
stem = src[:-4] if src.endswith(".csv") else src
rows = list(csv.reader(open(src)))
hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hi]
ki, mi, vi, ui, idi = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
per = collections.defaultdict(dict)
for r in rows[hi + 1:]:
    if len(r) > vi:
        per[(r[idi], r[ki])][r[mi]] = (float(r[vi].replace(",", "")), r[ui])
SCALE = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for (_, k), m in per.items():
    name = k.split("(")[0].split("<")[0].replace("dd::", "").replace("void ", "").strip()
    f = fam[name]
    f[0] += 1
    for j, key in ((1, "gpu__time_duration.sum"), (2, "dram__bytes_read.sum"), (3, "dram__bytes_write.sum")):
        v, u = m.get(key, (0.0, "byte"))
        f[j] += v * SCALE.get(u, 1)
tot = sum(f[1] for f in fam.values())
lines = ["ncu launch list of ONE 8-scene CFG denoising step (profiles/step_once.py EAGER=1; --metrics gpu__time_duration.sum,",
         "dram__bytes_read.sum,dram__bytes_write.sum --clock-control none).  Per-launch times are cold-cache and serialised: compare SHARES.",
         f"launches {sum(f[0] for f in fam.values())}  total {tot / 1e3:.2f} ms"]
out = {}
for k, f in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k:28s} x{f[0]:4d} {f[1] / 1e3:8.3f} ms {100 * f[1] / tot:5.1f}%  dram read {f[2] / 1e9:6.2f} GB "
                 f"write {f[3] / 1e9:6.2f} GB  traffic/launch {(f[2] + f[3]) / f[0] / 1e6:7.1f} MB")
    out[k] = {"launches": f[0], "ms": f[1] / 1e3, "share": f[1] / tot, "traffic_bytes_per_launch": (f[2] + f[3]) / f[0]}
open(stem + "_summary.txt", "w").write("\n".join(lines) + "\n")
json.dump(out, open(stem + "_traffic.json", "w"), indent=1)
print("\n".join(lines[:8]))
