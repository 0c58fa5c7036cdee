"""Summarise `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` by CUDA source line:
executed warp instructions and stall samples per line (top N)."""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, agg = None, None, {}
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].strip().isdigit() and len(r) > 8:
        d = dict(zip(hdr, r))
        key = (cur, int(r[0]))
        def num(v):
            try:
                return float(v)
            except (TypeError, ValueError):
                return 0.0
        ins = num(d.get("Instructions Executed"))
        smp = num(d.get("# Samples"))
        a = agg.setdefault(key, [0.0, 0.0, r[1].strip()[:90]])
        a[0] += ins
        a[1] += smp
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
print(f"total warp instructions {tot_i:.3e}, samples {tot_s:.0f}")
print("by samples:")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100 * a[1] / tot_s:5.1f}% smp {100 * a[0] / tot_i:5.1f}% ins  {f}:{ln}  {a[2]}")
