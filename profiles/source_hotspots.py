#!/usr/bin/env python
"""Per-source-line hot spots of an ncu report: `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv | python profiles/source_hotspots.py [top]`.
Sums warp-stall samples and executed instructions of the SASS rows under each CUDA line."""
import csv
import sys
from collections import defaultdict
top = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rows = list(csv.reader(sys.stdin))


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0

hdr = next(r for r in rows if r and r[0] == "Line No")
ci = {n: i for i, n in enumerate(hdr)}
samp_i, inst_i = ci["Warp Stall Sampling (All Samples)"], ci["Instructions Executed"]
stall_cols = [(n, i) for n, i in ci.items() if n.startswith("stall_") and "Not Issued" not in n]
agg = defaultdict(lambda: [0, 0, "", defaultdict(int)])
cur = None
for r in rows:
    if len(r) < len(hdr) or r[0] == "Line No":
        continue
    if r[0]:
        cur = int(r[0]); agg[cur][2] = r[1].strip()
        continue
    if cur is None:
        continue
    a = agg[cur]
    a[0] += num(r[samp_i]); a[1] += num(r[inst_i])
    for n, i in stall_cols:
        a[3][n] += num(r[i])
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print(f"total samples {tot_s}  total warp instructions {tot_i}")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(a[3].items(), key=lambda kv: -kv[1])[:3]
    print(f"{ln:5d} {100 * a[0] / tot_s:5.1f}% smp {100 * a[1] / tot_i:5.1f}% inst  {' '.join(f'{n[6:]}:{v}' for n, v in st if v):40s} | {a[2][:110]}")
if len(sys.argv) > 2:   # extra: cumulative share by line ranges "a-b,c-d,..."
    for rg in sys.argv[2].split(","):
        lo, hi = map(int, rg.split("-"))
        s_ = sum(a[0] for ln, a in agg.items() if lo <= ln <= hi); i_ = sum(a[1] for ln, a in agg.items() if lo <= ln <= hi)
        print(f"lines {lo}-{hi}: {100 * s_ / tot_s:5.1f}% samples {100 * i_ / tot_i:5.1f}% instructions ({i_})")
