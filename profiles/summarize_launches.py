#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total and share per kernel.
usage: python profiles/summarize_launches.py gpurun_out/launches_x.csv > profiles/rNN_launches_x.txt"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    n = 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        a = agg.setdefault(k, [0, 0.0, 0.0])
        a[0] += 1; a[1] += v; a[2] = max(a[2], v); n += 1
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {n} launches, {tot:.0f} us of device time (cold-cache, serialised under ncu: compare shares)")
    print(f"{'kernel':64s} {'n':>5s} {'sum_us':>11s} {'share':>7s} {'max_us':>9s}")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:64]:64s} {a[0]:5d} {a[1]:11.1f} {a[1] / tot * 100:6.1f}% {a[2]:9.1f}")


if __name__ == "__main__":
    main(sys.argv[1])
