#!/usr/bin/env python
"""Per-region executed-instruction / stall-sample / shared-wavefront breakdown of one kernel from an ncu report's source page.
usage: python profiles/region_breakdown.py REPORT.ncu-rep name:lo_hex:hi_hex ...   (offsets relative to the kernel's first instruction)"""
import csv
import io
import subprocess
import sys


def main(path, regs):
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]
    ia, ie, isamp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
    iw, iwi = h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Ideal")
    regs = [(n, int(lo, 16), int(hi, 16)) for n, lo, hi in (r.split(":") for r in regs)]
    tot = {n: [0, 0, 0, 0] for n, _, _ in regs}
    base = None
    for r in rows[2:]:
        try:
            a, n, s = int(r[ia], 16), int(r[ie]), int(r[isamp])
        except (ValueError, IndexError):
            continue
        if base is None:
            base = a
        a -= base
        for nm, lo, hi in regs:
            if lo <= a < hi:
                t = tot[nm]
                t[0] += n; t[1] += s; t[2] += int(r[iw] or 0); t[3] += int(r[iwi] or 0)
    T = sum(v[0] for v in tot.values()); S = sum(v[1] for v in tot.values())
    print(f"# {path}: {T} warp instructions, {S} stall samples")
    for k, v in tot.items():
        print(f"{k:12s} inst {v[0] / 1e6:10.1f} M {v[0] / T * 100:5.1f} %   samples {v[1] / max(S, 1) * 100:5.1f} %   smem wavefronts {v[2] / 1e6:9.1f} M (ideal {v[3] / 1e6:9.1f} M)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
