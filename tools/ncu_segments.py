#!/usr/bin/env python
"""Groups the SASS of `ncu --page source --csv` into runs of equal execution count and prints the
stall-sample share, instruction count and FP64 instruction count of each run."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.003
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
segs, cur = [], None
for n, r in enumerate(data):
    ex = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    if cur is None or abs(ex - cur["ex"]) > 0.02 * max(ex, cur["ex"], 1):
        cur = dict(start=n, ex=ex, s=0, n=0, fp64=0, lds=0, shfl=0, ldl=0)
        segs.append(cur)
    src = r[ix["Source"]]
    cur["s"] += s
    cur["n"] += 1
    cur["end"] = n
    cur["fp64"] += any(m in src for m in ("DFMA", "DADD", "DMUL", "MUFU.RCP64H", "DSETP"))
    cur["lds"] += ("LDS" in src) or ("STS" in src)
    cur["shfl"] += "SHFL" in src
    cur["ldl"] += ("LDL" in src) or ("STL" in src)
for s in segs:
    if s["s"] > tot * thr:
        print(f"{s['start']:5d}-{s['end']:5d} n={s['n']:4d} ex={s['ex']:>9d} samples={100*s['s']/tot:5.1f}% "
              f"fp64={s['fp64']:3d} smem={s['lds']:3d} shfl={s['shfl']:3d} local={s['ldl']:3d} "
              f"first={data[s['start']][ix['Source']][:50]}")
