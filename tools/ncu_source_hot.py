#!/usr/bin/env python
"""Top SASS instructions by stall samples from `ncu --page source --csv` output (file argument)."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
# cumulative by region: print every instruction with >= 0.4% of samples, in address order
for n, r in enumerate(data):
    s = int(r[ix["# Samples"]] or 0)
    if s >= tot * 0.004:
        main = max(stalls, key=lambda k: int(r[ix[k]] or 0))
        print(f"{n:5d} {s:6d} {100*s/tot:5.1f}% ex={r[ix['Instructions Executed']]:>10s} {main:22s} {r[ix['Source']][:90]}")
