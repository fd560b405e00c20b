#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --launch-skip N --launch-count 1`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = []
for r in rows[2:]:
    if len(r) != len(hdr) or r[0] == "Address":
        break                                  # only the first (SASS) section
    data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot)
agg = {}
for s in stalls:
    agg[s] = sum(int(r[ix[s]] or 0) for r in data)
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]
for r in top:
    n = int(r[ix["# Samples"]] or 0)
    why = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{n:7d} {100*n/tot:5.1f}%  {r[ix['Address']][-5:]}  {r[ix['Source']].strip()[:90]:90s} {why}")
