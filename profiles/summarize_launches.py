"""Summarise an ncu launch list (gpu__time_duration.sum per launch) of one bench step by kernel/grid."""
import collections, csv, json, re, sys
path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [x["Kernel Name"] for x in rows]
idx = [i for i, n in enumerate(names) if "len_to_i32" in n]
last = rows[idx[-2]:idx[-1]] if len(idx) >= 2 else rows[idx[-1]:]       # the last device-resident timed step
agg = collections.OrderedDict()
for x in last:
    n = re.sub(r"\(.*", "", x["Kernel Name"]).split("::")[-1]
    key = (n, x["Grid Size"], x["Block Size"])
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += float(x["Metric Value"])
tot = sum(v[1] for v in agg.values())
print(f"launches {len(last)}  total {tot/1e3:.1f} us (serialised, cold-cache: compare SHARES, not absolutes)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{v[1]/1e3:9.1f} us {100*v[1]/tot:5.1f}%  n={v[0]:3d}  avg {v[1]/v[0]/1e3:8.1f} us  {k[0]} grid={k[1]} block={k[2]}")
