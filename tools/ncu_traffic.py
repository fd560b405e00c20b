#!/usr/bin/env python
"""Join an ncu per-launch CSV of ONE step (tools/ncu_step.py) with the step's kernel classes and write the per-class
summary the bench line quotes: profiles/r2_dram_traffic.json (+ a readable .md).

    python tools/ncu_traffic.py gpurun_out/step_bf16.csv gpurun_out/step_classes_bf16.json [--tag r2_v1]
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASS_NAMES = ["dec_resblocks", "attention", "sine_source", "glue", "dec_pre_ups", "flow", "enc_linear", "unused"]


def main():
    csv_path, cls_path = sys.argv[1], sys.argv[2]
    tag = sys.argv[sys.argv.index("--tag") + 1] if "--tag" in sys.argv else "r2"
    meta = json.load(open(cls_path))
    lines = [l for l in open(csv_path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()            # launch id -> {metric: value}
    for r in rows:
        full = re.sub(r"\(.*", "", r["Kernel Name"])
        d = per.setdefault(r["ID"], {"name": full.split("::")[-1] if full.startswith(("rvc::", "void rvc::")) else "at::" + full.split("::")[-1], "grid": r["Grid Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        if r["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if r["Metric Name"] == "gpu__time_duration.sum":
            v *= {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1, "us": 1, "msecond": 1e3, "ms": 1e3}.get(unit, 1e-3)   # -> us
        d[r["Metric Name"]] = v
    launches = [l for l in per.values() if not l["name"].startswith("at::")]   # this repo's kernels only
    kernels = meta.get("kernels") or [1] * meta["launches"]
    assert len(launches) == sum(kernels), (len(launches), sum(kernels))
    cls_of, ev_of = [], []
    for c, k, ev in zip(meta["cls"], kernels, meta["event_ms"]):      # a scope of k kernels: its event time is split evenly
        cls_of += [c] * k
        ev_of += [ev / max(k, 1)] * k
    by_cls = collections.OrderedDict()
    by_kernel = collections.OrderedDict()
    for l, c, ev in zip(launches, cls_of, ev_of):
        a = by_cls.setdefault(CLASS_NAMES[c], {"launches": 0, "read": 0.0, "write": 0.0, "ncu_us": 0.0, "event_us": 0.0})
        k = by_kernel.setdefault((CLASS_NAMES[c], l["name"], l["grid"]), {"launches": 0, "read": 0.0, "write": 0.0, "ncu_us": 0.0, "event_us": 0.0})
        for t in (a, k):
            t["launches"] += 1
            t["read"] += l.get("dram__bytes_read.sum", 0.0)
            t["write"] += l.get("dram__bytes_write.sum", 0.0)
            t["ncu_us"] += l.get("gpu__time_duration.sum", 0.0)
            t["event_us"] += ev * 1e3
    out_path = os.path.join(ROOT, "profiles", "r2_dram_traffic.json")
    allp = json.load(open(out_path)) if os.path.exists(out_path) else {}
    rb = by_cls["dec_resblocks"]
    allp[meta["precision"]] = {
        "bytes_per_step": rb["read"] + rb["write"], "read_bytes_per_step": rb["read"], "write_bytes_per_step": rb["write"],
        "launches_per_step": rb["launches"], "workload": f"{meta['config']} B={meta['B']} T={meta['T']}",
        "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one step "
                  f"(tools/ncu_step.py --precision {meta['precision']}), joined with the step's launch classes "
                  f"(tools/ncu_traffic.py, capture {tag}); profiles/{tag}_step_{meta['precision']}.md",
        "by_class": by_cls,
    }
    json.dump(allp, open(out_path, "w"), indent=1)
    tot = sum(v["ncu_us"] for v in by_cls.values())
    with open(os.path.join(ROOT, "profiles", f"{tag}_step_{meta['precision']}.md"), "w") as f:
        f.write(f"# ncu launch list of one step, {meta['config']} B={meta['B']} T={meta['T']} {meta['precision']} ({tag})\n\n")
        f.write("ncu times are cold-cache and serialised: compare SHARES.  event_us = the library's own CUDA-event pair around the launch "
                "in an unprofiled step.\n\n| class | launches | ncu us | share | event us | DRAM read MB | DRAM write MB |\n|---|---|---|---|---|---|---|\n")
        for n, v in by_cls.items():
            f.write(f"| {n} | {v['launches']} | {v['ncu_us']:.1f} | {100 * v['ncu_us'] / tot:.1f} % | {v['event_us']:.1f} | {v['read'] / 1e6:.1f} | {v['write'] / 1e6:.1f} |\n")
        f.write(f"| total | {len(launches)} | {tot:.1f} | | {sum(v['event_us'] for v in by_cls.values()):.1f} | | |\n\n")
        f.write("| class | kernel | grid | n | ncu us (avg) | event us (avg) | read MB (avg) | write MB (avg) | DRAM GB/s (ncu time) |\n|---|---|---|---|---|---|---|---|---|\n")
        for (c, n, g), v in sorted(by_kernel.items(), key=lambda kv: -kv[1]["ncu_us"]):
            k = v["launches"]
            f.write(f"| {c} | {n} | {g} | {k} | {v['ncu_us'] / k:.1f} | {v['event_us'] / k:.1f} | {v['read'] / k / 1e6:.1f} | {v['write'] / k / 1e6:.1f} | "
                    f"{(v['read'] + v['write']) / max(v['ncu_us'], 1e-9) / 1e3:.0f} |\n")
    print(json.dumps({k: {kk: round(vv, 1) for kk, vv in v.items()} for k, v in by_cls.items()}, indent=1))


if __name__ == "__main__":
    main()
