#!/bin/bash
# Round-2 visit 12 (1 GPU): RMVPE with the st.async GRU exchange: tests, timing, launch list of one 60 s call.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_rmvpe_gpu.py -q -s --timeout 300 > gpurun_out/pytest_rmvpe_v2.log 2>&1
echo "rmvpe pytest rc=$?" | tee gpurun_out/status.txt; grep -E "passed|failed|Error|error|r[123]_|GRU|H=" gpurun_out/pytest_rmvpe_v2.log | cut -c1-260 | tail -40
timeout 300 python tools/bench_rmvpe.py --seconds 5,20,60 > gpurun_out/rmvpe_bench_v2.jsonl 2> gpurun_out/rmvpe_bench.err; cat gpurun_out/rmvpe_bench_v2.jsonl; tail -5 gpurun_out/rmvpe_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rmvpe_launches_60s_v2.csv python tools/rmvpe_step.py > gpurun_out/rmvpe_step.log 2>&1
tail -2 gpurun_out/rmvpe_step.log
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/rmvpe_launches_60s_v2.csv")) if len(r) > 10 and r[0].isdigit()]
n = len(rows) // 2
rows = rows[-n:]
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0][-40:]
    t = float(r[-1].replace(",", ""))
    tot[name][0] += 1; tot[name][1] += t
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:42s} x{c:4d} {t/1e3:9.1f} us")
print("total", sum(v[1] for v in tot.values()) / 1e3, "us in", n, "launches")
# the convolutions in launch order
for i, r in enumerate(rows):
    if "conv_tc" in r[4]:
        print(i, r[7], r[-1])
P
