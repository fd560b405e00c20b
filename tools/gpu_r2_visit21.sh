#!/bin/bash
# Round-2 visit 21 (1 GPU): RMVPE with pack pixels per 64-channel row on the wide levels: tests, timing, launch list
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_rmvpe_gpu.py -q -s --timeout 300 > gpurun_out/pytest_rmvpe_v7.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; grep -E "passed|failed|Error|error|r[1234]_" gpurun_out/pytest_rmvpe_v7.log | cut -c1-250 | tail -20
timeout 300 python tools/bench_rmvpe.py --seconds 5,20,60 --no-incumbent > gpurun_out/rmvpe_bench_v7.jsonl 2>> gpurun_out/rmvpe_bench.err; cat gpurun_out/rmvpe_bench_v7.jsonl; tail -3 gpurun_out/rmvpe_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rmvpe_launches_60s_v7.csv python tools/rmvpe_step.py > gpurun_out/rmvpe_step.log 2>&1
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/rmvpe_launches_60s_v7.csv")) if len(r) > 10 and r[0].isdigit()]
rows = rows[-(len(rows) // 2):]
print("convs:", [round(float(r[-1].replace(",", "")) / 1e3) for r in rows if "conv_tc" in r[4]])
print("total us", sum(float(r[-1].replace(",", "")) for r in rows) / 1e3)
P
