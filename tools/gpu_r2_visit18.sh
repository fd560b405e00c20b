#!/bin/bash
# Round-2 visit 18 (1 GPU): residual prefetch one tile ahead + bias staged once in the generic epilogue: whole GPU suite, timings
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_v18.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; tail -4 gpurun_out/pytest_v18.log
timeout 300 python tools/bench_rmvpe.py --seconds 5,60 --no-incumbent > gpurun_out/rmvpe_bench_v5.jsonl 2>> gpurun_out/rmvpe_bench.err; cat gpurun_out/rmvpe_bench_v5.jsonl
timeout 300 python tools/bench_hubert.py --seconds 5,60 > gpurun_out/hubert_bench_v5.jsonl 2>> gpurun_out/hubert_bench.err; cut -c1-200 gpurun_out/hubert_bench_v5.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-incumbent --no-front-end > gpurun_out/bench_v18.json 2> gpurun_out/bench_v18.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_v18.json"))
print(round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], {k: round(v,3) for k,v in d["time_by_class_ms_per_step"].items()}, d["parity"]["snr_db"])
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rmvpe_launches_60s_v5.csv python tools/rmvpe_step.py > gpurun_out/rmvpe_step.log 2>&1
python - <<'P'
import csv
rows = [r for r in csv.reader(open("gpurun_out/rmvpe_launches_60s_v5.csv")) if len(r) > 10 and r[0].isdigit()]
rows = rows[-(len(rows) // 2):]
print("convs:", [round(float(r[-1].replace(",", "")) / 1e3) for r in rows if "conv_tc" in r[4]])
P
