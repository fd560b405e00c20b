#!/bin/bash
# Round-2 visit 23 (1 GPU): 32-column passes in the generic epilogue: whole GPU suite, then A/B (RVCB200_EPI32=0 = 16-column passes)
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_v23.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; tail -4 gpurun_out/pytest_v23.log
for g in 0 1; do
  RVCB200_EPI32=$g timeout 300 python tools/bench_rmvpe.py --seconds 5,60 --no-incumbent > gpurun_out/rmvpe_bench_e$g.jsonl 2>> gpurun_out/rmvpe_bench.err; cut -c1-120 gpurun_out/rmvpe_bench_e$g.jsonl
  RVCB200_EPI32=$g timeout 300 python tools/bench_hubert.py --seconds 5,60 > gpurun_out/hubert_bench_e$g.jsonl 2>> gpurun_out/hubert_bench.err; cut -c1-120 gpurun_out/hubert_bench_e$g.jsonl
  RVCB200_EPI32=$g timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-incumbent --no-front-end > gpurun_out/bench_e$g.json 2> gpurun_out/bench_e$g.err; echo "bench rc=$?"
  python - <<P
import json
d = json.load(open("gpurun_out/bench_e$g.json"))
print("EPI32=$g", round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], {k: round(v,3) for k,v in d["time_by_class_ms_per_step"].items()}, d["parity"]["snr_db"])
P
done
