#!/bin/bash
# graph tests, then the default bench line three times in fresh processes (e2e stability with graph replay of repeated shapes)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_graphs_gpu.py tests/test_parity_gpu.py -x -q --timeout 300 2>&1 | tail -2
for v in 1 2 3; do
  timeout 300 python bench.py --no-cpu-baseline --no-gpu-incumbent > gpurun_out/bench_run$v.json 2> gpurun_out/bench_run$v.err
  python - $v <<'P'
import json, sys
d = json.load(open(f"gpurun_out/bench_run{sys.argv[1]}.json"))
print("RUN", sys.argv[1], round(d["value"]), "RT", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d.get("parity", {}).get("snr_db"), d["clocks"]["sm_mhz"], "fp16", round(d.get("fp16", {}).get("value", 0)), "launches", d.get("gpu_launches"), {k: round(x, 3) for k, x in d["time_by_class_ms_per_step"].items()})
P
done
