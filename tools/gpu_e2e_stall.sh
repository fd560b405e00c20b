#!/bin/bash
# does the occasional ~48 ms stall of the e2e leg depend on graph replay?  K = 60 steps, per-step completion stamps
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in RVCB200_GRAPH_REPEAT=1 RVCB200_GRAPH_REPEAT=0 RVCB200_GRAPH_REPEAT=1 RVCB200_GRAPH_REPEAT=0 RVCB200_GRAPH_REPEAT=1 RVCB200_GRAPH_REPEAT=0 RVCB200_GRAPH_REPEAT=1 RVCB200_GRAPH_REPEAT=0; do
  env $v timeout 300 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --no-gpu-incumbent --no-front-end --no-parity --no-extra-precision > gpurun_out/bench_stall.json 2> gpurun_out/bench_stall.err
  python - "$v" <<'P'
import json, sys
d = json.load(open("gpurun_out/bench_stall.json"))
e = d["e2e"]
print("STALL", sys.argv[1], "value ms", round(d["ms_per_step"], 3), "e2e ms", round(e["ms_per_step"], 3), "median", round(e["step_ms_median"], 3), "max", round(e["step_ms_max"], 2), d["clocks"]["sm_mhz"])
P
done
