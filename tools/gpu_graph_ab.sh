#!/bin/bash
# A/B: CUDA-graph replay of the whole 60 s step (RVCB200_GRAPH_FRAMES above B*T) against stream launches
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in 100000 2500 100000; do
  RVCB200_GRAPH_FRAMES=$v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-incumbent --no-front-end > gpurun_out/bench_graph$v.json 2> gpurun_out/bench_graph$v.err
  python - $v <<'P'
import json, sys
d = json.load(open(f"gpurun_out/bench_graph{sys.argv[1]}.json"))
print("GRAPH_FRAMES", sys.argv[1], round(d["value"]), "RT", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), d.get("parity", {}).get("snr_db"), d["clocks"]["sm_mhz"], "fp16", d.get("fp16", {}).get("value"))
P
done
