#!/bin/bash
# e2e stability of the default bench line: smoke first (another process, like the driver's order), then five default runs
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python __graft_entry__.py smoke > /dev/null 2>&1
for v in 1 2 3 4 5; do
  timeout 300 python bench.py --no-cpu-baseline --no-gpu-incumbent --no-front-end > gpurun_out/bench_chk$v.json 2> gpurun_out/bench_chk$v.err
  python - $v <<'P'
import json, sys
d = json.load(open(f"gpurun_out/bench_chk{sys.argv[1]}.json"))
print("RUN", sys.argv[1], round(d["value"]), "RT", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3), d["clocks"]["sm_mhz"], d["clocks"]["samples"])
P
done
