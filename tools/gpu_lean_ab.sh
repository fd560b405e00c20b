#!/bin/bash
# lean generic epilogue: full GPU suite with it on, then A/B of the bench line and of the two front ends
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 800 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_lean.log 2>&1; tail -3 gpurun_out/pytest_lean.log
for v in RVCB200_LEAN_EPI=0 RVCB200_LEAN_EPI=1 RVCB200_LEAN_EPI=0 RVCB200_LEAN_EPI=1; do
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-incumbent > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - "$v" <<'P'
import json, sys
d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
fe = d.get("front_end", {}); f0 = d.get("f0_front_end", {})
print(sys.argv[1], round(d["value"]), "RT  e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "parity", round(d["parity"]["snr_db"], 2), {k: round(x, 3) for k, x in d["time_by_class_ms_per_step"].items()}, d["clocks"]["sm_mhz"], "hubert", {k: v for k, v in fe.items() if "ms" in k}, "rmvpe", {k: v for k, v in f0.items() if "ms" in k})
P
done
