#!/bin/bash
# Last visit of round 2 (short: the long one, tools/gpu_r2_final.sh, ran two commits earlier): full GPU suite, smoke, the default bench
# line without its CPU / eager-PyTorch legs (those are in profiles/r2_bench_default_final2.json), twice.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_r2_final3.log 2>&1
echo "pytest -m gpu rc=$?"; tail -2 gpurun_out/pytest_r2_final3.log
timeout 400 python __graft_entry__.py smoke > gpurun_out/smoke_r2_final3.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke_r2_final3.log
for v in a b; do
  timeout 600 python bench.py --no-cpu-baseline --no-gpu-incumbent > gpurun_out/bench_r2_final3$v.json 2> gpurun_out/bench_r2_final3$v.err; echo "bench rc=$?"
  python - $v <<'P'
import json, sys
d = json.load(open(f"gpurun_out/bench_r2_final3{sys.argv[1]}.json"))
print(round(d["value"], 1), "RT  e2e", round(d["e2e"]["value"], 1), "ms", round(d["ms_per_step"], 3), {k: round(x, 3) for k, x in d["time_by_class_ms_per_step"].items()}, d["clocks"], d["parity"]["snr_db"], "fp16", round(d["fp16"]["value"]), "launches", d["gpu_launches"])
print({k: d["roofline"][k] for k in ("achieved", "frac", "frac_of_burst", "traffic")}, {k: v for k, v in d["front_end"].items() if "ms" in k}, {k: v for k, v in d["f0_front_end"].items() if k == "ms_per_segment"})
P
done
