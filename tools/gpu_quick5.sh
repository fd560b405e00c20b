#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_hubert_gpu.py tests/test_tc_gpu.py -x -q --timeout 300 -k "hubert or noise or inj or attention" 2>&1 | tail -2
for v in "X=0" "$@" "X=1"; do
  env $v timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-incumbent --no-front-end > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - "$v" <<'P'
import json, sys
d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
print(sys.argv[1], round(d["value"]), "RT  e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), "parity", round(d["parity"]["snr_db"], 2), {k: round(x, 3) for k, x in d["time_by_class_ms_per_step"].items()}, d["clocks"]["sm_mhz"])
P
done
