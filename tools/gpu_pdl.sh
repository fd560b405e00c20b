#!/bin/bash
# A/B: programmatic dependent launch with small shared-memory rings on the <= 2-wave launches (encoder / flow GEMMs).
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for cfg in "0 0" "1 0" "1 96" "0 96" "1 64"; do
  set -- $cfg
  RVCB200_PDL=$1 RVCB200_SMALL_SMEM_KB=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pdl$1_small$2.json 2> gpurun_out/bench_pdl$1_small$2.err
done
python - <<'P'
import json, glob
for n in sorted(glob.glob("gpurun_out/bench_pdl*.json")):
    try:
        d = json.load(open(n))
        print(n, round(d["value"]), "RT  ms", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["time_by_class_ms_per_step"].items()}, d["clocks"]["sm_mhz"])
    except Exception as e:
        print(n, "failed", e)
P
