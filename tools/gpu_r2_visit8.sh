#!/bin/bash
# Round-2 visit 8 (1 GPU): source injection as an im2col GEMM (A/B), deeper rings of the generic contractions (A/B),
# acc_nostore in the fused pair; full tests first.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_v8.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; tail -12 gpurun_out/pytest_v8.log
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-incumbent --no-front-end --no-extra-precision > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<P
import json
d = json.load(open("gpurun_out/bench_$tag.json"))
c = d["time_by_class_ms_per_step"]
print("$tag", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "snr", round(d["parity"]["snr_db"], 2), {k: round(v, 3) for k, v in c.items()}, d["clocks"]["sm_mhz"], d["gpu_launches"] // 10)
P
}
run default RVCB200_X=0
run injgemm0 RVCB200_INJECT_GEMM=0
run rings_old RVCB200_GEN_NA=3 RVCB200_GEN_NB=10
run rings_6_20 RVCB200_GEN_NA=6 RVCB200_GEN_NB=20
run default2 RVCB200_X=0
timeout 300 python tools/sweep.py --what sweep --reps 10 --max-frames 1000 > gpurun_out/sweep_v8.jsonl 2>> gpurun_out/sweep.err
grep '"batch": 1,' gpurun_out/sweep_v8.jsonl | cut -c1-120
RVCB200_GEN_NA=3 RVCB200_GEN_NB=10 timeout 300 python tools/sweep.py --what sweep --reps 10 --max-frames 1000 > gpurun_out/sweep_v8_oldrings.jsonl 2>> gpurun_out/sweep.err
grep '"batch": 1,' gpurun_out/sweep_v8_oldrings.jsonl | cut -c1-120
timeout 300 python tools/bench_hubert.py --seconds 5,60 > gpurun_out/hubert_bench_v3.jsonl 2> gpurun_out/hubert_bench.err; cat gpurun_out/hubert_bench_v3.jsonl
