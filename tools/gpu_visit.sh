#!/bin/bash
# One GPU-box visit: tcgen05 kernel tests first, then the full GPU suite, per-shape micro-benchmark and bench lines.
# Usage: bash tools/gpu_visit.sh [ncu]     (ncu: also capture the launch list and an ncu --set full of the resblock kernel)
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_tc_gpu.py -x -q -s --timeout 120 > gpurun_out/test_tc.log 2>&1
rc=$?; echo "tc tests rc=$rc" | tee gpurun_out/status.txt; tail -5 gpurun_out/test_tc.log
grep -h "SNR" gpurun_out/test_tc.log | tail -20
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 240 --deselect tests/test_tc_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/status.txt
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_conv_tc.py --reps 5 --rb 1 > gpurun_out/shapes_rb1.jsonl 2> gpurun_out/shapes_rb1.err
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
timeout 300 python bench.py --precision fp16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
RVCB200_UPS_DENSE=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16_nodense.json 2> gpurun_out/bench_bf16_nodense.err
python - <<'P'
import json
for n in ("bench_bf16", "bench_fp16", "bench_bf16_nodense"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        print(n, round(d["value"]), "RT  e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), d["time_by_class_ms_per_step"], d["clocks"])
    except Exception as e:
        print(n, "failed", e)
P
if [ "$1" == "ncu" ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
  timeout 700 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rbconv_tc \
      -o gpurun_out/prof_rbconv -f python tools/bench_conv_tc.py --reps 1 --profile --ks 3,11 --stages 1,3 --rb 1 > gpurun_out/ncu_full.log 2>&1
fi
