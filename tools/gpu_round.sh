#!/bin/bash
# One GPU-box visit: kernel tests first (bounded), then micro-benchmarks, full parity suite, bench lines, ncu captures.
# Usage: bash tools/gpu_round.sh [quick]
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_tc_gpu.py -x -q --timeout 120 -k "rbconv" > gpurun_out/test_rb.log 2>&1
rc=$?; echo "rb tests rc=$rc" | tee gpurun_out/status.txt; tail -5 gpurun_out/test_rb.log
if [ $rc -ne 0 ]; then
  timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/bench_conv_tc.py --T 200 --reps 1 --ks 3 --rb 1 > gpurun_out/sanitizer_rb.log 2>&1
  export RVCB200_RBCONV=0
fi
[ -z "$SKIP_SHAPES" ] && timeout 300 python tools/bench_conv_tc.py --reps 5 --rb 0 > gpurun_out/shapes_rb0.jsonl 2> gpurun_out/shapes_rb0.err
[ -z "$SKIP_SHAPES" ] && [ $rc -eq 0 ] && timeout 300 python tools/bench_conv_tc.py --reps 5 --rb 1 > gpurun_out/shapes_rb1.jsonl 2> gpurun_out/shapes_rb1.err
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 180 --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/status.txt
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
timeout 300 python bench.py --precision fp16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
[ "$1" == "quick" ] && { cat gpurun_out/bench_bf16.json; exit 0; }
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc \
    -o gpurun_out/prof_rbconv -f python tools/bench_conv_tc.py --reps 1 --profile --ks 3,11 --stages 1,3 --rb 1 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/bench_bf16.json
