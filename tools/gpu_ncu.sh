#!/bin/bash
# ncu --set full captures: (1) resblock kernel at chosen shapes, (2) optionally a kernel regex inside bench.py.
# Usage: bash tools/gpu_ncu.sh "<bench_conv_tc args>" [kernel-regex-in-bench] [launch-count]
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rbconv_tc \
    -o gpurun_out/prof_rbconv -f python tools/bench_conv_tc.py --reps 1 --profile --rb 1 $1 > gpurun_out/ncu_rb.log 2>&1
tail -3 gpurun_out/ncu_rb.log
if [ -n "$2" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 --launch-skip ${4:-0} --launch-count ${3:-4} \
      -o gpurun_out/prof_bench -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  tail -3 gpurun_out/ncu_bench.log
fi
