#!/bin/bash
# ncu --set full + source of <count> launches matching <regex> inside the bench step (skipping <skip> matches)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip ${2:-0} --launch-count ${3:-2} \
    -o gpurun_out/prof_$4 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-incumbent --no-front-end --no-parity > gpurun_out/ncu_$4.log 2>&1
tail -2 gpurun_out/ncu_$4.log
