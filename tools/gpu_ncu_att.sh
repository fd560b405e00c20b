#!/bin/bash
# ncu --set full + source of one attention launch inside the bench step (launch index chosen past the warm-ups)
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc --launch-skip ${1:-20} --launch-count 1 \
    -o gpurun_out/prof_att -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-incumbent --no-front-end --no-parity > gpurun_out/ncu_att.log 2>&1
tail -3 gpurun_out/ncu_att.log
