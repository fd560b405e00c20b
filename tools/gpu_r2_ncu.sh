#!/bin/bash
# Round-2 ncu --set full captures (1 GPU): the new fused C = 32, k = 11 pair, the generic contraction kernel after the
# epilogue prefetch (flow / encoder launches of a step), the attention kernel.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rbpair_tc \
    -o gpurun_out/prof_rbpair_k11 -f python tools/bench_conv_tc.py --pair --reps 1 --profile --ks 11 --stages 3 > gpurun_out/ncu_pair.log 2>&1
tail -2 gpurun_out/ncu_pair.log
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc_kernel --launch-skip 20 --launch-count 6 \
    -o gpurun_out/prof_generic -f python tools/ncu_step.py --precision bf16 > gpurun_out/ncu_generic.log 2>&1
tail -2 gpurun_out/ncu_generic.log
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_tc --launch-count 1 \
    -o gpurun_out/prof_attention -f python tools/ncu_step.py --precision bf16 > gpurun_out/ncu_att.log 2>&1
tail -2 gpurun_out/ncu_att.log
ls -la gpurun_out/*.ncu-rep
