#!/bin/bash
# Round-2 visit 17 (1 GPU): as visit 16, with the ncu captures cut to a few launches (gpurun_out is limited to 64 MiB).
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rmvpe_launches_60s_v4.csv python tools/rmvpe_step.py > gpurun_out/rmvpe_step.log 2>&1
tail -1 gpurun_out/rmvpe_step.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rmvpe_gru --launch-skip 1 --launch-count 1 \
    -o gpurun_out/prof_rmvpe_gru -f python tools/rmvpe_step.py --seconds 20 > gpurun_out/ncu_gru.log 2>&1
tail -1 gpurun_out/ncu_gru.log
# second call's convolutions (131 per call): encoder L0 16 -> 16 (row slabs, resident weights), intermediate 512 -> 512 (grouped ring stages)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 134 --launch-count 2 \
    -o gpurun_out/prof_rmvpe_conv_L0 -f python tools/rmvpe_step.py > gpurun_out/ncu_rmvpe_conv_L0.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 178 --launch-count 2 \
    -o gpurun_out/prof_rmvpe_conv_L5 -f python tools/rmvpe_step.py > gpurun_out/ncu_rmvpe_conv_L5.log 2>&1
ls -la gpurun_out/*.ncu-rep
timeout 600 python tools/sweep.py --what song --front-end b200 --f0 rmvpe --tiers 60,38 --reps 3 > gpurun_out/song_1gpu_real_front_ends.jsonl 2> gpurun_out/song.err
tail -2 gpurun_out/song.err
timeout 1500 python bench.py > gpurun_out/bench_default_v17.json 2> gpurun_out/bench_default_v17.err; echo "bench rc=$?"
du -sh gpurun_out
