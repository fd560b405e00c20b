#!/bin/bash
# Round-2 visit 16 (1 GPU): RMVPE profile evidence (launch list, ncu --set full of the GRU recurrence and of the UNet convolutions),
# the song with both real front ends, the default bench line with every leg.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rmvpe_launches_60s_v4.csv python tools/rmvpe_step.py > gpurun_out/rmvpe_step.log 2>&1
tail -1 gpurun_out/rmvpe_step.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rmvpe_gru --launch-skip 1 --launch-count 1 \
    -o gpurun_out/prof_rmvpe_gru -f python tools/rmvpe_step.py --seconds 20 > gpurun_out/ncu_gru.log 2>&1
tail -2 gpurun_out/ncu_gru.log
# second call's convolutions: launch 4 (L0 16->16, row slabs), a deep one (intermediate 512->512, grouped ring stages)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 134 --launch-count 80 \
    -o gpurun_out/prof_rmvpe_convs -f python tools/rmvpe_step.py > gpurun_out/ncu_rmvpe_convs.log 2>&1
tail -2 gpurun_out/ncu_rmvpe_convs.log
ls -la gpurun_out/*.ncu-rep
timeout 600 python tools/sweep.py --what song --front-end b200 --f0 rmvpe --tiers 60,38 --reps 3 > gpurun_out/song_1gpu_real_front_ends.jsonl 2> gpurun_out/song.err
python - <<'P'
import json
for l in open("gpurun_out/song_1gpu_real_front_ends.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print({k: d[k] for k in ("front_end", "f0", "tier", "segments", "wall_s", "audio_s_per_s", "device_ms_max_over_ranks", "device_audio_s_per_s", "host_s_rank0")})
P
tail -3 gpurun_out/song.err
timeout 1500 python bench.py > gpurun_out/bench_default_v16.json 2> gpurun_out/bench_default_v16.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_default_v16.json"))
for k in ("value", "ms_per_step", "e2e", "parity", "front_end", "f0_front_end", "time_by_class_ms_per_step", "clocks", "gpu_launches", "cpu_baseline", "gpu_incumbent"): print(k, d.get(k))
print("fp16", d["fp16"]["value"], d["fp16"]["parity"]["snr_db"])
print({k: d["roofline"][k] for k in ("achieved", "frac", "frac_of_burst", "traffic", "hbm_frac", "avg_launch_ms", "launches_per_step")})
P
