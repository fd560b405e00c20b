#!/bin/bash
# Round-2 visit 7 (1 GPU): HuBERT N-tile heuristic, pinned-buffer reuse in the song driver, full suite, default bench line
# (all legs, as the driver runs it), fp16 / fp32 bench lines, reference arm, launch list + DRAM traffic of the final build.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_v7.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; tail -5 gpurun_out/pytest_v7.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/status.txt; tail -3 gpurun_out/smoke.log
timeout 600 python tools/bench_hubert.py --seconds 5,20,60 > gpurun_out/hubert_bench_v2.jsonl 2> gpurun_out/hubert_bench.err; cat gpurun_out/hubert_bench_v2.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?" | tee -a gpurun_out/status.txt
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_default.json"))
for k in ("value", "ms_per_step", "e2e", "parity", "fp16", "front_end", "gpu_incumbent", "cpu_baseline", "time_by_class_ms_per_step", "clocks", "gpu_launches"): print(k, d.get(k))
print({k: d["roofline"][k] for k in ("achieved", "frac", "frac_of_burst", "traffic", "hbm_frac", "avg_launch_ms", "launches_per_step")})
P
timeout 300 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-incumbent --no-front-end > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
python -c "
import json; d=json.load(open('gpurun_out/bench_fp32.json')); print('fp32', d['value'], d['ms_per_step'], d['parity'])"
timeout 300 python tools/sweep.py --what song --reps 3 --tiers 60 > gpurun_out/song_1gpu_v3.jsonl 2> gpurun_out/song.err; cut -c1-700 gpurun_out/song_1gpu_v3.jsonl
timeout 300 python tools/sweep.py --what config3 --reps 5 > gpurun_out/config3.jsonl 2>> gpurun_out/song.err; cat gpurun_out/config3.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/step_bf16.csv python tools/ncu_step.py --precision bf16 > gpurun_out/ncu_step_bf16.log 2>&1
tail -1 gpurun_out/ncu_step_bf16.log
