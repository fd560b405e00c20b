#!/bin/bash
# Round-2 visit 6 (1 GPU): generic-epilogue prefetch (encoder / flow / HuBERT contractions) and 256-row split tiles at C = 256:
# tests, bench classes, short step, per-shape table of stage 1, HuBERT front end against the HuggingFace incumbent.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_v6.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; tail -6 gpurun_out/pytest_v6.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > gpurun_out/bench_bf16_v6.json 2> gpurun_out/bench_bf16_v6.err; echo "bench rc=$?" | tee -a gpurun_out/status.txt
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_bf16_v6.json"))
for k in ("value", "ms_per_step", "parity", "time_by_class_ms_per_step", "clocks", "gpu_launches"): print(k, d.get(k))
print({k: d["roofline"][k] for k in ("achieved", "frac", "frac_of_burst", "traffic", "hbm_frac", "avg_launch_ms", "launches_per_step")})
P
timeout 300 python tools/bench_conv_tc.py --reps 5 --rb 1 --stages 0 > gpurun_out/shapes_stage1_v6.jsonl 2> gpurun_out/shapes.err; cut -c1-260 gpurun_out/shapes_stage1_v6.jsonl
timeout 300 python tools/sweep.py --what sweep --reps 10 --max-frames 1000 > gpurun_out/sweep_v6.jsonl 2>> gpurun_out/sweep.err
grep '"batch": 1,' gpurun_out/sweep_v6.jsonl | cut -c1-150
timeout 600 python tools/bench_hubert.py --seconds 5,20,60 > gpurun_out/hubert_bench.jsonl 2> gpurun_out/hubert_bench.err; cat gpurun_out/hubert_bench.jsonl; tail -3 gpurun_out/hubert_bench.err
