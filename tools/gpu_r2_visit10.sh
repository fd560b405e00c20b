#!/bin/bash
# Round-2 visit 10 (1 GPU): uniform MMA issue + split epilogue in the generic kernel: tests, phase trace, bench, HuBERT, short sweep.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_v10.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; tail -12 gpurun_out/pytest_v10.log
timeout 200 python tools/trace_generic.py > gpurun_out/trace_generic_T6000_v2.jsonl 2> gpurun_out/trace.err; cut -c1-700 gpurun_out/trace_generic_T6000_v2.jsonl
timeout 200 python tools/trace_generic.py --T 100 > gpurun_out/trace_generic_T100_v2.jsonl 2>> gpurun_out/trace.err; cut -c1-200 gpurun_out/trace_generic_T100_v2.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err; echo "bench rc=$?" | tee -a gpurun_out/status.txt
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_v10.json"))
for k in ("value", "ms_per_step", "e2e", "parity", "front_end", "time_by_class_ms_per_step", "clocks", "gpu_launches"): print(k, d.get(k))
print("fp16", d["fp16"]["value"], d["fp16"]["parity"]["snr_db"])
print({k: d["roofline"][k] for k in ("achieved", "frac", "frac_of_burst", "traffic", "hbm_frac", "avg_launch_ms", "launches_per_step")})
P
timeout 300 python tools/bench_hubert.py --seconds 5,20,60 > gpurun_out/hubert_bench_v4.jsonl 2> gpurun_out/hubert_bench.err; cut -c1-330 gpurun_out/hubert_bench_v4.jsonl
timeout 300 python tools/sweep.py --what sweep --reps 10 --max-frames 1000 > gpurun_out/sweep_v10.jsonl 2>> gpurun_out/sweep.err
grep '"batch": 1,' gpurun_out/sweep_v10.jsonl | cut -c1-120
