#!/bin/bash
# Round-end visit: full GPU suite, smoke, bench lines (bf16 / fp16 / fp32 / reference arm), per-shape micro-benchmarks,
# ncu --set full of the fused-pair and resblock kernels, launch list of a bench step.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q --timeout 240 --timeout-method=thread > gpurun_out/pytest_gpu_full.log 2>&1
echo "pytest -m gpu rc=$?" | tee gpurun_out/status.txt; tail -2 gpurun_out/pytest_gpu_full.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/status.txt; tail -1 gpurun_out/smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
timeout 300 python bench.py --precision fp16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
timeout 300 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python - <<'P'
import json
for n in ("bench_bf16", "bench_fp16", "bench_fp32", "bench_reference"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        print(n, round(d["value"], 1), "RT  e2e", round(d["e2e"]["value"], 1), "ms", round(d.get("ms_per_step", 0), 3), d.get("time_by_class_ms_per_step"), d.get("clocks"), d.get("cpu_baseline"))
    except Exception as e:
        print(n, "failed", e)
P
timeout 300 python tools/bench_conv_tc.py --pair --reps 5 > gpurun_out/pairs.jsonl 2> gpurun_out/pairs.err
timeout 300 python tools/bench_conv_tc.py --reps 5 --rb 1 > gpurun_out/shapes_rb1.jsonl 2> gpurun_out/shapes_rb1.err
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rbpair_tc \
    -o gpurun_out/prof_rbpair -f python tools/bench_conv_tc.py --pair --reps 1 --profile --ks 3 > gpurun_out/ncu_pair.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rbconv_tc \
    -o gpurun_out/prof_rbconv -f python tools/bench_conv_tc.py --reps 1 --profile --ks 3,11 --stages 1,3 --rb 1 > gpurun_out/ncu_rb.log 2>&1
if [ "$1" == "launches" ]; then
  timeout 450 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
fi
ls -la gpurun_out | head -40
