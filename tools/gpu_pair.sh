#!/bin/bash
# One GPU-box visit for the fused ResBlock-pair kernel: op-level bit-exactness tests first (bounded), then the
# per-shape micro-benchmark, bench lines with and without fusion, and the launch list of the fused step.
# Usage: bash tools/gpu_pair.sh [ncu]
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_tc_gpu.py -x -q --timeout 90 --timeout-method=thread -k "rbpair" > gpurun_out/test_pair.log 2>&1
rc=$?; echo "pair op tests rc=$rc" | tee gpurun_out/status.txt; tail -15 gpurun_out/test_pair.log
if [ $rc -ne 0 ]; then
  timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_tc_gpu.py -x -q --timeout 250 --timeout-method=thread -k "rbpair_tc and pair_c32_k3_d1 and fp16-s" > gpurun_out/sanitizer_pair.log 2>&1
  tail -30 gpurun_out/sanitizer_pair.log
  exit 0
fi
timeout 900 python -m pytest tests/test_tc_gpu.py -x -q -s --timeout 200 --timeout-method=thread -k "fused_pairs or snr" > gpurun_out/test_pair_e2e.log 2>&1
echo "pair e2e tests rc=$?" | tee -a gpurun_out/status.txt; tail -5 gpurun_out/test_pair_e2e.log
timeout 300 python tools/bench_conv_tc.py --pair --reps 5 > gpurun_out/pairs.jsonl 2> gpurun_out/pairs.err; cat gpurun_out/pairs.jsonl
RVCB200_FUSE_PAIRS=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16_nofuse.json 2> gpurun_out/bench_bf16_nofuse.err
RVCB200_FUSE_PAIRS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16_fuse.json 2> gpurun_out/bench_bf16_fuse.err
python - <<'P'
import json
for n in ("bench_bf16_nofuse", "bench_bf16_fuse"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        print(n, round(d["value"]), "RT  e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), d["time_by_class_ms_per_step"], d["clocks"])
    except Exception as e:
        print(n, "failed", e)
P
if [ "$1" == "ncu" ]; then
  RVCB200_FUSE_PAIRS=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
fi
