#!/bin/bash
# One GPU-box visit for the fused ResBlock-pair kernel: op-level bit-exactness tests first (bounded), then the
# per-shape micro-benchmark, bench lines with and without fusion, and the launch list of the fused step.
# Usage: bash tools/gpu_pair.sh [ncu]
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_tc_gpu.py -x -q --timeout 90 --timeout-method=thread -k "rbpair or rbconv or conv_tc or attention" > gpurun_out/test_pair.log 2>&1
rc=$?; echo "pair op tests rc=$rc" | tee gpurun_out/status.txt; tail -15 gpurun_out/test_pair.log
if [ $rc -ne 0 ]; then
  timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_tc_gpu.py -x -q --timeout 250 --timeout-method=thread -k "rbpair_tc and pair_c32_k3_d1 and fp16-s" > gpurun_out/sanitizer_pair.log 2>&1
  tail -30 gpurun_out/sanitizer_pair.log
  exit 0
fi
timeout 900 python -m pytest tests/test_tc_gpu.py -x -q -s --timeout 200 --timeout-method=thread -k "fused_pairs or snr" > gpurun_out/test_pair_e2e.log 2>&1
echo "pair e2e tests rc=$?" | tee -a gpurun_out/status.txt; tail -5 gpurun_out/test_pair_e2e.log
RVCB200_PAIR_CFG=0 timeout 300 python tools/bench_conv_tc.py --pair --reps 5 > gpurun_out/pairs_cfg0.jsonl 2> gpurun_out/pairs.err
RVCB200_PAIR_CFG=1 timeout 300 python tools/bench_conv_tc.py --pair --reps 5 > gpurun_out/pairs_cfg1.jsonl 2>> gpurun_out/pairs.err
python - <<'P'
import json
a = [json.loads(l) for l in open("gpurun_out/pairs_cfg0.jsonl")]
b = [json.loads(l) for l in open("gpurun_out/pairs_cfg1.jsonl")]
for x, y in zip(a, b):
    print(f"C={x['C']} k={x['k']} d={x['dil']} {x['kind']}: cfg0 {x['fused_us']} us  cfg1 {y['fused_us']} us  two-launch {x['two_us']} us   (cfg0 {x['fused_hbm_gbs']} GB/s {x['fused_tflops']} TF)")
P
[ -z "$SKIP_SHAPES" ] && timeout 300 python tools/bench_conv_tc.py --reps 5 --rb 1 > gpurun_out/shapes_rb1.jsonl 2> gpurun_out/shapes_rb1.err
RVCB200_PAIR_CFG=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16_cfg0.json 2> gpurun_out/bench_bf16_cfg0.err
RVCB200_PAIR_CFG=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16_cfg1.json 2> gpurun_out/bench_bf16_cfg1.err
python - <<'P'
import json
for n in ("bench_bf16_cfg0", "bench_bf16_cfg1"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        print(n, round(d["value"]), "RT  e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), d["time_by_class_ms_per_step"], d["clocks"])
    except Exception as e:
        print(n, "failed", e)
P
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rbpair_tc \
      -o gpurun_out/prof_rbpair -f python tools/bench_conv_tc.py --pair --reps 1 --profile --ks 3 > gpurun_out/ncu_pair.log 2>&1
  tail -3 gpurun_out/ncu_pair.log
fi
if [ "$1" == "launches" ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
fi
