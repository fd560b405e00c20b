#!/bin/bash
# attention kernel tests + end-to-end SNR gates + bench, then an ncu --set full of the small encoder/flow GEMM launches
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_tc_gpu.py -x -q -s --timeout 200 --timeout-method=thread -k "attention or snr or fused_pairs" > gpurun_out/test_att.log 2>&1
echo "attention/e2e tests rc=$?" | tee gpurun_out/status.txt; grep -h "attention_tc\|SNR" gpurun_out/test_att.log | tail -24; tail -3 gpurun_out/test_att.log
RVCB200_POST_TC=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16_nopost.json 2> gpurun_out/bench_bf16_nopost.err
python -c "
import json; d = json.load(open('gpurun_out/bench_bf16_nopost.json')); print('POST_TC=0', round(d['value']), 'RT ms', round(d['ms_per_step'], 3), d['time_by_class_ms_per_step'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_bf16.json"))
print(round(d["value"]), "RT  e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), d["time_by_class_ms_per_step"], d["clocks"])
P
if [ "$1" == "ncu" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"::conv_tc_kernel|noise_add16|attention_tc" --launch-skip 100 --launch-count 44 \
      -o gpurun_out/prof_small -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_small.log 2>&1
  tail -3 gpurun_out/ncu_small.log
fi
