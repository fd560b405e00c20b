#!/bin/bash
# One GPU-box visit: A/B the conv_tc variants, then parity tests, bench lines, ncu launch list + full capture
# with the newest variant that passes.  Every step is bounded by its own timeout.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
GOOD=""
for v in $VARIANTS; do
  lib=$PWD/build_variants/librvcb200_$v.so
  RVCB200_LIB=$lib timeout 240 python -m pytest tests/test_tc_gpu.py -x -q --timeout 90 > gpurun_out/tc_$v.log 2>&1
  rc=$?
  echo "variant $v test_tc rc=$rc" | tee -a gpurun_out/variants.txt
  if [ $rc -eq 0 ]; then
    RVCB200_LIB=$lib timeout 200 python tools/bench_conv_tc.py --reps 5 --ks 3,11 > gpurun_out/shapes_$v.jsonl 2> gpurun_out/shapes_$v.err
    echo "variant $v shapes rc=$?" | tee -a gpurun_out/variants.txt
    GOOD=$v
  else
    RVCB200_LIB=$lib timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/bench_conv_tc.py --T 40 --reps 1 --ks 3 \
        > gpurun_out/sanitizer_$v.log 2>&1
  fi
done
echo "GOOD=$GOOD" | tee -a gpurun_out/variants.txt
[ -z "$GOOD" ] && exit 0
export RVCB200_LIB=$PWD/build_variants/librvcb200_$GOOD.so
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 180 --durations=25 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --precision fp16 --steps 10 --warmup 3 > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
timeout 300 python bench.py --precision bf16 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --precision fp16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc \
    -o gpurun_out/prof_conv_tc -f python tools/bench_conv_tc.py --reps 1 --profile --ks 3,11 --stages 1,3 > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/bench_fp16.json
