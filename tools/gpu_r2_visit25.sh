#!/bin/bash
# Round-2 visit 25 (1 GPU): GRU recurrence on 16-CTA clusters (A/B: RVCB200_GRU_CLUSTER=8)
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_rmvpe_gpu.py -q -s --timeout 300 > gpurun_out/pytest_rmvpe_v9.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; grep -E "passed|failed|Error|error|r[1234]_|GRU" gpurun_out/pytest_rmvpe_v9.log | cut -c1-220 | tail -12
for c in 8 16; do
  RVCB200_GRU_CLUSTER=$c timeout 300 python tools/bench_rmvpe.py --seconds 5,20,60 --no-incumbent > gpurun_out/rmvpe_bench_cl$c.jsonl 2>> gpurun_out/rmvpe_bench.err; cat gpurun_out/rmvpe_bench_cl$c.jsonl
done
tail -3 gpurun_out/rmvpe_bench.err
