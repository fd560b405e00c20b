#!/bin/bash
# Round-2 visit 19 (1 GPU): GRU wait without the cluster-scope acquire: RMVPE tests + timing
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_rmvpe_gpu.py -q -s --timeout 300 > gpurun_out/pytest_rmvpe_v6.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; grep -E "passed|failed|r[1234]_|GRU" gpurun_out/pytest_rmvpe_v6.log | cut -c1-250
timeout 300 python tools/bench_rmvpe.py --seconds 5,20,60 > gpurun_out/rmvpe_bench_v6.jsonl 2>> gpurun_out/rmvpe_bench.err; cat gpurun_out/rmvpe_bench_v6.jsonl
