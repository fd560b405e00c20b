#!/bin/bash
# Round-2 visit 14 (1 GPU): row-slab mode of the generic kernel: RMVPE tests, generic-kernel op tests, RMVPE timing.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_rmvpe_gpu.py tests/test_tc_gpu.py tests/test_hubert_gpu.py -q -s --timeout 300 > gpurun_out/pytest_rmvpe_v4.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; grep -E "passed|failed|Error|error|r[123]_|H=" gpurun_out/pytest_rmvpe_v4.log | cut -c1-260 | tail -30
timeout 300 python tools/bench_rmvpe.py --seconds 5,20,60 --no-incumbent > gpurun_out/rmvpe_bench_v3.jsonl 2> gpurun_out/rmvpe_bench.err; cat gpurun_out/rmvpe_bench_v3.jsonl; tail -5 gpurun_out/rmvpe_bench.err
