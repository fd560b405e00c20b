#!/bin/bash
# Round-2 visit 11 (1 GPU): first run of the RMVPE path: its GPU tests, timing against eager PyTorch, then the whole GPU suite.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_rmvpe_gpu.py -q -x -s --timeout 300 > gpurun_out/pytest_rmvpe_v1.log 2>&1
echo "rmvpe pytest rc=$?" | tee gpurun_out/status.txt; tail -40 gpurun_out/pytest_rmvpe_v1.log
timeout 300 python tools/bench_rmvpe.py --seconds 5,20,60 > gpurun_out/rmvpe_bench_v1.jsonl 2> gpurun_out/rmvpe_bench.err; cat gpurun_out/rmvpe_bench_v1.jsonl; tail -5 gpurun_out/rmvpe_bench.err
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 --deselect tests/test_rmvpe_gpu.py > gpurun_out/pytest_v11.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/status.txt; tail -5 gpurun_out/pytest_v11.log
