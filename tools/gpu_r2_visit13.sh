#!/bin/bash
# Round-2 visit 13 (1 GPU): RMVPE tests after the decode fix (float32 weight sum like numpy), phase trace of the deep UNet convolutions.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_rmvpe_gpu.py -q -s --timeout 300 > gpurun_out/pytest_rmvpe_v3.log 2>&1
echo "rmvpe pytest rc=$?" | tee gpurun_out/status.txt; grep -E "passed|failed|Error|error|r[123]_|GRU|H=" gpurun_out/pytest_rmvpe_v3.log | cut -c1-260 | tail -40
timeout 200 python tools/trace_generic.py --rmvpe --T 6000 > gpurun_out/trace_rmvpe_T6000_v1.jsonl 2> gpurun_out/trace.err; cut -c1-900 gpurun_out/trace_rmvpe_T6000_v1.jsonl; tail -3 gpurun_out/trace.err
