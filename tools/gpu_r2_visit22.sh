#!/bin/bash
# Round-2 visit 22 (1 GPU): HuBERT positional convolution as one grouped launch: HuBERT tests + timing, generic op tests
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_hubert_gpu.py tests/test_tc_gpu.py tests/test_rmvpe_gpu.py -q -s --timeout 300 > gpurun_out/pytest_hubert_v2.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; grep -E "passed|failed|Error|error|h[123]_" gpurun_out/pytest_hubert_v2.log | cut -c1-250 | tail -12
timeout 300 python tools/bench_hubert.py --seconds 5,20,60 > gpurun_out/hubert_bench_v6.jsonl 2>> gpurun_out/hubert_bench.err; cut -c1-330 gpurun_out/hubert_bench_v6.jsonl; tail -3 gpurun_out/hubert_bench.err
