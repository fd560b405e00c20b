#!/bin/bash
# Round-2 visit 5 (1 GPU): HuBERT front end tests, graph/PDL/two-device fixes, quick bench.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_hubert_gpu.py -m gpu -q -x -s --timeout 300 > gpurun_out/pytest_hubert.log 2>&1
echo "hubert rc=$?" | tee gpurun_out/status.txt; grep -E "SNR|passed|failed|Error|error" gpurun_out/pytest_hubert.log | head -30; tail -25 gpurun_out/pytest_hubert.log
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_hubert_gpu.py > gpurun_out/pytest_v5.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/status.txt; tail -8 gpurun_out/pytest_v5.log
