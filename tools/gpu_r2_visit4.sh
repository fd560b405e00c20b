#!/bin/bash
# Round-2 visit 4 (2 GPUs): tests that need two devices, the sharded song-level driver over NCCL (bit-identical to the
# unsharded run), graphs/PDL tests.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_graphs_gpu.py tests/test_pipeline_gpu.py tests/test_parity_gpu.py "tests/test_tc_gpu.py::test_infer_fused_pairs_equals_two_launch_form" -m gpu -q -x --timeout 600 > gpurun_out/pytest_v4.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; tail -6 gpurun_out/pytest_v4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sweep.py --what song --reps 3 --check > gpurun_out/song_2gpu.jsonl 2> gpurun_out/song2.err
echo "song2 rc=$?" | tee -a gpurun_out/status.txt; cut -c1-900 gpurun_out/song_2gpu.jsonl; tail -5 gpurun_out/song2.err
timeout 300 python tools/sweep.py --what song --reps 3 > gpurun_out/song_1gpu_v2.jsonl 2> gpurun_out/song1.err; cut -c1-900 gpurun_out/song_1gpu_v2.jsonl
