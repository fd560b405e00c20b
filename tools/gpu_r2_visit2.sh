#!/bin/bash
# Round-2 visit 2: new tests (CUDA graphs, k = 11 fused pairs, attention at T = 6000 / 8600, ResBlock2 on the tensor path),
# pair micro-benchmark of the new shape in both pipeline shapes, bench line, short-segment sweep with graph replay.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_graphs_gpu.py tests/test_tc_gpu.py tests/test_parity_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_v2.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; tail -15 gpurun_out/pytest_v2.log
timeout 300 python tools/bench_conv_tc.py --pair --reps 5 --stages 3 --ks 11 > gpurun_out/pairs_k11_cfg0.jsonl 2> gpurun_out/pairs.err; cat gpurun_out/pairs_k11_cfg0.jsonl
RVCB200_PAIR_CFG=1 timeout 300 python tools/bench_conv_tc.py --pair --reps 5 --stages 3 --ks 11 > gpurun_out/pairs_k11_cfg1.jsonl 2>> gpurun_out/pairs.err; cat gpurun_out/pairs_k11_cfg1.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > gpurun_out/bench_bf16_v2.json 2> gpurun_out/bench_bf16_v2.err; echo "bench rc=$?" | tee -a gpurun_out/status.txt
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_bf16_v2.json"))
for k in ("value", "ms_per_step", "parity", "time_by_class_ms_per_step", "clocks"): print(k, d.get(k))
print({k: d["roofline"][k] for k in ("achieved", "frac", "frac_of_burst", "traffic", "hbm_frac", "avg_launch_ms", "launches_per_step")})
P
timeout 900 python tools/sweep.py --what sweep --reps 5 > gpurun_out/sweep_graphs.jsonl 2> gpurun_out/sweep.err; tail -60 gpurun_out/sweep_graphs.jsonl
RVCB200_GRAPH_FRAMES=0 timeout 600 python tools/sweep.py --what sweep --reps 5 --max-frames 2500 > gpurun_out/sweep_nographs.jsonl 2>> gpurun_out/sweep.err; tail -30 gpurun_out/sweep_nographs.jsonl
