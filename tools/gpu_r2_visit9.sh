#!/bin/bash
# Round-2 visit 9 (1 GPU): rotated K loop of the generic contractions (A/B), rate= branch, full tests.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_v9.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; tail -12 gpurun_out/pytest_v9.log
run() {
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-incumbent --no-front-end --no-extra-precision > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python - <<P
import json
d = json.load(open("gpurun_out/bench_$tag.json"))
c = d["time_by_class_ms_per_step"]
print("$tag", round(d["value"], 1), "ms", round(d["ms_per_step"], 3), "snr", round(d["parity"]["snr_db"], 2), {k: round(v, 3) for k, v in c.items()}, d["clocks"]["sm_mhz"], d["gpu_launches"] // 10)
P
}
run krot1 RVCB200_KROT=1
run krot0 RVCB200_KROT=0
run krot1b RVCB200_KROT=1
run krot0b RVCB200_KROT=0
timeout 300 python tools/bench_hubert.py --seconds 5,60 > gpurun_out/hubert_krot1.jsonl 2> gpurun_out/hubert_bench.err; cut -c1-200 gpurun_out/hubert_krot1.jsonl
RVCB200_KROT=0 timeout 300 python tools/bench_hubert.py --seconds 5,60 > gpurun_out/hubert_krot0.jsonl 2>> gpurun_out/hubert_bench.err; cut -c1-200 gpurun_out/hubert_krot0.jsonl
timeout 300 python tools/sweep.py --what song --reps 3 --tiers 60 --front-end b200 > gpurun_out/song_1gpu_b200fe.jsonl 2> gpurun_out/song.err; cut -c1-900 gpurun_out/song_1gpu_b200fe.jsonl; tail -3 gpurun_out/song.err
