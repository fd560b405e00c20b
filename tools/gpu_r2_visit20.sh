#!/bin/bash
# Round-2 visit 20 (1 GPU): TMA-staged 16-bit output in the generic epilogue: whole GPU suite, then A/B (RVCB200_GENERIC_TMA=0 = direct stores)
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_v20.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; tail -12 gpurun_out/pytest_v20.log
for g in 0 1; do
  RVCB200_GENERIC_TMA=$g timeout 300 python tools/bench_rmvpe.py --seconds 5,60 --no-incumbent > gpurun_out/rmvpe_bench_gt$g.jsonl 2>> gpurun_out/rmvpe_bench.err; cat gpurun_out/rmvpe_bench_gt$g.jsonl
  RVCB200_GENERIC_TMA=$g timeout 300 python tools/bench_hubert.py --seconds 5,60 > gpurun_out/hubert_bench_gt$g.jsonl 2>> gpurun_out/hubert_bench.err; cut -c1-200 gpurun_out/hubert_bench_gt$g.jsonl
  RVCB200_GENERIC_TMA=$g timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-incumbent --no-front-end > gpurun_out/bench_gt$g.json 2> gpurun_out/bench_gt$g.err; echo "bench rc=$?"
  python - <<P
import json
d = json.load(open("gpurun_out/bench_gt$g.json"))
print("GENERIC_TMA=$g", round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], {k: round(v,3) for k,v in d["time_by_class_ms_per_step"].items()}, d["parity"]["snr_db"], d["fp16"]["parity"]["snr_db"])
P
done
timeout 200 python tools/trace_generic.py --T 6000 > gpurun_out/trace_generic_T6000_v5_tma16.jsonl 2> gpurun_out/trace.err
python - <<'P'
import json
for l in open("gpurun_out/trace_generic_T6000_v5_tma16.jsonl"):
    d = json.loads(l); print(d["shape"], d["event_us_back_to_back"], "mma", d["slab0_landed->mmas_issued_us"], "epi", d["acc_complete->epilogue_done_us"])
P
