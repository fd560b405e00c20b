#!/bin/bash
# Round-2 visit 24 (1 GPU): generic epilogue with its descriptor fields in registers: traces first (fast signal), whole suite, timings
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python tools/trace_generic.py --T 6000 > gpurun_out/trace_generic_T6000_v6_regs.jsonl 2> gpurun_out/trace.err
timeout 200 python tools/trace_generic.py --rmvpe --T 6000 >> gpurun_out/trace_generic_T6000_v6_regs.jsonl 2>> gpurun_out/trace.err
python - <<'P'
import json
for l in open("gpurun_out/trace_generic_T6000_v6_regs.jsonl"):
    d = json.loads(l); print(d["shape"], d["event_us_back_to_back"], "mma", d["slab0_landed->mmas_issued_us"], "epi", d["acc_complete->epilogue_done_us"])
P
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_v24.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; tail -4 gpurun_out/pytest_v24.log
timeout 300 python tools/bench_rmvpe.py --seconds 5,60 --no-incumbent > gpurun_out/rmvpe_bench_v8.jsonl 2>> gpurun_out/rmvpe_bench.err; cut -c1-120 gpurun_out/rmvpe_bench_v8.jsonl
timeout 300 python tools/bench_hubert.py --seconds 5,60 > gpurun_out/hubert_bench_v7.jsonl 2>> gpurun_out/hubert_bench.err; cut -c1-120 gpurun_out/hubert_bench_v7.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-incumbent --no-front-end > gpurun_out/bench_v24.json 2> gpurun_out/bench_v24.err; echo "bench rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_v24.json"))
print(round(d["ms_per_step"],3), d["clocks"]["sm_mhz"], {k: round(v,3) for k,v in d["time_by_class_ms_per_step"].items()}, d["parity"]["snr_db"], d["fp16"]["parity"]["snr_db"])
P
