#!/bin/bash
# Round-2 final visit (1 GPU): full GPU suite, smoke, the default bench line with every leg, fp32 line, reference arm, launch list of a bench step.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_r2_final.log 2>&1
echo "pytest -m gpu rc=$?" | tee gpurun_out/status.txt; tail -3 gpurun_out/pytest_r2_final.log
timeout 400 python __graft_entry__.py smoke > gpurun_out/smoke_r2_final.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/status.txt; tail -5 gpurun_out/smoke_r2_final.log
timeout 1500 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo "bench rc=$?" | tee -a gpurun_out/status.txt
timeout 400 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-incumbent --no-front-end > gpurun_out/bench_r2_final_fp32.json 2> gpurun_out/bench_fp32.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_final_reference.json 2> gpurun_out/bench_reference.err
python - <<'P'
import json
for n in ("bench_r2_final", "bench_r2_final_fp32", "bench_r2_final_reference"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        print(n, round(d["value"], 1), "RT  e2e", round(d["e2e"]["value"], 1), "ms", round(d.get("ms_per_step", 0), 3), d.get("time_by_class_ms_per_step"), d.get("clocks"), d.get("parity", {}).get("snr_db"))
        if n == "bench_r2_final":
            print({k: d["roofline"][k] for k in ("achieved", "frac", "frac_of_burst", "traffic", "hbm_frac")}, d["front_end"], d["f0_front_end"], d["gpu_incumbent"], d["cpu_baseline"])
    except Exception as e:
        print(n, "failed", e)
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-incumbent --no-front-end --no-parity > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-200
du -sh gpurun_out
