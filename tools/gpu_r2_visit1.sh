#!/bin/bash
# Round-2 visit 1: full GPU suite incl. the BASELINE-size parity tests, smoke, the reworked bench line (parity block, fp16 leg,
# GPU incumbent, full-step CPU legs), the reference arm on the full step, and the ncu DRAM-traffic launch list of one step.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L; nproc
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 --timeout-method=thread -x --durations=15 > gpurun_out/pytest_gpu_full.log 2>&1
echo "pytest -m gpu rc=$?" | tee gpurun_out/status.txt; tail -30 gpurun_out/pytest_gpu_full.log
timeout 300 python -m pytest tests/test_baseline_size_gpu.py -m gpu -q -s --timeout 600 > gpurun_out/pytest_baseline_size.log 2>&1
echo "baseline-size rc=$?" | tee -a gpurun_out/status.txt; grep -E "SNR|LSB|passed|failed" gpurun_out/pytest_baseline_size.log | tail -40
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/status.txt; tail -4 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench rc=$?" | tee -a gpurun_out/status.txt
tail -5 gpurun_out/bench_bf16.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?" | tee -a gpurun_out/status.txt
python - <<'P'
import json
for n in ("bench_bf16", "bench_reference"):
    try:
        d = json.load(open(f"gpurun_out/{n}.json"))
        for k in ("value", "ms_per_step", "e2e", "parity", "fp16", "gpu_incumbent", "cpu_baseline", "time_by_class_ms_per_step", "clocks", "oracle_vs_reference_fixture"):
            if k in d: print(n, k, d[k])
        if "roofline" in d: print({k: d["roofline"][k] for k in ("achieved", "frac", "frac_of_burst", "traffic", "hbm_frac", "avg_launch_ms")})
    except Exception as e:
        print(n, "failed", e)
P
for prec in bf16; do
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --profile-from-start off --csv --log-file gpurun_out/step_$prec.csv python tools/ncu_step.py --precision $prec > gpurun_out/ncu_step_$prec.log 2>&1
tail -2 gpurun_out/ncu_step_$prec.log
done
ls -la gpurun_out | head -40
