#!/bin/bash
# Round-2 visit 3 (1 GPU): graph / pipeline tests, launch list of a SHORT step (T = 100) to see where its 2 ms go,
# programmatic dependent launch A/B on short segments, song-level driver timing split on one GPU.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_graphs_gpu.py tests/test_pipeline_gpu.py tests/test_tc_gpu.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_v3.log 2>&1
echo "pytest rc=$?" | tee gpurun_out/status.txt; tail -8 gpurun_out/pytest_v3.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/step_T100.csv \
    python tools/ncu_step.py --precision bf16 --seconds 1 --out gpurun_out/short > gpurun_out/ncu_short.log 2>&1; tail -1 gpurun_out/ncu_short.log
python - <<'P'
import csv, re, collections
rows = list(csv.DictReader([l for l in open("gpurun_out/step_T100.csv") if not l.startswith("==")]))
agg = collections.OrderedDict(); tot = 0
for r in rows:
    n = re.sub(r"\(.*", "", r["Kernel Name"]).split("::")[-1]
    v = float(r["Metric Value"].replace(",", "")) / 1e3
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print("T=100 step: kernels", len(rows), "sum of kernel durations us", round(tot, 1))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]: print(f"{v[1]:8.1f} us n={v[0]:3d} avg {v[1]/v[0]:6.1f}  {k}")
P
for pdl in 0 1; do
RVCB200_PDL=$pdl timeout 300 python tools/sweep.py --what sweep --reps 10 --max-frames 1000 > gpurun_out/sweep_pdl$pdl.jsonl 2>> gpurun_out/sweep.err
grep '"batch": 1,' gpurun_out/sweep_pdl$pdl.jsonl | cut -c1-150
done
timeout 600 python tools/sweep.py --what song --reps 3 > gpurun_out/song_1gpu.jsonl 2> gpurun_out/song.err; cat gpurun_out/song_1gpu.jsonl | cut -c1-1200; tail -3 gpurun_out/song.err
