#!/bin/bash
# quick visit: chosen tests (-k "$1"), bench lines for each "ENV=VAL" variant in $2.. (plus the default)
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
sel="$1"; shift
timeout 900 python -m pytest tests/test_tc_gpu.py -x -q -s --timeout 200 --timeout-method=thread -k "$sel" > gpurun_out/test_quick.log 2>&1
echo "tests rc=$?" | tee gpurun_out/status.txt; grep -h "SNR\|fused vs" gpurun_out/test_quick.log | tail -30; tail -3 gpurun_out/test_quick.log
for v in "X=0" "$@"; do
  env $v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err
  python - "$v" <<'P'
import json, sys
d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
print(sys.argv[1], round(d["value"]), "RT  e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), {k: round(x, 3) for k, x in d["time_by_class_ms_per_step"].items()}, d["clocks"]["sm_mhz"])
P
done
