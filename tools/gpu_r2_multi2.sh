#!/bin/bash
# Round-2 multi-GPU visit 2 (8 GPUs of one box): BASELINE configs[3] with BOTH real front ends (HuBERT + RMVPE on the same kernels),
# segments sharded over 8 / 4 / 2 GPUs and checked bit for bit against the unsharded run; bench.py at N = 8.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | wc -l
for n in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
      tools/sweep.py --what song --front-end b200 --f0 rmvpe --tiers 60,38 --reps 3 --check > gpurun_out/song_real_${n}gpu.jsonl 2> gpurun_out/song_real_${n}gpu.err
  echo "song N=$n rc=$?" | tee -a gpurun_out/status_multi2.txt
  python - <<P
import json
for l in open("gpurun_out/song_real_${n}gpu.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print({k: d[k] for k in ("tier", "n_gpus", "segments", "makespan_bound", "wall_s", "audio_s_per_s", "device_ms_max_over_ranks", "device_audio_s_per_s", "host_s_rank0", "sharded_equals_unsharded")})
P
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu_v2.json 2> gpurun_out/bench_8gpu_v2.err
echo "bench N=8 rc=$?" | tee -a gpurun_out/status_multi2.txt
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_8gpu_v2.json"))
print({k: d.get(k) for k in ("value", "n_gpus", "ms_per_step", "e2e", "clocks", "scaling")})
P
