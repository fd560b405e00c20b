#!/bin/bash
# Round-2 multi-GPU visit (8 GPUs of one box): BASELINE configs[3] (10 min song, segments sharded over 2/4/8 GPUs, checked
# bit for bit against the unsharded run) and configs[4] (length x batch sweep on 8 GPUs, outputs checked against the fp32 path).
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | wc -l
for n in 8 4 2; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      tools/sweep.py --what song --reps 3 --check > gpurun_out/song_${n}gpu.jsonl 2> gpurun_out/song_${n}gpu.err
  echo "song N=$n rc=$?" | tee -a gpurun_out/status_multi.txt
  python - <<P
import json
for l in open("gpurun_out/song_${n}gpu.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print({k: d[k] for k in ("tier", "n_gpus", "segments", "makespan_bound", "wall_s", "audio_s_per_s", "device_ms_max_over_ranks", "device_audio_s_per_s", "host_s_rank0", "sharded_equals_unsharded")})
P
done
timeout 300 python tools/sweep.py --what song --reps 3 > gpurun_out/song_1gpu.jsonl 2> gpurun_out/song_1gpu.err
python - <<'P'
import json
for l in open("gpurun_out/song_1gpu.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print({k: d[k] for k in ("tier", "n_gpus", "segments", "wall_s", "audio_s_per_s", "device_ms_max_over_ranks", "device_audio_s_per_s", "host_s_rank0")})
P
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 \
    tools/sweep.py --what sweep --reps 3 --check > gpurun_out/sweep_8gpu.jsonl 2> gpurun_out/sweep_8gpu.err
echo "sweep N=8 rc=$?" | tee -a gpurun_out/status_multi.txt
grep -c '"ms"' gpurun_out/sweep_8gpu.jsonl; python - <<'P'
import json
rows = [json.loads(l) for l in open("gpurun_out/sweep_8gpu.jsonl") if l.startswith("{") and '"ms"' in l]
print("min snr", min(r["snr_db_item0_vs_fp32_path"] for r in rows), "max rate", max(r["audio_s_per_s"] for r in rows))
for r in rows:
    if r["batch"] in (1, 256) or r["seconds"] == 30: print(r["config"], r["seconds"], r["batch"], r["ms"], r["audio_s_per_s"], r["snr_db_item0_vs_fp32_path"])
P
tail -3 gpurun_out/sweep_8gpu.err
