#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations on one GPU (bench.py measures configs[1]):

    config3   32k_v2, 64 x 8 s segments in ONE batched infer() (bf16)                      (BASELINE configs[2])
    song      10 min synthetic 16 kHz song through VC.pipeline at 48k_v2 (fake HuBERT / synthetic f0): planning,
              H2D once, segments back to back, device-side trim / peak-normalise / int16, one D2H   (configs[3];
              under torchrun the segments are sharded over the ranks and gathered on rank 0)
    sweep     segment length {1,2,5,10,20,30} s x batch {1,4,16,64} at 40k v1 and 48k_v2    (configs[4], memory-capped)

    python tools/sweep.py [--what config3,song,sweep] [--precision bf16] [--reps 3]
    python -m torch.distributed.run --nproc-per-node N tools/sweep.py --what song

Every line is JSON: audio seconds produced per wall second (CUDA events, max over ranks for the song).
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import comfy_rvc_b200 as rvc  # noqa: E402
from comfy_rvc_b200 import pipeline as pl, synthetic  # noqa: E402
from comfy_rvc_b200.config import NAMED_CONFIGS  # noqa: E402


def build(cfg, precision, dev):
    sd = synthetic.make_state_dict(cfg)
    cls = rvc.SynthesizerTrnMs256NSFsid if cfg.feat_dim == 256 else rvc.SynthesizerTrnMs768NSFsid
    net = cls(*cfg.to_positional(), is_half=precision != "fp32")
    del net.enc_q
    net.load_state_dict({k: v.half() for k, v in sd.items()}, strict=False)
    return net.eval().to(dev).set_precision(precision)


def time_infer(net, cfg, B, T, reps, dev):
    ins = [t.to(dev) for t in synthetic.make_inputs(cfg, B, T, seed=3)]
    for _ in range(2):
        net.infer(*ins)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        net.infer(*ins)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    return ms, B * T * cfg.upp / cfg.sr / (ms / 1e3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="config3,song,sweep")
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--song-seconds", type=float, default=600.0)
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    what = args.what.split(",")

    if "config3" in what and rank == 0:
        cfg = NAMED_CONFIGS["32k_v2"]
        net = build(cfg, args.precision, dev)
        ms, rate = time_infer(net, cfg, 64, 800, args.reps, dev)
        print(json.dumps({"config": "32k_v2 batched 64 x 8 s", "precision": args.precision, "ms_per_batch": round(ms, 2),
                          "audio_s_per_s": round(rate, 1), "launches": net.last_launches}), flush=True)
        del net
        torch.cuda.empty_cache()

    if "song" in what:
        cfg = NAMED_CONFIGS["48k_v2"]
        net = build(cfg, args.precision, dev)
        audio = synthetic.make_song(args.song_seconds, seed=0)
        vc = pl.VC(cfg.sr, pl.PipelineConfig(3, 10, 60, 64, is_half=False, device=str(dev)), noise="device")   # the half-mode tier
        vc.f0_method_dict["synthetic"] = synthetic.pipeline_f0
        hubert = synthetic.FakeHubert(cfg.feat_dim)
        walls = []
        for it in range(1 + args.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            out = vc.pipeline(hubert, net, 0, audio.copy(), [0, 0, 0], 0, "synthetic", "median", "", 0.0, 1, 3, cfg.sr, 0, 1.0,
                              "v2", 0.5, 160, False, False, None, 50, 1100)
            torch.cuda.synchronize(dev)
            if world > 1:
                dist.barrier()
            if it:
                walls.append(time.perf_counter() - t0)
        if rank == 0:
            plan = vc.last_plan
            secs = out.shape[0] / cfg.sr
            w = float(np.median(walls))
            print(json.dumps({"config": f"48k_v2 VC.pipeline, {args.song_seconds:.0f} s song, tier (3,10,60,64)", "n_gpus": world,
                              "precision": args.precision, "segments": len(plan["segments"]),
                              "segment_seconds": [round(s.n_samples / 16000, 1) for s in plan["segments"]],
                              "makespan_bound": plan["makespan_bound"], "wall_s": round(w, 4),
                              "audio_s_per_s": round(secs / w, 1),
                              "note": "wall clock of the whole call: host planning + filtfilt + f0 post-processing, H2D, all "
                                      "segments, device-side finalise, D2H, gather"}), flush=True)
        del net
        torch.cuda.empty_cache()

    if "sweep" in what and rank == 0:
        for cname in ("40k", "48k_v2"):
            cfg = NAMED_CONFIGS[cname]
            net = build(cfg, args.precision, dev)
            for secs in (1, 2, 5, 10, 20, 30):
                for B in (1, 4, 16, 64):
                    T = secs * 100
                    if B * T > 64 * 3000:          # keep the workspace under ~60 GB
                        continue
                    try:
                        ms, rate = time_infer(net, cfg, B, T, args.reps, dev)
                    except RuntimeError as e:     # out of memory on this box: report and go on
                        print(json.dumps({"config": cname, "seconds": secs, "batch": B, "error": str(e)[:80]}), flush=True)
                        torch.cuda.empty_cache()
                        continue
                    print(json.dumps({"config": cname, "precision": args.precision, "seconds": secs, "batch": B,
                                      "ms": round(ms, 3), "audio_s_per_s": round(rate, 1)}), flush=True)
            del net
            torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
