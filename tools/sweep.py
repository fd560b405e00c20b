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
    ap.add_argument("--front-end", default="fake", choices=["fake", "b200"],
                    help="song: `fake` = seeded random features drawn on the device (isolates the synthesis path), `b200` = the "
                         "HuBERT / ContentVec front end on the same kernels (comfy_rvc_b200.HubertB200, seeded weights)")
    ap.add_argument("--f0", default="synthetic", choices=["synthetic", "rmvpe"],
                    help="song: `synthetic` = a seeded contour (isolates the synthesis path), `rmvpe` = the RMVPE f0 estimator on the "
                         "same kernels (comfy_rvc_b200.RMVPE, seeded weights), as the nodes' default f0_method does")
    ap.add_argument("--tiers", default="", help="song: comma list of x_center values to keep (60,38,30); default all")
    ap.add_argument("--check", action="store_true", help="song under torchrun: also run unsharded and compare bit for bit")
    ap.add_argument("--max-frames", type=int, default=0, help="sweep: only points with batch * frames <= this (0 = all)")
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
        # configs[3]: one 10 min song through VC.pipeline, the reference's own segments sharded over the ranks.  The three
        # x_pad/x_query/x_center/x_max tiers are all reference behaviour (config.py:124-141); they change the segment count.
        cfg = NAMED_CONFIGS["48k_v2"]
        net = build(cfg, args.precision, dev)
        audio = synthetic.make_song(args.song_seconds, seed=0)
        if args.front_end == "b200":
            from comfy_rvc_b200.hubert import HubertB200
            hubert = HubertB200(synthetic.HUBERT_BASE, synthetic.make_hubert_state_dict(0), dev)
        else:
            hubert = synthetic.FakeHubert(cfg.feat_dim, device_rng=True)
        rmvpe_model = None
        if args.f0 == "rmvpe":
            from comfy_rvc_b200.rmvpe import RMVPE
            rmvpe_model = RMVPE(synthetic.make_rmvpe_state_dict(0), is_half=True, device=dev)
        solo = [dist.new_group([r]) for r in range(world)] if world > 1 else None       # 1-rank groups: the unsharded run
        for tier in [t for t in ((3, 10, 60, 64), (1, 6, 38, 41), (1, 5, 30, 32)) if not args.tiers or str(t[2]) in args.tiers.split(",")]:
            def make_vc(group=None):
                v = pl.VC(cfg.sr, pl.PipelineConfig(*tier, is_half=False, device=str(dev)), noise="device", group=group)
                v.f0_method_dict["synthetic"] = synthetic.pipeline_f0
                if rmvpe_model is not None:
                    v.model_rmvpe = rmvpe_model
                return v

            def run(v):
                return v.pipeline(hubert, net, 0, audio.copy(), [0, 0, 0], 0, args.f0, "median", "", 0.0, 1, 3, cfg.sr, 0, 1.0,
                                  "v2", 0.5, 160, False, False, None, 50, 1100)
            vc = make_vc()
            walls, dev_ms, host = [], [], []
            for it in range(1 + args.reps):
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                out = run(vc)
                torch.cuda.synchronize(dev)
                if world > 1:
                    dist.barrier()
                if it:
                    walls.append(time.perf_counter() - t0)
                    dev_ms.append(vc.last_plan["device_ms"])
                    host.append(vc.last_plan["host_s"])
            same = None
            if world > 1 and args.check:
                ref = run(make_vc(solo[rank]))                                           # every rank: the whole song on its own GPU
                same = bool(np.array_equal(ref, out)) if rank == 0 else None
            dm = torch.tensor([float(np.median(dev_ms))], device=dev)
            if world > 1:
                dist.all_reduce(dm, op=dist.ReduceOp.MAX)
            if rank == 0:
                plan = vc.last_plan
                secs = out.shape[0] / cfg.sr
                w = float(np.median(walls))
                h = {k: round(float(np.median([x[k] for x in host])), 4) for k in host[0]}
                print(json.dumps({"config": f"48k_v2 VC.pipeline, {args.song_seconds:.0f} s song", "front_end": args.front_end, "f0": args.f0,
                                  "tier": list(tier), "n_gpus": world,
                                  "precision": args.precision, "segments": len(plan["segments"]),
                                  "segment_seconds": [round(s.n_samples / 16000, 1) for s in plan["segments"]],
                                  "assignment": plan["assignment"], "makespan_bound": round(plan["makespan_bound"], 3),
                                  "wall_s": round(w, 4), "audio_s_per_s": round(secs / w, 1),
                                  "device_ms_max_over_ranks": round(float(dm), 2),
                                  "device_audio_s_per_s": round(secs / (float(dm) / 1e3), 1), "host_s_rank0": h,
                                  "sharded_equals_unsharded": same,
                                  "note": "wall = whole call on rank 0 (C filtfilt, H2D, device quiet-point search, segments, device "
                                          "gather, D2H); device_ms = CUDA events from the first segment to the end of the gather"}),
                      flush=True)
        del net
        torch.cuda.empty_cache()

    if "sweep" in what:
        # configs[4]: every rank runs the same (length, batch) point on its own GPU (weak scaling, no collective on the data
        # path); the time of a point is the max over ranks and the rate the aggregate of all ranks
        for cname in ("40k", "48k_v2"):
            cfg = NAMED_CONFIGS[cname]
            net = build(cfg, args.precision, dev)
            ref32 = build(cfg, "fp32", dev) if args.check else None
            for secs in (1, 2, 5, 10, 20, 30):
                for B in (1, 4, 16, 64, 256):
                    T = secs * 100
                    if B * T > 256 * 1000 or (args.max_frames and B * T > args.max_frames):   # workspace under ~80 GB
                        continue
                    try:
                        ms, rate = time_infer(net, cfg, B, T, args.reps, dev)
                        snr = None
                        if ref32 is not None:            # item 0 of the batch against the +-1 LSB fp32 path, same inputs and noise
                            ins = [t.to(dev) for t in synthetic.make_inputs(cfg, B, T, seed=3)]
                            noise = net.draw_noise(B, T)
                            o = net.infer(*ins, noise=noise)[0][0, 0]
                            r = ref32.infer(*[t[:1] for t in ins], noise=tuple(t[:1] for t in noise))[0][0, 0]
                            snr = synthetic.snr_db(r.cpu().numpy(), o.cpu().numpy())
                    except RuntimeError as e:     # out of memory on this box: report and go on
                        if rank == 0:
                            print(json.dumps({"config": cname, "seconds": secs, "batch": B, "error": str(e)[:80]}), flush=True)
                        torch.cuda.empty_cache()
                        ms = float("nan")
                    t = torch.tensor([ms], device=dev)
                    if world > 1:
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    if rank == 0 and ms == ms:
                        ms_all = float(t)
                        print(json.dumps({"config": cname, "precision": args.precision, "seconds": secs, "batch": B, "n_gpus": world,
                                          "ms": round(ms_all, 3), "audio_s_per_s": round(world * B * T * cfg.upp / cfg.sr / (ms_all / 1e3), 1),
                                          "graph_replay": bool(net.last_graph_replay),
                                          "snr_db_item0_vs_fp32_path": None if snr is None else round(snr, 1)}), flush=True)
            del net, ref32
            torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
