#!/usr/bin/env python
"""One infer() of the bench workload inside a cudaProfilerStart/Stop bracket (for `ncu --profile-from-start off`).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --profile-from-start off --csv --log-file gpurun_out/step.csv python tools/ncu_step.py --precision bf16

Before the bracket it runs warm-ups and one step under the library's own per-launch recorder and writes
gpurun_out/step_classes_<precision>.json: class and CUDA-event time of every launch in launch order, so that
tools/ncu_traffic.py can join ncu's launch list with the bench's kernel classes.  The noise is pre-drawn so that the
bracket holds nothing but this repo's kernels.
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import comfy_rvc_b200 as rvc  # noqa: E402
from comfy_rvc_b200 import _lib, synthetic  # noqa: E402
from comfy_rvc_b200.config import NAMED_CONFIGS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--config", default="48k_v2")
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out"))
    args = ap.parse_args()
    cfg = NAMED_CONFIGS[args.config]
    T = int(round(args.seconds * 100))
    sd = synthetic.make_state_dict(cfg)
    cls = rvc.SynthesizerTrnMs256NSFsid if cfg.feat_dim == 256 else rvc.SynthesizerTrnMs768NSFsid
    net = cls(*cfg.to_positional(), is_half=args.precision != "fp32")
    del net.enc_q
    net.load_state_dict({k: v.half() for k, v in sd.items()}, strict=False)
    net.eval().to("cuda:0").set_precision(args.precision)
    net.graph_max_frames = 0                  # eager launches: the per-launch recorder sits between them
    ins = [t.cuda() for t in synthetic.make_inputs(cfg, args.batch, T)]
    noise = net.draw_noise(args.batch, T)
    for _ in range(3):
        net.infer(*ins, noise=noise)
    torch.cuda.synchronize()
    lib = _lib.load()
    lib.rvcb200_profile_enable(net._ctx, 1)
    net.infer(*ins, noise=noise)
    torch.cuda.synchronize()
    cap = 4096
    cl, nk, ms = (C.c_int32 * cap)(), (C.c_int32 * cap)(), (C.c_float * cap)()
    n = int(lib.rvcb200_profile_launches(net._ctx, cl, nk, ms, cap))
    lib.rvcb200_profile_enable(net._ctx, 0)
    os.makedirs(args.out, exist_ok=True)
    json.dump({"precision": args.precision, "config": args.config, "T": T, "B": args.batch, "launches": n,
               "cls": [int(cl[i]) for i in range(n)], "kernels": [int(nk[i]) for i in range(n)],
               "event_ms": [float(ms[i]) for i in range(n)]},
              open(os.path.join(args.out, f"step_classes_{args.precision}.json"), "w"))
    torch.cuda.cudart().cudaProfilerStart()
    net.infer(*ins, noise=noise)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print(f"captured one step: {net.last_launches} launches ({n} recorded)")


if __name__ == "__main__":
    main()
