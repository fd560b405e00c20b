#!/usr/bin/env python
"""Micro-benchmark of the transposed-conv ladder launches (48k_v2, 60 s segment) in their output modes:

    f32    fp32 planar-vector output (then a separate source-injection kernel reads it back)
    h16    16-bit stream output, no source injection (the no-f0 classes)
    fused  16-bit stream output + source injection in the epilogue (engine default for noise kernels <= 16 taps)

    python tools/bench_ups.py [--reps 5] [--stages 1,2,3] [--modes f32,h16,fused] [--profile]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comfy_rvc_b200 import _lib, weights  # noqa: E402
from comfy_rvc_b200.config import NAMED_CONFIGS  # noqa: E402

PADF = 32


def pitch(L):
    return ((L + 127) // 128) * 128 + 128


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--T", type=int, default=6000)
    ap.add_argument("--config", default="48k_v2")
    ap.add_argument("--stages", default="0,1,2,3")
    ap.add_argument("--modes", default="f32,h16,fused")
    ap.add_argument("--profile", action="store_true")
    args = ap.parse_args()
    cfg = NAMED_CONFIGS[args.config]
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    L_har = args.T * cfg.upp
    har = torch.randn(1, L_har, device=dev)
    Lc, Cc = args.T, cfg.upsample_initial_channel
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        Cn, Ln = Cc // 2, Lc * u
        if str(i) in args.stages.split(","):
            pad, ntaps, g_off = weights.up_geometry(k, u)
            w = weights.pack_conv_transpose(torch.randn(Cc, Cn, k) / (Cc * 2) ** 0.5, u)
            w16 = weights.pack_tc(w, torch.float16).to(dev)
            x16 = torch.randn(1, Lc, Cc, dtype=torch.float16, device=dev)
            bias = torch.randn(Cn, device=dev)
            nk, ns, npad = cfg.noise_conv_geometry(i)
            wn = torch.randn(nk, Cn, device=dev)
            nb = torch.randn(Cn, device=dev)
            y32 = torch.zeros(1, Cn // 4, pitch(Ln), 4, device=dev)
            y16 = torch.zeros(1, Ln, Cn, dtype=torch.float16, device=dev)
            for mode in args.modes.split(","):
                d = _lib.TcConvDesc()
                d.x16, d.L_in, d.padf = x16.data_ptr(), Lc, PADF
                d.w16, d.bias = w16.data_ptr(), bias.data_ptr()
                d.Cin, d.ntaps, d.dil, d.G = Cc, ntaps, 1, u
                for p_, o in enumerate(g_off):
                    d.g_off[p_] = o
                d.N, d.Cout_total = min(256, Cn), Cn
                d.Lj, d.out_stride, d.Lp_out = Lc, u, pitch(Ln)
                d.div, d.out_slope = 1.0, 0.1
                if mode == "f32":
                    d.y32 = y32.data_ptr()
                else:
                    d.y16 = y16.data_ptr()
                st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                for _ in range(2):
                    assert lib.rvcb200_op_conv_tc(C.byref(d), 1, st) == 0
                torch.cuda.synchronize()
                if args.profile:
                    torch.cuda.profiler.start()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.reps):
                    lib.rvcb200_op_conv_tc(C.byref(d), 1, st)
                e1.record()
                torch.cuda.synchronize()
                if args.profile:
                    torch.cuda.profiler.stop()
                us = e0.elapsed_time(e1) * 1e3 / args.reps
                flops = 2.0 * Ln * Cc * Cn * ntaps
                bytes_ = Lc * Cc * 2 + Ln * Cn * (4 if mode == "f32" else 2)
                print(json.dumps(dict(stage=i + 1, Cin=Cc, Cout=Cn, u=u, k=k, L_out=Ln, noise_k=nk, mode=mode, us=round(us, 1),
                                      tflops=round(flops / us / 1e6, 1), hbm_gbs=round(bytes_ / us / 1e3, 1))), flush=True)
        Lc, Cc = Ln, Cn


if __name__ == "__main__":
    main()
