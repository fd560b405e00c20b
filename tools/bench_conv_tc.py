#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 convolution at the decoder's real shapes (48k_v2, 60 s segment).

    python tools/bench_conv_tc.py [--reps 5] [--only STAGE] [--profile]

For each stage (C, L) x kernel size x epilogue type (c1: 16-bit store only; c2: + fp32 residual read,
fp32 store) reports device time (CUDA events), TFLOP/s and algorithmic HBM GB/s.  With --profile the
timed launches are bracketed by cudaProfilerStart/Stop for `ncu --profile-from-start off`.
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comfy_rvc_b200 import _lib, weights  # noqa: E402

PADF = 32


def pitch(L):
    return ((L + 127) // 128) * 128 + 128


def bench_pairs(args, lib):
    """One ResBlock1 pair x' = x + c2(lrelu(c1(lrelu(x)))) at the decoder's stage-3/4 shapes: one fused launch vs two."""
    dev = torch.device("cuda", 0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rows = []
    for si, (Cc, L) in ((3, (64, args.T * 240)), (4, (32, args.T * 480))):
        if args.stages and (si - 1) not in {int(v) for v in args.stages.split(",")}:
            continue
        Lp = pitch(L)
        x16 = torch.randn(1, L, Cc, dtype=torch.float16, device=dev)
        h16 = torch.zeros(1, L, Cc, dtype=torch.float16, device=dev)
        y16 = torch.zeros(1, L, Cc, dtype=torch.float16, device=dev)
        xs = torch.zeros(1, Cc // 8, Lp, 8, dtype=torch.float16, device=dev)
        bias = torch.randn(Cc, device=dev)
        keep = {int(v) for v in args.ks.split(",")} if args.ks else None
        for k, dil in ((3, 1), (3, 5), (7, 1), (7, 5), (11, 1), (11, 5)):
            if (keep is not None and k not in keep) or (k == 11 and Cc != 32):     # C = 64, k = 11 is not a fused shape
                continue
            w = weights.pack_tc(torch.randn(1, k, Cc, Cc) / (Cc * k) ** 0.5, torch.float16).to(dev)
            for kind in ("s", "a"):
                d1, d2 = _lib.TcConvDesc(), _lib.TcConvDesc()
                for d, dl in ((d1, dil), (d2, 1)):
                    d.L_in, d.padf, d.Cin, d.ntaps, d.dil, d.G = L, PADF, Cc, k, dl, 1
                    d.g_off[0] = -((k - 1) // 2) * dl
                    d.N, d.Cout_total, d.Lj, d.out_stride, d.Lp_out = Cc, Cc, L, 1, Lp
                    d.div, d.out_slope, d.w16, d.bias = 1.0, 0.1, w.data_ptr(), bias.data_ptr()
                d1.x16, d1.y16 = x16.data_ptr(), h16.data_ptr()
                d2.x16, d2.res16, d2.res_neg_scale = h16.data_ptr(), x16.data_ptr(), 10.0
                if kind == "s":
                    d2.y16 = y16.data_ptr()
                else:
                    d2.y32, d2.acc_f16, d2.accum = xs.data_ptr(), 1, 1

                def fused():
                    return lib.rvcb200_op_rbpair_tc(C.byref(d1), C.byref(d2), 1, st)

                def two():
                    return lib.rvcb200_op_rbconv_tc(C.byref(d1), 1, st) | lib.rvcb200_op_rbconv_tc(C.byref(d2), 1, st)

                res = {}
                for nm, fn in (("fused", fused), ("two", two)):
                    for _ in range(2):
                        assert fn() == 0
                    torch.cuda.synchronize()
                    if args.profile and nm == "fused":
                        torch.cuda.profiler.start()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(args.reps):
                        fn()
                    e1.record()
                    torch.cuda.synchronize()
                    if args.profile and nm == "fused":
                        torch.cuda.profiler.stop()
                    res[nm] = e0.elapsed_time(e1) * 1e3 / args.reps
                flops = 2 * 2.0 * L * Cc * Cc * k
                b_f = L * Cc * (4 if kind == "s" else 6)
                b_t = L * Cc * (10 if kind == "s" else 12)
                rows.append(dict(pair=1, stage=si, C=Cc, L=L, k=k, dil=dil, kind=kind, fused_us=round(res["fused"], 1),
                                 two_us=round(res["two"], 1), fused_tflops=round(flops / res["fused"] / 1e6, 1),
                                 fused_hbm_gbs=round(b_f / res["fused"] / 1e3, 1), two_hbm_gbs=round(b_t / res["two"] / 1e3, 1)))
                print(json.dumps(rows[-1]), flush=True)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--T", type=int, default=6000)
    ap.add_argument("--a-mode", type=int, default=0)
    ap.add_argument("--dbg-alt", type=int, default=0)
    ap.add_argument("--rb", type=int, default=1, help="1: specialised resblock kernel (rbconv_tc.cu), 0: generic conv_tc")
    ap.add_argument("--stages", default="", help="comma list of stage indices (0-3) to keep")
    ap.add_argument("--kinds", default="c1,c2s,c2a", help="c1: 16-bit store only; c2: + fp32 planar residual in/out; "
                    "c2s: + residual from the fp16 lrelu-domain stream; c2a: stream residual, planar fp16 branch-sum accumulate, no 16-bit store")
    ap.add_argument("--ks", default="", help="comma list of kernel sizes to keep (default all)")
    ap.add_argument("--pair", action="store_true", help="fused ResBlock pair (rbpair_tc.cu) against its two-launch form")
    args = ap.parse_args()
    lib = _lib.load()
    if args.pair:
        return bench_pairs(args, lib)
    conv_fn = lib.rvcb200_op_rbconv_tc if args.rb else lib.rvcb200_op_conv_tc
    dev = torch.device("cuda", 0)
    stages = [(256, args.T * 12), (128, args.T * 120), (64, args.T * 240), (32, args.T * 480)]
    rows = []
    for si, (Cc, L) in enumerate(stages):
        if args.only >= 0 and si != args.only:
            continue
        if args.stages and si not in {int(v) for v in args.stages.split(",")}:
            continue
        Lp = pitch(L)
        x16 = torch.randn(1, L, Cc, dtype=torch.float16, device=dev)
        r32 = torch.randn(1, Cc // 4, Lp, 4, device=dev)
        y32 = torch.zeros(1, Cc // 4, Lp, 4, device=dev)
        y16 = torch.zeros(1, L, Cc, dtype=torch.float16, device=dev)
        r16 = torch.randn(1, L, Cc, dtype=torch.float16, device=dev)
        bias = torch.randn(Cc, device=dev)
        keep = {int(v) for v in args.ks.split(",")} if args.ks else None
        for k, dil in ((3, 1), (7, 3), (11, 5), (11, 1)):
            if keep is not None and (k not in keep or (k == 11 and dil == 1)):
                continue
            w = weights.pack_tc(torch.randn(1, k, Cc, Cc) / (Cc * k) ** 0.5, torch.float16).to(dev)
            for kind in args.kinds.split(","):
                d = _lib.TcConvDesc()
                d.x16, d.L_in, d.padf = x16.data_ptr(), L, PADF
                d.w16, d.bias = w.data_ptr(), bias.data_ptr()
                d.Cin, d.ntaps, d.dil, d.G = Cc, k, dil, 1
                d.a_mode = args.a_mode
                d.dbg_alt = args.dbg_alt
                d.g_off[0] = -((k - 1) // 2) * dil
                d.N, d.Cout_total = min(256, Cc), Cc
                d.Lj, d.out_stride, d.Lp_out = L, 1, Lp
                d.y16, d.out_slope, d.div = y16.data_ptr(), 0.1, 1.0
                if kind == "c2":
                    d.y32, d.res32 = y32.data_ptr(), r32.data_ptr()
                elif kind == "c2s":      # residual recovered from the fp16 lrelu-domain stream, one 16-bit store
                    d.res16, d.res_neg_scale = r16.data_ptr(), 10.0
                elif kind == "c2a":      # last pair of a resblock: stream residual + fp32 planar branch accumulate
                    d.res16, d.res_neg_scale, d.y32, d.accum, d.y16, d.acc_f16 = r16.data_ptr(), 10.0, y32.data_ptr(), 1, None, 1
                st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
                for _ in range(2):
                    assert conv_fn(C.byref(d), 1, st) == 0
                torch.cuda.synchronize()
                if args.profile:
                    torch.cuda.profiler.start()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.reps):
                    conv_fn(C.byref(d), 1, st)
                e1.record()
                torch.cuda.synchronize()
                if args.profile:
                    torch.cuda.profiler.stop()
                us = e0.elapsed_time(e1) * 1e3 / args.reps
                flops = 2.0 * L * Cc * Cc * k
                bytes_ = L * Cc * (2 + 2 + {"c1": 0, "c2": 8, "c2s": 2, "c2a": 4}[kind])
                rows.append(dict(rb=args.rb, a_mode=args.a_mode, stage=si + 1, C=Cc, L=L, k=k, dil=dil, kind=kind, us=round(us, 1),
                                 tflops=round(flops / us / 1e6, 1), hbm_gbs=round(bytes_ / us / 1e3, 1)))
                print(json.dumps(rows[-1]), flush=True)
    return rows


if __name__ == "__main__":
    main()
