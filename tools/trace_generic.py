#!/usr/bin/env python
"""Where does a one-tile-per-CTA contraction spend its ~20 us?  Launches text-encoder / flow shaped contractions through
`rvcb200_op_conv_tc` with the kernel's phase recorder on (`rvcb200_debug_trace_conv_tc`) and prints, per shape, the median
over CTAs of the time between the recorded points (globaltimer, ns) plus the launch's CUDA-event time.

    python tools/trace_generic.py [--T 6000] [--reps 20]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comfy_rvc_b200 import _lib, weights  # noqa: E402

POINTS = ["start", "prologue", "producer_go", "slab0_landed", "mmas_issued", "acc_complete", "epilogue_done"]
SHAPES = [  # name, Cin, ntaps, Cout, gate, relu, res
    ("flow.in   K=5x192 N=384 gate", 192, 5, 384, 1, 0, 0),
    ("flow.res  K=192   N=192 +res", 192, 1, 192, 0, 0, 1),
    ("enc.qkv   K=192   N=768", 192, 1, 768, 0, 0, 0),
    ("enc.ffn1  K=3x192 N=768 relu", 192, 3, 768, 0, 1, 0),
    ("enc.ffn2  K=3x768 N=192 +res", 768, 3, 192, 0, 0, 1),
]
# RMVPE DeepUnet shapes (--rmvpe): name, image lines per 100 frames, W, Cin, Cout, N tile (rows = lines * (W + 1); 3 x 3 taps)
RMVPE_SHAPES = [
    ("unet L5 512->512 3x3 W=4", 100 / 32, 4, 512, 512, 64),
    ("unet L4 256->256 3x3 W=8", 100 / 16, 8, 256, 256, 64),
    ("unet L3 128->128 3x3 W=16", 100 / 8, 16, 128, 128, 64),
    ("unet L2 64->64 3x3 W=32", 100 / 4, 32, 64, 64, 64),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=6000)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--rmvpe", action="store_true", help="the DeepUnet's 3 x 3 image convolutions instead (T = frames)")
    ap.add_argument("--micro", action="store_true", help="who limits the weight stream: 1 CTA alone, 8 CTAs on the same weights, "
                    "8 CTAs on the same rows, 64 CTAs (K = 9 x 512, N tile 64)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    T = args.T
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    shapes = SHAPES
    if args.rmvpe:
        shapes = [(name, Cin, 9, Cout, 0, 0, 1, W, int(round(T / 100 * lines)) * (W + 1), nt)
                  for name, lines, W, Cin, Cout, nt in RMVPE_SHAPES]
    if args.micro:      # name, Cin, taps, Cout, gate, relu, res, W, rows, N tile
        args.rmvpe = True
        shapes = [("1 CTA (1 m x 1 n)", 512, 9, 64, 0, 0, 1, 4, 125, 64), ("8 CTAs, same weights (8 m x 1 n)", 512, 9, 64, 0, 0, 1, 4, 1000, 64),
                  ("8 CTAs, same rows (1 m x 8 n)", 512, 9, 512, 0, 0, 1, 4, 125, 64), ("64 CTAs (8 m x 8 n)", 512, 9, 512, 0, 0, 1, 4, 1000, 64),
                  ("1 CTA, N tile 128", 512, 9, 128, 0, 0, 1, 4, 125, 128), ("1 CTA, N tile 256", 512, 9, 256, 0, 0, 1, 4, 125, 256),
                  ("1 CTA, 1 tap x 4608 ch (no slab reuse)", 4608, 1, 64, 0, 0, 1, 4, 125, 64)]
    frames = T
    for shape in shapes:
        name, Cin, ntaps, Cout, gate, relu, res = shape[:7]
        Wimg, T, ntile = (shape[7], shape[8], shape[9]) if args.rmvpe else (0, frames, 64)
        x16 = torch.randn(1, T, Cin, device=dev).half()
        w = torch.randn(ntaps, Cin, Cout) / (Cin * ntaps) ** 0.5
        w16 = weights.pack_tc(w, torch.float16, ntile).to(dev)
        bias = torch.randn(Cout, device=dev)
        cout_eff = Cout // 2 if gate else Cout
        y32 = torch.zeros(1, T, cout_eff, device=dev)
        y16 = torch.zeros(1, T, cout_eff, device=dev, dtype=torch.float16)
        r32 = torch.randn(1, T, cout_eff, device=dev)
        d = _lib.TcConvDesc()
        d.x16, d.L_in, d.padf = x16.data_ptr(), T, 32
        d.w16, d.bias = w16.data_ptr(), bias.data_ptr()
        d.Cin, d.ntaps, d.dil, d.G = Cin, ntaps, 1, 1
        d.g_off[0] = -((ntaps - 1) // 2)
        d.N, d.Cout_total, d.Lj, d.out_stride, d.Lp_out = ntile, Cout, T, 1, ((T + 127) // 128) * 128 + 128
        if args.rmvpe and ntaps == 9:
            d.tap_w, d.dil2, d.g_off[0] = 3, Wimg + 1, -(Wimg + 2)
            d.pad_period, d.pad_valid, d.mask_post, d.pre_slope = Wimg + 1, Wimg, 1, 0.0
        d.div, d.out_slope, d.alpha, d.pre_slope = 1.0, 1.0, 1.0, 1.0
        d.generic, d.f32_cl, d.gate, d.relu = 1, 1, gate, relu
        d.y32, d.ldy32, d.y16 = y32.data_ptr(), cout_eff, y16.data_ptr()
        if res:
            d.res32, d.ldr32, d.res_mode = r32.data_ptr(), cout_eff, 1
        n_cta = min(((T + 127) // 128) * (Cout // ntile), 148)
        trace = torch.zeros(148 * 16 + 3 * 512, dtype=torch.int64, device=dev)
        for _ in range(3):
            assert lib.rvcb200_op_conv_tc(C.byref(d), 1, st) == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            lib.rvcb200_op_conv_tc(C.byref(d), 1, st)
        e1.record()
        torch.cuda.synchronize()
        us_plain = e0.elapsed_time(e1) / args.reps * 1e3
        assert lib.rvcb200_debug_trace_conv_tc(C.c_void_p(trace.data_ptr())) == 0
        segs = []
        for _ in range(args.reps):
            trace.zero_()
            torch.cuda.synchronize()
            lib.rvcb200_op_conv_tc(C.byref(d), 1, st)
            torch.cuda.synchronize()
            full = trace.cpu().numpy()
            steps = full[148 * 16:].reshape(3, 512).astype(np.float64)
            t = full[:148 * 16].reshape(148, 16)[:n_cta].astype(np.float64)
            t0 = t[:, 0].min()
            segs.append(np.concatenate([[np.median(t[:, 0] - t0)], np.median(np.diff(t[:, :7], axis=1), axis=0),
                                        [t[:, 6].max() - t0]]))
        lib.rvcb200_debug_trace_conv_tc(None)
        m = np.median(np.array(segs), axis=0) / 1e3
        line = {"shape": name, "T": T, "ctas": n_cta, "event_us_back_to_back": round(us_plain, 2),
                "cta_start_skew_us": round(m[0], 2)}
        for i in range(6):
            line[f"{POINTS[i]}->{POINTS[i + 1]}_us"] = round(m[1 + i], 2)
        line["first_start->last_epilogue_done_us"] = round(m[7], 2)
        nst = int((steps[0] > 0).sum())                                      # CTA 0, first tile, last repetition
        if nst >= 12:
            iss, rdy, mma = steps[0, :nst], steps[1, :nst], steps[2, :nst]
            line["cta0_steps"] = nst
            line["cta0_issue_interval_first10_us"] = round(float(np.diff(iss[:10]).mean()) / 1e3, 3)
            line["cta0_issue_interval_steady_us"] = round(float(np.diff(iss[10:]).mean()) / 1e3, 3)
            line["cta0_issue->full_first10_us"] = [round(float(v) / 1e3, 2) for v in (rdy - iss)[:10]]
            line["cta0_issue->full_steady_us"] = round(float(np.median((rdy - iss)[10:])) / 1e3, 3)
            line["cta0_full->mma_issued_us"] = round(float(np.median(mma - rdy)) / 1e3, 3)
            line["cta0_mma_issued->slot_reissued_us"] = round(float(np.median(iss[10:] - mma[:nst - 10])) / 1e3, 3)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
