#!/usr/bin/env python
"""RMVPE f0 estimator: the B200 kernels (comfy_rvc_b200.RMVPE.infer_from_audio: host audio in, host f0 out) against the incumbent
the reference runs on a GPU -- the same network in eager PyTorch (cuDNN convolutions, cuDNN GRU), fp32 (TF32 off) and fp16
(`is_half`), through `bench.py`'s `rmvpe_incumbent` leg -- on one utterance of `--seconds` at 16 kHz.
One JSON line per length: ms per call, audio seconds per second, launches, the GRU recurrence alone, parity against the incumbent's
fp32 salience.

    python tools/bench_rmvpe.py [--seconds 5,20,60] [--reps 5] [--no-incumbent]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comfy_rvc_b200 import _lib, synthetic  # noqa: E402
from comfy_rvc_b200.rmvpe import RMVPE  # noqa: E402


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def logit(p):
    p = np.clip(np.asarray(p, dtype=np.float64), 1e-7, 1 - 1e-7)
    return np.log(p) - np.log1p(-p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", default="5,20,60")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-incumbent", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    sd = synthetic.make_rmvpe_state_dict(0)
    ours = RMVPE(sd, is_half=True, device=dev)
    lib = _lib.load()
    inc = None
    if not args.no_incumbent:                                      # the incumbent leg lives in bench.py (oracle/ is test infrastructure)
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        import bench as _bench
        inc = lambda audio_np, dtype: _bench.rmvpe_incumbent(audio_np, dev, dtype == torch.float16, reps=1)[1]
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for secs in [float(s) for s in args.seconds.split(",")]:
        audio = synthetic.make_speech(secs, seed=1)[0].numpy()
        taps = {}
        ours.infer_from_audio(audio, taps=taps)
        hidden = taps["hidden"].cpu().numpy()
        ms, f0 = timed(lambda: ours.infer_from_audio(audio), args.reps)
        T = ours._padded_frames(audio.shape[0] // 160 + 1)
        gi = torch.randn(T, 1536, device=dev)
        o16 = torch.empty(T, 512, dtype=torch.float16, device=dev)
        W = ours._w
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        ms_gru, _ = timed(lambda: lib.rvcb200_op_rmvpe_gru(C.c_void_p(gi.data_ptr()), C.c_void_p(W["gru.hh.w"].data_ptr()),
                                                           C.c_void_p(W["gru.hh.b"].data_ptr()), C.c_void_p(o16.data_ptr()), None, T, st),
                          args.reps)
        line = {"seconds": secs, "frames": int(f0.shape[0]), "b200_ms": round(ms, 3), "b200_audio_s_per_s": round(secs / (ms / 1e3), 1),
                "launches": ours.last_launches, "gru_recurrence_ms": round(ms_gru, 3), "gru_us_per_step": round(ms_gru * 1e3 / T, 3)}
        if inc is not None:
            ms32, ref = timed(lambda: inc(audio, torch.float32), max(2, args.reps // 2))
            ms16, h16 = timed(lambda: inc(audio, torch.float16), args.reps)
            line.update(incumbent_fp32_ms=round(ms32, 3), incumbent_fp16_ms=round(ms16, 3), speedup_vs_fp16=round(ms16 / ms, 2),
                        speedup_vs_fp32=round(ms32 / ms, 2),
                        logit_snr_db_vs_incumbent_fp32=round(synthetic.snr_db(logit(ref), logit(hidden)), 1),
                        incumbent_fp16_logit_snr_db=round(synthetic.snr_db(logit(ref), logit(h16)), 1),
                        hidden_max_abs_err=float(np.abs(ref - hidden).max()))
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
