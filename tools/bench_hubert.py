#!/usr/bin/env python
"""HuBERT / ContentVec front end: the B200 kernels (comfy_rvc_b200.HubertB200) against the incumbent the reference runs --
HuggingFace `transformers.HubertModel` (eager PyTorch, cuDNN / cuBLAS) on the same GPU, fp16 and fp32 (TF32 off) -- on
one utterance of `--seconds` at 16 kHz.  Prints one JSON line per length: ms per call, audio seconds per second, and the
SNR of the B200 features against the incumbent's fp32 features (same seeded weights and audio).

    python tools/bench_hubert.py [--seconds 10,60] [--reps 5] [--version v2]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comfy_rvc_b200 import synthetic  # noqa: E402
from comfy_rvc_b200.hubert import HubertB200  # noqa: E402


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", default="10,60")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--version", default="v2")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    sd = synthetic.make_hubert_state_dict(0)
    ours = HubertB200(synthetic.HUBERT_BASE, sd, dev)
    inc = {}
    try:                                                   # the incumbent: unmodified HuggingFace model, library code
        from transformers import HubertConfig, HubertModel
        cfg = HubertConfig(**{k: (list(v) if isinstance(v, tuple) else v) for k, v in synthetic.HUBERT_BASE.items()})
        hf = HubertModel(cfg)
        hf.load_state_dict({k: v for k, v in sd.items() if not k.startswith("final_proj")})
        hf = hf.eval().to(dev)
        inc["hf"] = hf
    except Exception as e:  # noqa: BLE001
        inc["error"] = f"{type(e).__name__}: {e}"[:200]
    layer = (9 if args.version == "v1" else 12) - 1
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    for secs in [float(s) for s in args.seconds.split(",")]:
        src = synthetic.make_speech(secs, seed=1).to(dev)
        ms, f = timed(lambda: ours.extract_features(version=args.version, source=src), args.reps)
        line = {"seconds": secs, "version": args.version, "frames": int(f.shape[1]), "b200_ms": round(ms, 3),
                "b200_audio_s_per_s": round(secs / (ms / 1e3), 1), "launches": ours.last_launches}
        if "hf" in inc:
            hf = inc["hf"]
            with torch.no_grad():
                ms32, ref = timed(lambda: hf.float()(src, output_hidden_states=True)["hidden_states"][layer], max(2, args.reps // 2))
                ref = ref.float().cpu().numpy()
                ms16, _ = timed(lambda: hf.half()(src.half(), output_hidden_states=True)["hidden_states"][layer], args.reps)
                hf.float()
            line.update(incumbent_fp32_ms=round(ms32, 3), incumbent_fp16_ms=round(ms16, 3),
                        speedup_vs_fp16=round(ms16 / ms, 2), speedup_vs_fp32=round(ms32 / ms, 2))
            if args.version == "v2":
                line["snr_db_vs_incumbent_fp32"] = round(synthetic.snr_db(ref, f.float().cpu().numpy()), 1)
        else:
            line["incumbent"] = inc["error"]
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
