#!/usr/bin/env python
"""One RMVPE call on `--seconds` of audio after a warm-up call, for `ncu --metrics gpu__time_duration.sum` launch lists
(tools/gpu_r2_visit12.sh); prints the launch count so the list can be cut to the last call."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comfy_rvc_b200 import synthetic  # noqa: E402
from comfy_rvc_b200.rmvpe import RMVPE  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=60.0)
args = ap.parse_args()
m = RMVPE(synthetic.make_rmvpe_state_dict(0), is_half=True, device="cuda:0")
audio = synthetic.make_speech(args.seconds, seed=1)[0].numpy()
m.infer_from_audio(audio)
torch.cuda.synchronize()
f0 = m.infer_from_audio(audio)
torch.cuda.synchronize()
print("launches", m.last_launches, "frames", f0.shape[0])
