#!/usr/bin/env python
"""One HuBERT / ContentVec call on `--seconds` of audio after a warm-up call, for `ncu --metrics gpu__time_duration.sum` launch lists."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comfy_rvc_b200 import synthetic  # noqa: E402
from comfy_rvc_b200.hubert import HubertB200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=60.0)
args = ap.parse_args()
m = HubertB200(synthetic.HUBERT_BASE, synthetic.make_hubert_state_dict(0), "cuda:0")
src = synthetic.make_speech(args.seconds, seed=1).cuda()
m.extract_features(version="v2", source=src)
torch.cuda.synchronize()
f = m.extract_features(version="v2", source=src)
torch.cuda.synchronize()
print("launches", m.last_launches, "frames", f.shape[1])
