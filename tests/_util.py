"""Shared helpers for the parity tests: golden fixtures + seeded inputs."""
from __future__ import annotations

import ast
import os

import numpy as np
import torch

from comfy_rvc_b200 import synthetic
from comfy_rvc_b200.config import NAMED_CONFIGS, resolve

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["c1_40k_v1", "c2_48k_v2", "c3_32k_v2_ragged", "c4_48k_v2_unvoiced", "c5_48k_v1_5stage", "c6_40k_v1_tiny",
                "c7_40k_v1_nono", "c8_48k_v2_nono_ragged", "c9_40k_v1_resblock2", "c10_48k_v2_resblock2x"]


def load_golden(name):
    """Returns (cfg, state_dict, inputs, noise, golden arrays) for one fixture minted by make_golden.py."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=True)
    cfg_name, B, T, lengths, f0v, wseed, iseed, nseed, _torch_ver = [str(x) for x in z["meta"]]
    cfg = resolve(cfg_name)                        # ":nono" = the no-f0 classes (models.py:812-1021), ":rb2" = ResBlock2
    B, T = int(B), int(T)
    lengths = ast.literal_eval(lengths)
    sd = synthetic.make_state_dict(cfg, seed=int(wseed))
    inputs = synthetic.make_inputs(cfg, B, T, seed=int(iseed), lengths=lengths, f0_variant=f0v)
    noise = synthetic.draw_noise(cfg, B, T, seed=int(nseed))
    return cfg, sd, inputs, noise, z


def oracle_infer(cfg, w, inputs, noise, taps=None):
    """The oracle call matching the synthesizer class: `infer` (f0) or `infer_nono` (models.py:905-915)."""
    from oracle import rvc_oracle
    phone, lens, pitch, pitchf, sid = inputs
    if cfg.f0:
        return rvc_oracle.infer(w, cfg, phone, lens, pitch, pitchf, sid, *noise, taps=taps)
    return rvc_oracle.infer_nono(w, cfg, phone, lens, sid, noise[0], taps=taps)


def net_infer(net, cfg, inputs, noise=None, taps=None, device="cuda"):
    """`net.infer` with the reference's positional signature for the class (5 tensors with f0, 3 without)."""
    phone, lens, pitch, pitchf, sid = [t.to(device) for t in inputs]
    kw = {}
    if noise is not None:
        kw["noise"] = noise if cfg.f0 else noise[:1]
    if taps is not None:
        kw["taps"] = taps
    if cfg.f0:
        return net.infer(phone, lens, pitch, pitchf, sid, **kw)
    return net.infer(phone, lens, sid, **kw)


def int16_lsb_diff(ref_f32: np.ndarray, est_f32: np.ndarray) -> int:
    """max |int16(ref) - int16(est)| with the reference conversion (vc_infer_pipeline.py:188-189)."""
    a = synthetic.to_int16(ref_f32).astype(np.int32)
    b = synthetic.to_int16(est_f32).astype(np.int32)
    return int(np.abs(a - b).max())


PIPELINE_CASES = ["p1_40k_v1_4seg", "p2_32k_v2_protect_index", "p3_48k_v2_single"]


def load_pipeline_golden(name):
    """Fixture minted by tests/golden/make_pipeline_golden.py → dict of everything needed to replay it."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=True)
    cfg_name, secs, tiers, protect, index_rate, f0_up_key, aseed, rseed = [str(x) for x in z["meta"][:8]]
    cfg = NAMED_CONFIGS[cfg_name]
    index_rate = float(index_rate)
    case = dict(cfg=cfg, sd=synthetic.make_state_dict(cfg, seed=0), tiers=ast.literal_eval(tiers), protect=float(protect),
                index_rate=index_rate, f0_up_key=int(f0_up_key), rseed=int(rseed),
                audio=synthetic.make_song(float(secs), seed=int(aseed)), version="v1" if cfg.feat_dim == 256 else "v2",
                hubert=synthetic.FakeHubert(cfg.feat_dim), gold=z)
    idx = synthetic.FakeIndex(cfg.feat_dim, seed=5)
    case["file_index"] = (idx, idx.big_npy) if index_rate > 0 else ""
    return case
