"""CPU: the pipeline oracle reproduces the reference's own `VC.pipeline` outputs (golden fixtures)."""
import hashlib

import numpy as np
import pytest
import torch

from comfy_rvc_b200 import synthetic
from comfy_rvc_b200.config import NAMED_CONFIGS
from oracle import pipeline_oracle, rvc_oracle
from tests._util import PIPELINE_CASES, load_pipeline_golden


@pytest.mark.parametrize("name", PIPELINE_CASES)
def test_pipeline_oracle_matches_reference_golden(name):
    torch.set_num_threads(1)
    case = load_pipeline_golden(name)
    cfg, gold = case["cfg"], case["gold"]
    c = pipeline_oracle.Constants(*case["tiers"], tgt_sr=cfg.sr)
    w = rvc_oracle.fold_weight_norm(case["sd"])
    torch.manual_seed(case["rseed"])                     # the reference drew its noise from the global CPU RNG
    out, parts, opt_ts = pipeline_oracle.pipeline(
        w, cfg, case["hubert"], case["audio"].copy(), c, synthetic.pipeline_f0, f0_up_key=case["f0_up_key"], sid=0,
        file_index=case["file_index"], index_rate=case["index_rate"], version=case["version"], protect=case["protect"],
        return_parts=True)
    # segmentation is integer work: exact
    assert [p.shape[0] + 2 * c.t_pad_tgt for p in parts] == list(gold["seg_out_len"])
    assert out.shape == gold["out_i16"].shape and out.dtype == np.int16
    # the north_star gate for the fp32 path: int16 PCM within +-1 LSB of the reference
    assert np.abs(out.astype(np.int32) - gold["out_i16"].astype(np.int32)).max() <= 1
    peaks = [float(np.abs(p).max()) for p in parts]
    assert max(peaks) <= float(gold["seg_peak"].max()) * (1 + 1e-4)


def test_constants_match_reference_tiers():
    # config.py:124-141: half tier (3,10,60,64), fp32 tier (1,6,38,41), <=4 GB tier (1,5,30,32)
    c = pipeline_oracle.Constants(3, 10, 60, 64, tgt_sr=48000)
    assert (c.t_pad, c.t_pad_tgt, c.t_pad2, c.t_query, c.t_center, c.t_max) == (48000, 144000, 96000, 160000, 960000, 1024000)


def test_split_points_and_segments_cover_audio():
    audio = synthetic.make_song(9.0, seed=3)
    c = pipeline_oracle.Constants(1, 1, 2, 3, tgt_sr=40000)
    ts = pipeline_oracle.split_points(audio, c)
    assert len(ts) == 4 and all(abs(t - (i + 1) * c.t_center) <= c.t_query for i, t in enumerate(ts))
    segs = pipeline_oracle.segments(audio.shape[0], ts, c)
    # trimmed spans tile the audio once; every split point re-synthesises one extra hop (":171" end = t + t_pad2 + window)
    n_pad = audio.shape[0] + 2 * c.t_pad
    covered = sum(((n_pad if e is None else e) - s) - 2 * c.t_pad for s, e in segs)
    assert covered == audio.shape[0] + len(ts) * c.window


def test_pipeline_oracle_with_real_front_ends_matches_reference():
    """p4: the reference pipeline with its own RMVPE (f0_method="rmvpe") and its own HuBERT, nothing stubbed
    (tests/golden/make_pipeline_real_golden.py).  The oracles of the three models chained by the pipeline oracle reproduce the
    reference song to +-1 LSB -- which pins, among other things, a side effect of the reference: HuggingFace's HubertEncoder draws
    `torch.rand([])` per layer from the global generator on every forward, so the synthesizer's noise stream depends on it."""
    import ast
    import os
    import types
    from oracle import hubert_oracle, rmvpe_oracle
    from tests._util import GOLDEN_DIR
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    z = np.load(os.path.join(GOLDEN_DIR, "p4_48k_v2_real_front_ends.npz"), allow_pickle=True)
    cfg_name, secs, tiers, protect, f0_up_key, aseed, rseed = [str(x) for x in z["meta"][:7]]
    cfg = NAMED_CONFIGS[cfg_name]
    hsd, rsd = synthetic.make_hubert_state_dict(0), synthetic.make_rmvpe_state_dict(0)
    hcfg = types.SimpleNamespace(**synthetic.HUBERT_BASE)

    class Hubert:
        def extract_features(self, version="v2", source=None, **k):
            return hubert_oracle.extract_features(hsd, hcfg, source, version)

    seen = {}

    def f0_fn(x=None, **k):
        # the RMVPE oracle's f0 is checked below; the SYNTHESIS gets the reference's own f0: the NSF source integrates f0 into a
        # phase, so even the oracle's 0.002-cent deviations drift into a few LSB over a segment (8 LSB when it is used instead)
        seen["f0"] = rmvpe_oracle.infer_from_audio(rsd, x, thred=0.03)
        return z["f0"].copy()

    c = pipeline_oracle.Constants(*ast.literal_eval(tiers), tgt_sr=cfg.sr)
    audio = synthetic.make_song(float(secs), seed=int(aseed))
    torch.manual_seed(int(rseed))
    out = pipeline_oracle.pipeline(rvc_oracle.fold_weight_norm(synthetic.make_state_dict(cfg, seed=0)), cfg, Hubert(), audio.copy(), c,
                                   f0_fn, f0_up_key=int(f0_up_key), version="v2", protect=float(protect))
    cents = 1200 * np.abs(np.log2(np.maximum(seen["f0"], 1e-3) / np.maximum(z["f0"], 1e-3)))
    d = np.abs(out.astype(np.int32) - z["out_i16"].astype(np.int32))
    print(f"p4: oracle f0 max {cents.max():.4f} cents from the reference's; song max {d.max()} LSB, {100 * np.mean(d > 0):.2f} % differ")
    assert out.shape == z["out_i16"].shape and cents.max() < 0.5 and d.max() <= 1
