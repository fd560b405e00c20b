"""CPU: the pipeline oracle reproduces the reference's own `VC.pipeline` outputs (golden fixtures)."""
import hashlib

import numpy as np
import pytest
import torch

from comfy_rvc_b200 import synthetic
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
