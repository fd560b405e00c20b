"""GPU: the HuBERT / ContentVec front end on the B200 kernels (comfy_rvc_b200/hubert.py, through the C ABI) against the
features the reference's own `HubertModelWithFinalProj.extract_features` produced (tests/golden/make_hubert_golden.py).

The contractions run on fp16 operands with fp32 accumulation (like the reference on CUDA with `is_half`); gate: SNR of the
feature tensor >= 45 dB against the reference's fp32 CPU features (the tolerance BASELINE.json states for the 16-bit path)."""
import numpy as np
import pytest
import torch

import comfy_rvc_b200 as rvc
from comfy_rvc_b200 import synthetic
from comfy_rvc_b200.hubert import HubertB200
from tests.test_hubert_oracle import HUBERT_CASES, load_hubert_golden

pytestmark = pytest.mark.gpu

SNR_GATE_DB = 45.0


@pytest.mark.parametrize("name", HUBERT_CASES)
def test_hubert_features_match_reference(name):
    sd, source, gold = load_hubert_golden(name)
    model = HubertB200(synthetic.HUBERT_BASE, {k: v.half() for k, v in sd.items()}, "cuda:0").eval()
    for version in ("v1", "v2"):
        f = model.extract_features(version=version, source=source.cuda(), padding_mask=None, output_layer=9 if version == "v1" else 12)
        torch.cuda.synchronize()
        ref = gold[f"feats_{version}"]
        got = f.float().cpu().numpy()
        assert got.shape == ref.shape and model.last_launches > 0
        snr = synthetic.snr_db(ref, got)
        print(f"{name} {version}: SNR {snr:.1f} dB, max |err| {np.abs(got - ref).max():.3e} (|ref| max {np.abs(ref).max():.2f}), "
              f"{model.last_launches} launches")
        assert snr >= SNR_GATE_DB


def test_hubert_half_input_and_determinism():
    sd, source, gold = load_hubert_golden("h1_hubert_2s")
    model = HubertB200(synthetic.HUBERT_BASE, sd, "cuda:0")
    a = model.extract_features(version="v2", source=source.cuda().half())
    b = model.extract_features(version="v2", source=source.cuda().half())
    assert a.dtype == torch.float16 and torch.equal(a, b)
    assert synthetic.snr_db(gold["feats_v2"], a.float().cpu().numpy()) >= 40.0      # fp16 audio + fp16 output rounding
    with pytest.raises(ValueError):
        model.extract_features(version="v2", source=torch.zeros(2, 1000).cuda())


def test_hubert_feeds_the_segment_driver():
    """`VC.vc` with the B200 front end as `model` (vc_infer_pipeline.py:48-55) end to end on the device."""
    from comfy_rvc_b200 import pipeline as pl
    from comfy_rvc_b200.config import NAMED_CONFIGS
    from tests.test_parity_gpu import build_net
    cfg = NAMED_CONFIGS["48k_v2"]
    net = build_net(cfg, synthetic.make_state_dict(cfg), "fp16")
    model = HubertB200(synthetic.HUBERT_BASE, synthetic.make_hubert_state_dict(0), "cuda:0")
    vc = pl.VC(cfg.sr, pl.PipelineConfig(1, 6, 38, 41, is_half=True, device="cuda:0"))
    audio = synthetic.make_speech(3.0, seed=4)[0].numpy()
    frames = audio.shape[0] // 160
    f0 = synthetic.make_f0(frames)
    pitch = torch.from_numpy(synthetic.coarse_pitch(f0))[None].cuda()
    pitchf = torch.from_numpy(f0)[None].cuda()
    out = vc.vc(model, net, torch.zeros(1, dtype=torch.int64).cuda(), audio, pitch, pitchf, [0, 0, 0], None, None, 0.0, "v2", 0.33)
    assert out.dtype == np.float32 and out.shape[0] == min(frames, 2 * ((audio.shape[0] - 400) // 320 + 1)) * cfg.upp
    assert np.isfinite(out).all() and np.abs(out).max() > 1e-3
