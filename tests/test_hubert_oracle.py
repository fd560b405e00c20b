"""CPU: the HuBERT / ContentVec oracle (oracle/hubert_oracle.py) reproduces what the reference's own
`HubertModelWithFinalProj.extract_features` (/root/reference/lib/infer_pack/loaders.py:52-61) produced for the same
weights and audio in the build container (tests/golden/make_hubert_golden.py)."""
import os
import types

import numpy as np
import pytest
import torch

from comfy_rvc_b200 import synthetic
from oracle import hubert_oracle
from tests._util import GOLDEN_DIR

HUBERT_CASES = ["h1_hubert_2s", "h2_hubert_400samples", "h3_hubert_10s"]


def load_hubert_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=True)
    secs, wseed, aseed = float(z["meta"][0]), int(z["meta"][1]), int(z["meta"][2])
    return synthetic.make_hubert_state_dict(wseed), synthetic.make_speech(secs, seed=aseed), z


@pytest.mark.parametrize("name", HUBERT_CASES)
def test_hubert_oracle_matches_reference_golden(name):
    torch.set_num_threads(1)
    sd, source, gold = load_hubert_golden(name)
    hcfg = types.SimpleNamespace(**synthetic.HUBERT_BASE)
    assert hubert_oracle.frames_for(source.shape[1], hcfg) == (source.shape[1] - 400) // 320 + 1 == gold["feats_v2"].shape[1]
    for version in ("v1", "v2"):
        f = hubert_oracle.extract_features(sd, hcfg, source, version).numpy()
        ref = gold[f"feats_{version}"]
        assert f.shape == ref.shape
        err = np.abs(f - ref).max()
        print(f"{name} {version}: max |err| {err:.2e} (|ref| max {np.abs(ref).max():.2f})")
        assert err < 5e-5
