"""Mint a pipeline-level golden vector with BOTH real front ends from the UNMODIFIED reference (build container only).

    python tests/golden/make_pipeline_real_golden.py        # writes tests/golden/p4_48k_v2_real_front_ends.npz

Like make_pipeline_golden.py (same import recipe of /root/reference/vc_infer_pipeline.py), but nothing in front of the synthesizer
is a stand-in: the reference `VC.pipeline` runs with
  * `f0_method="rmvpe"` -- the reference's own `RMVPE` class (lib/rmvpe.py) built by `FeatureExtractor.get_rmvpe`'s code path
    (pitch_extraction.py:191-195; the object is attached as `vc.model_rmvpe`, which is what that method caches) on our seeded
    synthetic weights (`librosa` stubbed as in make_rmvpe_golden.py: the mel filterbank is the oracle's restatement);
  * the reference's own `HubertModelWithFinalProj` (lib/infer_pack/loaders.py) on our seeded synthetic weights;
  * the reference synthesizer on the CPU in fp32.
The fixture holds the int16 song, the f0 the reference estimated and the coarse pitch it derived, so a test can tell a front-end
difference from a synthesis difference.
"""
from __future__ import annotations

import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from comfy_rvc_b200.config import NAMED_CONFIGS  # noqa: E402
from comfy_rvc_b200 import synthetic  # noqa: E402
import make_pipeline_golden as mpg  # noqa: E402
import make_rmvpe_golden as mrg  # noqa: E402

NAME, CFG, SECS, TIERS, PROTECT, F0_UP_KEY, ASEED, RSEED = "p4_48k_v2_real_front_ends", "48k_v2", 4.0, (1, 1, 2, 3), 0.33, 0, 4, 14


def main():
    torch.set_num_threads(1)
    cwd = os.getcwd()
    work = tempfile.mkdtemp(prefix="rvc_ref_")
    ref = mpg.import_reference_pipeline(work)
    mrg.install_librosa_stub()                                   # real helpers instead of the import-only placeholders
    import comfy_rvc_ref.lib.rmvpe as ref_rmvpe  # type: ignore
    import librosa.filters as stub_filters
    import librosa.util as stub_util
    ref_rmvpe.mel = stub_filters.mel                             # the names lib/rmvpe.py bound at import time
    ref_rmvpe.pad_center, ref_rmvpe.tiny, ref_rmvpe.normalize = stub_util.pad_center, stub_util.tiny, stub_util.normalize
    from comfy_rvc_ref.lib.infer_pack import models as ref_models  # type: ignore
    from comfy_rvc_ref.lib.infer_pack.loaders import HubertModelWithFinalProj  # type: ignore
    from transformers import HubertConfig
    cfg = NAMED_CONFIGS[CFG]
    sd = synthetic.make_state_dict(cfg, seed=0)
    net_g = ref_models.SynthesizerTrnMs768NSFsid(*cfg.to_positional(), is_half=False)
    del net_g.enc_q
    net_g.load_state_dict({k: v.half() for k, v in sd.items()}, strict=False)
    net_g.eval().float()
    hcfg = HubertConfig(**{k: (list(v) if isinstance(v, tuple) else v) for k, v in synthetic.HUBERT_BASE.items()})
    hubert = HubertModelWithFinalProj(hcfg)
    hubert.load_state_dict({k: v.half().float() for k, v in synthetic.make_hubert_state_dict(0).items()})
    hubert.eval()
    conf = types.SimpleNamespace(x_pad=TIERS[0], x_query=TIERS[1], x_center=TIERS[2], x_max=TIERS[3], is_half=False, device="cpu")
    vc = ref.VC(cfg.sr, conf)
    with tempfile.NamedTemporaryFile(suffix=".pt") as f:
        torch.save(synthetic.make_rmvpe_state_dict(0), f.name)
        vc.model_rmvpe = ref_rmvpe.RMVPE(f.name, is_half=False, device="cpu")
    audio = synthetic.make_song(SECS, seed=ASEED)
    seen = {}
    orig_get_f0 = vc.get_f0

    def spy_f0(*a, **k):
        coarse, f0 = orig_get_f0(*a, **k)
        seen["coarse"], seen["f0"] = np.array(coarse), np.array(f0)
        return coarse, f0

    vc.get_f0 = spy_f0
    torch.manual_seed(RSEED)
    out = vc.pipeline(hubert, net_g, 0, audio.copy(), [0, 0, 0], F0_UP_KEY, "rmvpe", "median", "", 0.0, 1, 3, cfg.sr, 0, 1.0, "v2",
                      PROTECT, 160, False, False, None, 50, 1100)
    assert out.dtype == np.int16
    path = os.path.join(HERE, NAME + ".npz")
    np.savez_compressed(path, out_i16=out, f0=seen["f0"].astype(np.float64), coarse=seen["coarse"].astype(np.int16),
                        meta=np.array([CFG, str(SECS), str(TIERS), str(PROTECT), str(F0_UP_KEY), str(ASEED), str(RSEED),
                                       torch.__version__, np.__version__], dtype=object), allow_pickle=True)
    print(f"{NAME}: int16[{out.shape[0]}] peak {np.abs(out).max()}, f0 frames {seen['f0'].shape[0]} "
          f"range {seen['f0'].min():.1f}-{seen['f0'].max():.1f} Hz -> {path} ({os.path.getsize(path) / 1e3:.0f} kB)")
    os.chdir(cwd)
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
