"""Mint golden vectors for the HuBERT / ContentVec front end from the UNMODIFIED reference (build container only).

    python tests/golden/make_hubert_golden.py

Imports the reference's own class `HubertModelWithFinalProj` (/root/reference/lib/infer_pack/loaders.py:10-61,
read-only; it subclasses HuggingFace `transformers.HubertModel`), loads our seeded synthetic weights
(`comfy_rvc_b200.synthetic.make_hubert_state_dict`, HuggingFace key layout) through its `load_state_dict`, and calls its
`extract_features(source, version=...)` for v1 (layer 9 + final_proj -> 256-d) and v2 (layer 12 -> 768-d) exactly as
`VC.vc` does (vc_infer_pipeline.py:48-55).  Fixtures hold the float32 features.
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from comfy_rvc_b200 import synthetic  # noqa: E402

CASES = [
    # name, seconds, weight seed, audio seed
    ("h1_hubert_2s", 2.0, 0, 1),
    ("h2_hubert_400samples", 400 / 16000.0, 0, 2),        # exactly one frame (the receptive field)
    ("h3_hubert_10s", 10.013, 0, 3),                      # odd lengths through every stride-2 layer
]


def reference_model(sd):
    sys.path.insert(0, "/root/reference")
    warnings.filterwarnings("ignore")
    from transformers import HubertConfig
    from lib.infer_pack.loaders import HubertModelWithFinalProj            # type: ignore
    cfg = HubertConfig(**{k: (list(v) if isinstance(v, tuple) else v) for k, v in synthetic.HUBERT_BASE.items()})
    model = HubertModelWithFinalProj(cfg)
    model.load_state_dict({k: v.half().float() for k, v in sd.items()})     # strict: every key must match
    return model.eval()


def main():
    torch.set_num_threads(1)
    for name, secs, wseed, aseed in CASES:
        sd = synthetic.make_hubert_state_dict(wseed)
        model = reference_model(sd)
        source = synthetic.make_speech(secs, seed=aseed)
        out = {}
        for version in ("v1", "v2"):
            with torch.no_grad():
                f = model.extract_features(source, version=version)
            out[f"feats_{version}"] = f.numpy().astype(np.float32)
        out["meta"] = np.array([str(secs), str(wseed), str(aseed), torch.__version__, __import__("transformers").__version__,
                                json.dumps({k: (list(v) if isinstance(v, tuple) else v) for k, v in synthetic.HUBERT_BASE.items()})],
                               dtype=object)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out, allow_pickle=True)
        print(f"{name}: n={source.shape[1]} v1 {out['feats_v1'].shape} v2 {out['feats_v2'].shape} |v2| rms "
              f"{np.sqrt((out['feats_v2'] ** 2).mean()):.3f} -> {path} ({os.path.getsize(path) / 1e3:.0f} kB)")


if __name__ == "__main__":
    main()
