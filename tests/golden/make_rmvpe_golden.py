"""Mint golden vectors for the RMVPE f0 estimator from the UNMODIFIED reference (build container only).

    python tests/golden/make_rmvpe_golden.py

Imports the reference's own module /root/reference/lib/rmvpe.py (read-only) and drives its public class exactly as
`FeatureExtractor.get_rmvpe` does (pitch_extraction.py:191-195): `RMVPE(model_path, is_half=False, device="cpu")`, then
`infer_from_audio(audio, thred=0.03)` (and `infer_from_audio_with_pitch`).  `model_path` is a `torch.save` of our seeded synthetic
weights (`comfy_rvc_b200.synthetic.make_rmvpe_state_dict`; the reference loads it with a strict `load_state_dict`, so the key
layout is checked by the reference itself).  The intermediate `mel` and `hidden` tensors are captured by calling the same
object's `mel_extractor` / `mel2hidden`, as `infer_from_audio` does (rmvpe.py:617-624).

`librosa` is not installed in this image.  The module imports three helpers from it at import time; they are stubbed here:
`pad_center`, `tiny`, `normalize` (only used by the inverse STFT, never on this path; `pad_center` with size == len is the
identity) and `librosa.filters.mel`, for which the stub is `oracle.rmvpe_oracle.mel_filterbank` -- a restatement of librosa's
published algorithm, cross-checked against `transformers.audio_utils.mel_filter_bank` in tests/test_rmvpe_oracle.py.
Everything else (conv1d STFT, DeepUnet, BiGRU, decode) is the reference's code, unmodified.
"""
import os
import sys
import tempfile
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from comfy_rvc_b200 import synthetic  # noqa: E402
from oracle import rmvpe_oracle  # noqa: E402

CASES = [
    # name, samples, weight seed, audio seed
    ("r1_rmvpe_0p5s", 8000, 0, 11),            # 51 frames -> reflect-padded to 64
    ("r2_rmvpe_5s", 80000, 0, 12),             # 501 frames -> 512
    ("r3_rmvpe_1024frames", 163700, 1, 13),    # 1024 frames, no padding; second weight seed
    ("r4_rmvpe_60s", 960000, 0, 14),           # the benchmark's own segment length: 6001 frames -> 6016 (a 6016-step recurrence)
]
# fixtures of long inputs keep every SUB-th frame of mel / hidden (f0 is kept whole)
SUB = {"r4_rmvpe_60s": 8}


def install_librosa_stub():
    lib = types.ModuleType("librosa")
    util = types.ModuleType("librosa.util")
    filters = types.ModuleType("librosa.filters")

    def pad_center(data, size=None, axis=-1, **kw):
        n = data.shape[axis]
        lpad = (size - n) // 2
        widths = [(0, 0)] * data.ndim
        widths[axis] = (lpad, size - n - lpad)
        return np.pad(data, widths)

    util.pad_center = pad_center
    util.tiny = lambda x: np.finfo(np.asarray(x).dtype if np.issubdtype(np.asarray(x).dtype, np.floating) else np.float32).tiny
    util.normalize = lambda S, norm=None, **kw: S
    filters.mel = lambda sr, n_fft, n_mels, fmin, fmax, htk: rmvpe_oracle.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
    lib.util, lib.filters = util, filters
    sys.modules["librosa"], sys.modules["librosa.util"], sys.modules["librosa.filters"] = lib, util, filters


def reference_module():
    import importlib.util
    install_librosa_stub()
    warnings.filterwarnings("ignore")
    spec = importlib.util.spec_from_file_location("ref_rmvpe", "/root/reference/lib/rmvpe.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    torch.set_num_threads(1)
    ref = reference_module()
    models = {}
    only = sys.argv[1:]
    for name, n, wseed, aseed in CASES:
        if only and name not in only:
            continue
        if wseed not in models:
            sd = synthetic.make_rmvpe_state_dict(wseed)
            with tempfile.NamedTemporaryFile(suffix=".pt") as f:
                torch.save(sd, f.name)
                models[wseed] = ref.RMVPE(f.name, is_half=False, device="cpu")
        model = models[wseed]
        audio = synthetic.make_speech(n / 16000.0, seed=aseed)[0].numpy()
        assert audio.shape[0] == n
        f0 = model.infer_from_audio(audio, thred=0.03)
        f0_clip = model.infer_from_audio_with_pitch(audio, thred=0.03, f0_min=50, f0_max=1100)
        with torch.no_grad():
            mel = model.mel_extractor(torch.from_numpy(audio).float().unsqueeze(0), center=True)
            hidden = model.mel2hidden(mel)
        assert np.array_equal(model.decode(hidden.squeeze(0).numpy(), thred=0.03), f0)
        sub = SUB.get(name, 1)
        out = dict(n_samples=np.int64(n), weight_seed=np.int64(wseed), audio_seed=np.int64(aseed), frame_step=np.int64(sub),
                   mel=mel[0].numpy().astype(np.float32)[:, ::sub], hidden=hidden[0].numpy().astype(np.float32)[::sub],
                   f0=np.asarray(f0, dtype=np.float64), f0_with_pitch=np.asarray(f0_clip, dtype=np.float64))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "frames", f0.shape[0], "voiced", int((f0 > 0).sum()), "f0 range", float(f0.min()), float(f0.max()),
              os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
