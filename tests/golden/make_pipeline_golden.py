"""Mint pipeline-level golden vectors from the UNMODIFIED reference `VC.pipeline` (build container only).

    python tests/golden/make_pipeline_golden.py        # writes tests/golden/p*.npz

Imports /root/reference/vc_infer_pipeline.py read-only under a synthetic parent package (so its relative
imports resolve without executing the ComfyUI root `__init__.py`), stubbing the absent third-party modules
that the synthesis path never calls (librosa, soundfile, ffmpeg, monotonic_align).  The
reference `VC` object then runs its own `pipeline()` → `vc()` → `net_g.infer()` on the CPU (fp32) with
  * our seeded synthetic checkpoint loaded through the reference's own classes,
  * a deterministic stand-in for HuBERT (`synthetic.FakeHubert`; the real extractor is upstream of the path),
  * a registered synthetic f0 method (`synthetic.pipeline_f0`),
  * optionally a brute-force stand-in for the faiss index (`synthetic.FakeIndex`, passed preloaded as the
    reference allows, pitch_extraction.py:53-55).
The fixtures hold the int16 song exactly as the reference returns it, the segmentation it chose and
per-segment checksums.  The GPU box has no /root/reference; tests replay these files.
"""
from __future__ import annotations

import hashlib
import importlib
import importlib.machinery
import os
import shutil
import sys
import tempfile
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from comfy_rvc_b200.config import NAMED_CONFIGS  # noqa: E402
from comfy_rvc_b200 import synthetic  # noqa: E402

# name, config, seconds, (x_pad, x_query, x_center, x_max), protect, index_rate, f0_up_key, audio seed, rng seed
CASES = [
    ("p1_40k_v1_4seg", "40k", 7.0, (1, 1, 2, 3), 0.5, 0.0, 0, 0, 11),
    ("p2_32k_v2_protect_index", "32k_v2", 5.0, (1, 1, 2, 3), 0.33, 0.75, 3, 1, 12),
    ("p3_48k_v2_single", "48k_v2", 1.5, (1, 1, 2, 3), 0.5, 0.0, -2, 2, 13),
]


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference_pipeline(workdir: str):
    """SURVEY.md §8c recipe.  Returns the reference module `vc_infer_pipeline`."""
    warnings.filterwarnings("ignore")
    import transformers  # noqa: F401  (must be imported before librosa is stubbed)
    _stub("monotonic_align")
    _stub("soundfile")
    _stub("ffmpeg")
    lib = _stub("librosa")
    lib.util = _stub("librosa.util", pad_center=None, tiny=None, normalize=None)
    lib.filters = _stub("librosa.filters", mel=None)
    pkg = types.ModuleType("comfy_rvc_ref")
    pkg.__path__ = ["/root/reference"]
    pkg.__spec__ = importlib.machinery.ModuleSpec("comfy_rvc_ref", None, is_package=True)
    sys.modules["comfy_rvc_ref"] = pkg
    sys.argv = ["x"]
    shutil.copytree("/root/reference/configs", os.path.join(workdir, "configs"))   # config.py:8-19 rewrites them
    os.chdir(workdir)
    return importlib.import_module("comfy_rvc_ref.vc_infer_pipeline")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    torch.set_num_threads(1)
    cwd = os.getcwd()
    work = tempfile.mkdtemp(prefix="rvc_ref_")
    ref = import_reference_pipeline(work)
    from comfy_rvc_ref.lib.infer_pack import models as ref_models  # type: ignore
    for name, cfg_name, secs, tiers, protect, index_rate, f0_up_key, aseed, rseed in CASES:
        cfg = NAMED_CONFIGS[cfg_name]
        sd = synthetic.make_state_dict(cfg, seed=0)
        cls = ref_models.SynthesizerTrnMs256NSFsid if cfg.feat_dim == 256 else ref_models.SynthesizerTrnMs768NSFsid
        net_g = cls(*cfg.to_positional(), is_half=False)
        del net_g.enc_q
        net_g.load_state_dict({k: v.half() for k, v in sd.items()}, strict=False)
        net_g.eval().float()
        conf = types.SimpleNamespace(x_pad=tiers[0], x_query=tiers[1], x_center=tiers[2], x_max=tiers[3],
                                     is_half=False, device="cpu")
        vc = ref.VC(cfg.sr, conf)
        vc.f0_method_dict["synthetic"] = synthetic.pipeline_f0
        version = "v1" if cfg.feat_dim == 256 else "v2"
        hubert = synthetic.FakeHubert(cfg.feat_dim)
        audio = synthetic.make_song(secs, seed=aseed)
        file_index = ""
        if index_rate > 0:
            file_index = (synthetic.FakeIndex(cfg.feat_dim, seed=5), synthetic.FakeIndex(cfg.feat_dim, seed=5).big_npy)
        # record the segments the reference hands to vc()
        calls = []
        orig_vc = vc.vc

        def spy(model, net, sid, audio0, pitch, pitchf, *a, **k):
            out = orig_vc(model, net, sid, audio0, pitch, pitchf, *a, **k)
            calls.append((audio0.shape[0], int(pitch.shape[1]), out.shape[0], sha(out.astype(np.float32)),
                          float(np.abs(out).max())))
            return out

        vc.vc = spy
        torch.manual_seed(rseed)
        times = [0, 0, 0]
        out = vc.pipeline(hubert, net_g, 0, audio.copy(), times, f0_up_key, "synthetic", "median", file_index, index_rate,
                          1, 3, cfg.sr, 0, 1.0, version, protect, 160, False, False, None, 50, 1100)
        assert out.dtype == np.int16
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(
            path, out_i16=out,
            seg_audio_len=np.array([c[0] for c in calls]), seg_frames=np.array([c[1] for c in calls]),
            seg_out_len=np.array([c[2] for c in calls]), seg_sha=np.array([c[3] for c in calls]),
            seg_peak=np.array([c[4] for c in calls]),
            meta=np.array([cfg_name, str(secs), str(tiers), str(protect), str(index_rate), str(f0_up_key), str(aseed),
                           str(rseed), torch.__version__, np.__version__], dtype=object),
            allow_pickle=True)
        print(f"{name}: {len(calls)} segments {[c[0] for c in calls]} -> int16[{out.shape[0]}] peak {np.abs(out).max()} "
              f"-> {path} ({os.path.getsize(path)/1e3:.0f} kB)")
    os.chdir(cwd)
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
