"""Mint golden vectors from the UNMODIFIED reference (runs only in the build container).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (/root/reference, read-only, Python) has no tests or known-answer vectors for the
synthesis path, so the pin is the reference's own output: this script imports
`lib/infer_pack/models.py` read-only (stubbing the absent, unused `monotonic_align`,
models.py:11), loads our seeded synthetic checkpoint (`comfy_rvc_b200.synthetic`) through the
reference's own `load_state_dict`, seeds the global RNG and calls the reference `infer()`.
Outputs are stored as small fixtures (int16 PCM exactly as vc_infer_pipeline.py:188-189 would
emit it, a float32 copy, per-stage statistics and SHA-256s).  The GPU box has no
/root/reference; tests there replay these files.
"""
from __future__ import annotations

import hashlib
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from comfy_rvc_b200.config import resolve, state_dict_shapes  # noqa: E402
from comfy_rvc_b200 import synthetic  # noqa: E402

CASES = [
    # name,       config,   B, T,   lengths,     f0 variant, weight seed, input seed, noise seed
    ("c1_40k_v1", "40k",    1, 120, None,        "contour",  0, 1, 7),
    ("c2_48k_v2", "48k_v2", 1, 200, None,        "contour",  0, 1, 7),
    ("c3_32k_v2_ragged", "32k_v2", 2, 96, [96, 71], "uniform", 0, 2, 8),
    ("c4_48k_v2_unvoiced", "48k_v2", 1, 64, None, "unvoiced", 0, 3, 9),
    ("c5_48k_v1_5stage", "48k", 1, 48, None,     "contour",  0, 4, 10),
    ("c6_40k_v1_tiny", "40k", 1, 7, None,        "contour",  0, 5, 11),   # T < window+1
    # no-f0 synthesizers (SynthesizerTrnMs{256,768}NSFsid_nono, models.py:812-1021): config name suffixed ":nono"
    ("c7_40k_v1_nono", "40k:nono", 1, 110, None, "contour",  0, 6, 12),
    ("c8_48k_v2_nono_ragged", "48k_v2:nono", 2, 80, [80, 57], "contour", 0, 7, 13),
    # ResBlock2 decoders (modules.py:311-355, selected at models.py:496 when resblock != "1"); ":rb2x" has kernel/dilation
    # values outside the shipped table
    ("c9_40k_v1_resblock2", "40k:rb2", 1, 100, None, "contour", 0, 8, 14),
    ("c10_48k_v2_resblock2x", "48k_v2:rb2x", 2, 72, [72, 50], "contour", 0, 9, 15),
]


def import_reference():
    sys.modules.setdefault("monotonic_align", types.ModuleType("monotonic_align"))
    sys.path.insert(0, "/root/reference")
    warnings.filterwarnings("ignore")
    from lib.infer_pack import models as ref_models  # type: ignore
    return ref_models


def build_reference_model(ref_models, cfg, sd):
    if cfg.f0:
        cls = ref_models.SynthesizerTrnMs256NSFsid if cfg.feat_dim == 256 else ref_models.SynthesizerTrnMs768NSFsid
        net = cls(*cfg.to_positional(), is_half=False)
    else:
        cls = ref_models.SynthesizerTrnMs256NSFsid_nono if cfg.feat_dim == 256 else ref_models.SynthesizerTrnMs768NSFsid_nono
        net = cls(*cfg.to_positional())
    del net.enc_q                                   # vc_infer_pipeline.py:219
    ref_keys = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    ours = state_dict_shapes(cfg)
    assert ref_keys == ours, (set(ref_keys) ^ set(ours), [k for k in ours if k in ref_keys and ref_keys[k] != ours[k]])
    missing = net.load_state_dict({k: v.half() for k, v in sd.items()}, strict=False)   # fp16 on disk
    assert not missing.missing_keys and not missing.unexpected_keys, missing
    return net.eval().float()


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    torch.set_num_threads(1)        # pin: 1 vs 8 threads already moves 1 LSB
    ref_models = import_reference()
    only = set(sys.argv[1:])
    for name, cfg_name, B, T, lengths, f0v, wseed, iseed, nseed in CASES:
        if only and name not in only:
            continue
        cfg = resolve(cfg_name)
        sd = synthetic.make_state_dict(cfg, seed=wseed)
        net = build_reference_model(ref_models, cfg, sd)
        phone, lens, pitch, pitchf, sid = synthetic.make_inputs(cfg, B, T, seed=iseed, lengths=lengths, f0_variant=f0v)
        taps = {}
        hooks = []
        for i in range(cfg.num_upsamples):
            pass
        torch.manual_seed(nseed)
        with torch.no_grad():
            if cfg.f0:
                o, x_mask, (z, z_p, m_p, logs_p) = net.infer(phone, lens, pitch, pitchf, sid)
            else:
                o, x_mask, (z, z_p, m_p, logs_p) = net.infer(phone, lens, sid)
        # har_source on its own, same RNG position as inside infer: replay the stream
        torch.manual_seed(nseed)
        _ = torch.randn(B, cfg.inter_channels, T)
        with torch.no_grad():
            har = net.dec.m_source(pitchf, net.dec.upp)[0] if cfg.f0 else torch.zeros(B, T * cfg.upp, 1)
        o_np = o[:, 0].numpy()
        out = {
            "o_f32": o_np.astype(np.float32),
            "o_i16": np.stack([synthetic.to_int16(o_np[b, : int(lens[b]) * cfg.upp]) if int(lens[b]) == T
                               else np.pad(synthetic.to_int16(o_np[b, : int(lens[b]) * cfg.upp]), (0, (T - int(lens[b])) * cfg.upp))
                               for b in range(B)]),
            "x_mask": x_mask.numpy(),
            "m_p": m_p.numpy(), "logs_p": logs_p.numpy(), "z_p": z_p.numpy(), "z": z.numpy(),
            "har_source": har.transpose(1, 2).numpy().astype(np.float32),
            "meta": np.array([cfg_name, str(B), str(T), str(lengths), f0v, str(wseed), str(iseed), str(nseed),
                              torch.__version__], dtype=object),
            "sha_o_f32": np.array(sha(o_np.astype(np.float32))),
        }
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out, allow_pickle=True)
        print(f"{name}: o peak {np.abs(o_np).max():.4f} rms {np.sqrt((o_np**2).mean()):.4f} "
              f"|z| {z.abs().mean():.3f} |z-z_p| {(z - z_p).abs().mean():.3f} -> {path} ({os.path.getsize(path)/1e3:.0f} kB)")


if __name__ == "__main__":
    main()
