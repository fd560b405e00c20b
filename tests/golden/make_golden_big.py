"""Mint the BASELINE-size golden vectors from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_big.py [g1 g2 g3]

Same recipe as make_golden.py (reference `lib/infer_pack/models.py` imported read-only, our seeded
synthetic checkpoint through the reference's own `load_state_dict`, seeded global RNG, reference
`infer()`), at the sizes BASELINE.json's `configs` are quoted on:

  g1_40k_v1_T1000    configs[0]: 40k v1, B=1, T=1000 (10 s)               — 1 CPU thread
  g2_48k_v2_T6000    configs[1]: 48k_v2, B=1, T=6000 (60 s, 2.88 M samples) — 1 CPU thread
  g3_32k_v2_B64      configs[2]: 32k_v2, B=64 ragged lengths 600..800     — 6 CPU threads

To stay small the fixtures hold the int16 PCM exactly as vc_infer_pipeline.py:188-189 emits it, the
peak it was normalised with (so float = i16 * audio_max / 32768 to within one truncation step), the SHA-256 of
the float32 output, and every `STRIDE`-th frame of m_p / logs_p / z_p / z.  g3 keeps 8 of its 64 items.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from comfy_rvc_b200.config import NAMED_CONFIGS  # noqa: E402
from comfy_rvc_b200 import synthetic  # noqa: E402
from make_golden import build_reference_model, import_reference, sha  # noqa: E402

STRIDE = 25
G3_ITEMS = [0, 5, 13, 21, 34, 47, 55, 63]


def g3_lengths():
    rng = np.random.default_rng(64)
    lens = rng.integers(600, 801, size=64)
    lens[0], lens[63], lens[13] = 800, 600, 800
    return [int(x) for x in lens]


BIG_CASES = {
    # name: (file, config, B, T, lengths, f0 variant, weight seed, input seed, noise seed, threads, kept items)
    "g1": ("g1_40k_v1_T1000", "40k", 1, 1000, None, "contour", 0, 1, 7, 1, [0]),
    "g2": ("g2_48k_v2_T6000", "48k_v2", 1, 6000, None, "contour", 0, 1, 7, 1, [0]),
    "g3": ("g3_32k_v2_B64", "32k_v2", 64, 800, g3_lengths(), "contour", 0, 2, 8, 6, G3_ITEMS),
}


def main():
    ref_models = import_reference()
    for key in (sys.argv[1:] or list(BIG_CASES)):
        name, cfg_name, B, T, lengths, f0v, wseed, iseed, nseed, threads, keep = BIG_CASES[key]
        torch.set_num_threads(threads)
        cfg = NAMED_CONFIGS[cfg_name]
        sd = synthetic.make_state_dict(cfg, seed=wseed)
        net = build_reference_model(ref_models, cfg, sd)
        phone, lens, pitch, pitchf, sid = synthetic.make_inputs(cfg, B, T, seed=iseed, lengths=lengths, f0_variant=f0v)
        torch.manual_seed(nseed)
        t0 = time.time()
        with torch.no_grad():
            o, x_mask, (z, z_p, m_p, logs_p) = net.infer(phone, lens, pitch, pitchf, sid)
        dt = time.time() - t0
        o_np = o[:, 0].numpy().astype(np.float32)
        i16, peaks, shas = [], [], []
        for b in keep:
            n = int(lens[b]) * cfg.upp
            seg = o_np[b, :n]
            peaks.append(np.float64(np.abs(seg).max() / 0.99))
            i16.append(np.pad(synthetic.to_int16(seg), (0, T * cfg.upp - n)))
            shas.append(sha(seg))
        out = {
            "o_i16": np.stack(i16), "audio_max": np.array(peaks), "sha_o_f32": np.array(shas), "items": np.array(keep),
            "lengths": lens.numpy(),
            "x_mask_sum": x_mask.sum(dim=(1, 2)).numpy(),
            "m_p": m_p[keep][:, :, ::STRIDE].numpy(), "logs_p": logs_p[keep][:, :, ::STRIDE].numpy(),
            "z_p": z_p[keep][:, :, ::STRIDE].numpy(), "z": z[keep][:, :, ::STRIDE].numpy(),
            "meta": np.array([cfg_name, str(B), str(T), str(lengths), f0v, str(wseed), str(iseed), str(nseed),
                              torch.__version__, str(threads), str(STRIDE)], dtype=object),
        }
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out, allow_pickle=True)
        print(f"{name}: reference infer {dt:.1f} s on {threads} thread(s); o peak {np.abs(o_np).max():.4f} "
              f"rms {np.sqrt((o_np ** 2).mean()):.4f} -> {path} ({os.path.getsize(path) / 1e6:.2f} MB)", flush=True)


if __name__ == "__main__":
    main()
