"""Test double: the product `VC` driver with its three device hooks replaced by the CPU oracle.

The product class (`comfy_rvc_b200.pipeline.VC`) has no CPU path; this subclass exists only so the HOST logic of the
song-level driver (planning, f0 post-processing, segment→rank assignment, reference-order noise under sharding, the
peak exchange and the ordered gather over `torch.distributed`) can run in world_size-2 gloo tests without a GPU."""
import numpy as np
import torch

from comfy_rvc_b200.pipeline import VC
from oracle import pipeline_oracle, rvc_oracle


class OracleNet:
    """Looks like a loaded synthesizer to the driver (`.cfg`), computes with the oracle."""

    def __init__(self, cfg, sd):
        self.cfg = cfg
        self.w = rvc_oracle.fold_weight_norm(sd)


class HostVC(VC):
    def _stage(self, audio_pad, pitch, pitchf, sid, net_g, staged_audio=None):
        return {"dev": torch.device("cpu"), "audio": audio_pad, "sid": torch.tensor(sid).reshape(1).long(),
                "pitch": torch.from_numpy(pitch).unsqueeze(0), "pitchf": torch.from_numpy(pitchf).unsqueeze(0)}

    def _convert(self, staged, model, net_g, s, T_formula, index, big_npy, index_rate, version, protect, noise):
        c = pipeline_oracle.Constants(self.x_pad, self.x_query, self.x_center, self.x_max, tgt_sr=net_g.cfg.sr)

        def infer_fn(feats, p_len_t, pitch_s, pitchf_s, sid_s):
            assert feats.shape[1] == T_formula
            nz = noise() if callable(noise) else noise          # "reference" mode: drawn after the front end, like the reference
            return rvc_oracle.infer(net_g.w, net_g.cfg, feats, p_len_t, pitch_s, pitchf_s, sid_s, *nz)[0][0, 0].numpy()

        out = pipeline_oracle.vc_segment(infer_fn, model, net_g.cfg, staged["audio"][s.start:s.end],
                                         staged["pitch"][:, s.f0_start:s.f0_end], staged["pitchf"][:, s.f0_start:s.f0_end],
                                         staged["sid"], c, index, big_npy, index_rate, version, protect)
        return out[self.t_pad_tgt: out.shape[0] - self.t_pad_tgt]

    def _finalize(self, staged, parts, exchange_peak):
        local = np.concatenate([p for _, p in parts]) if parts else np.zeros(0, np.float32)
        peak = np.abs(local).max() if local.size else np.float32(0)
        if exchange_peak is not None:
            peak = np.float32(exchange_peak(float(peak)))
        audio_max = peak / 0.99
        return {i: (p * 32768 / audio_max).astype(np.int16) for i, p in parts}
