"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU restatement (numpy / torch-CPU fp32) of the reference's segmented conversion driver around
`net_g.infer`: `VC.pipeline` and `VC.vc` (/root/reference/vc_infer_pipeline.py:116-196, :25-114),
the constants of `FeatureExtractor.__init__` (/root/reference/pitch_extraction.py:14-45) and the f0
post-processing / coarse-pitch quantiser of `get_f0` (pitch_extraction.py:252-302).  The synthesizer
inside is `oracle.rvc_oracle.infer`.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU
legs may import this module.

Parity pin: the reference holds no tests for this driver, so the oracle is pinned against
the reference's own `VC.pipeline` run in the build container by `tests/golden/make_pipeline_golden.py`
(fixtures `tests/golden/p*.npz`, replayed by `tests/test_pipeline_oracle.py`).

NumPy semantics: the final int16 conversion follows NumPy >= 2 promotion (the fixtures were minted with
numpy 2.3): `np.abs(x).max() / 0.99` stays float32.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from scipy import signal

from . import rvc_oracle

MAX_INT16 = 32768  # lib/audio.py:14
_BH, _AH = signal.butter(N=5, Wn=48, btype="high", fs=16000)  # vc_infer_pipeline.py:21


@dataclass
class Constants:
    """pitch_extraction.py:14-32."""
    x_pad: int
    x_query: int
    x_center: int
    x_max: int
    tgt_sr: int
    sr: int = 16000
    window: int = 160

    @property
    def t_pad(self): return self.sr * self.x_pad
    @property
    def t_pad_tgt(self): return self.tgt_sr * self.x_pad
    @property
    def t_pad2(self): return self.t_pad * 2
    @property
    def t_query(self): return self.sr * self.x_query
    @property
    def t_center(self): return self.sr * self.x_center
    @property
    def t_max(self): return self.sr * self.x_max


def split_points(audio: np.ndarray, c: Constants) -> List[int]:
    """Quiet-point search, vc_infer_pipeline.py:124-135 (audio already high-passed)."""
    audio_pad = np.pad(audio, (c.window // 2, c.window // 2), mode="reflect")
    opt_ts: List[int] = []
    if audio_pad.shape[0] > c.t_max:
        audio_sum = np.zeros_like(audio)
        for i in range(c.window):
            audio_sum += audio_pad[i: i - c.window]
        for t in range(c.t_center, audio.shape[0], c.t_center):
            a = np.abs(audio_sum[t - c.t_query: t + c.t_query])
            opt_ts.append(t - c.t_query + int(np.where(a == a.min())[0][0]))
    return opt_ts


def segments(n_audio: int, opt_ts: Sequence[int], c: Constants) -> List[Tuple[int, Optional[int]]]:
    """(start, end) sample ranges into the t_pad-padded audio, vc_infer_pipeline.py:167-180; end None = to the end."""
    segs = []
    s = 0
    t = None
    for t in opt_ts:
        t = t // c.window * c.window
        segs.append((s, t + c.t_pad2 + c.window))
        s = t
    segs.append((t if t is not None else 0, None))
    return segs


def hz_to_mel(hz):  # lib/audio.py:302-304
    return 2595 * np.log10(1 + hz / 700)


def f0_post(f0: np.ndarray, f0_up_key: float, f0_min=50, f0_max=1100, bins: int = 256):
    """pitch_extraction.py:279-302 (no autotune, no f0 file): transpose, mel-quantise to 1..255."""
    f0 = np.array(f0, copy=True)
    f0 *= pow(2, f0_up_key / 12)
    mel_min, mel_max = hz_to_mel(f0_min), hz_to_mel(f0_max)
    f0_mel = hz_to_mel(f0)
    f0_mel = (f0_mel - mel_min) * (bins - 2) / (mel_max - mel_min) + 1
    f0_mel = np.clip(f0_mel, a_min=1, a_max=bins - 1)
    return np.rint(f0_mel).astype(np.int16), f0


def vc_segment(infer_fn, hubert, cfg, audio0: np.ndarray, pitch, pitchf, sid, c: Constants, index, big_npy,
               index_rate: float, version: str, protect: float) -> np.ndarray:
    """`VC.vc`, vc_infer_pipeline.py:25-114, fp32 mode.  `infer_fn(feats, p_len, pitch, pitchf, sid)` → [L] float32."""
    feats = torch.from_numpy(audio0).float().view(1, -1)                                   # :39-47
    feats = hubert.extract_features(version=version, source=feats, padding_mask=torch.zeros_like(feats, dtype=torch.bool),
                                    output_layer=9 if version == "v1" else 12)             # :48-55
    use_f0 = pitch is not None and pitchf is not None
    feats0 = feats.clone() if (protect < 0.5 and use_f0) else None                         # :57-58
    if index is not None and big_npy is not None and index_rate > 0:                       # :59-75
        npy = feats[0].numpy()
        score, ix = index.search(npy, k=1)
        weight = np.square(1 / score)
        weight /= weight.sum(axis=1, keepdims=True)
        npy = np.sum(big_npy[ix] * np.expand_dims(weight, axis=2), axis=1)
        feats = torch.from_numpy(npy).unsqueeze(0) * index_rate + (1 - index_rate) * feats
    feats = F.interpolate(feats.permute(0, 2, 1), scale_factor=2).permute(0, 2, 1)         # :77
    if feats0 is not None:
        feats0 = F.interpolate(feats0.permute(0, 2, 1), scale_factor=2).permute(0, 2, 1)   # :78-81
    p_len = min(audio0.shape[0] // c.window, feats.shape[1])                               # :83
    if use_f0:
        pitch, pitchf = pitch[:, :p_len], pitchf[:, :p_len]                                # :85-87
    if use_f0 and protect < 0.5:                                                           # :89-95
        pitchff = pitchf.clone()
        pitchff[pitchf > 0] = 1
        pitchff[pitchf < 1] = protect
        pitchff = pitchff.unsqueeze(-1)
        feats = feats * pitchff + feats0 * (1 - pitchff)
        feats = feats.to(feats0.dtype)
    return infer_fn(feats, torch.tensor([p_len]).long(), pitch, pitchf, sid)               # :96-105


def to_int16(audio_opt: np.ndarray) -> np.ndarray:
    """vc_infer_pipeline.py:188-189."""
    audio_max = np.abs(audio_opt).max() / 0.99
    return (audio_opt * MAX_INT16 / audio_max).astype(np.int16)


def pipeline(sd_folded, cfg, hubert, audio: np.ndarray, c: Constants, f0_fn, f0_up_key=0, sid: int = 0, file_index="",
             index_rate: float = 0.0, version: str = "v2", protect: float = 0.5, f0_min=50, f0_max=1100,
             noise_fn=None, return_parts: bool = False, if_f0: int = 1):
    """`VC.pipeline`, vc_infer_pipeline.py:116-196 with rms_mix_rate = 1 and no resampling.

    `noise_fn(i, T)` returns the three RNG draws of segment i; default: the global torch CPU RNG in the
    reference's call order (so `torch.manual_seed(s)` before the call reproduces the reference stream)."""
    index, big_npy = file_index if isinstance(file_index, tuple) else (None, None)       # pitch_extraction.py:49-73
    audio = signal.filtfilt(_BH, _AH, audio)                                               # :122
    opt_ts = split_points(audio, c)                                                        # :123-135
    audio_pad = np.pad(audio, (c.t_pad, c.t_pad), mode="reflect")                          # :141
    sid_t = torch.tensor(sid).unsqueeze(0).long()                                          # :151
    pitch = pitchf = None                                                                  # :152
    if if_f0:                                                                              # :153-162
        coarse, f0 = f0_post(f0_fn(x=audio_pad, f0_up_key=f0_up_key, f0_min=f0_min, f0_max=f0_max), f0_up_key, f0_min, f0_max)
        p_len = min(coarse.shape[0], f0.shape[0])
        pitch = torch.from_numpy(coarse[:p_len].astype(np.int64)).unsqueeze(0)
        pitchf = torch.from_numpy(f0[:p_len].astype(np.float32)).unsqueeze(0)

    def infer_fn_factory(i):
        def infer_fn(feats, p_len_t, pitch_s, pitchf_s, sid_s):
            T = int(feats.shape[1])
            if not if_f0:                                                                  # :102-105, one RNG draw (models.py:908)
                nz = noise_fn(i, T)[0] if noise_fn is not None else torch.randn(1, cfg.inter_channels, T)
                return rvc_oracle.infer_nono(sd_folded, cfg, feats, p_len_t, sid_s, nz)[0][0, 0].float().numpy()
            if noise_fn is not None:
                nz, ri, ns = noise_fn(i, T)
            else:
                nz = torch.randn(1, cfg.inter_channels, T)
                ri = torch.rand(1, 1)
                ns = torch.randn(1, T * cfg.upp, 1)
            o = rvc_oracle.infer(sd_folded, cfg, feats, p_len_t, pitch_s, pitchf_s, sid_s, nz, ri, ns)[0]
            return o[0, 0].float().numpy()
        return infer_fn

    parts = []
    for i, (start, end) in enumerate(segments(audio.shape[0], opt_ts, c)):                 # :167-180
        a = audio_pad[start:end]
        ps = pitch[:, start // c.window: (end // c.window if end is not None else None)] if if_f0 else None
        pfs = pitchf[:, start // c.window: (end // c.window if end is not None else None)] if if_f0 else None
        out = vc_segment(infer_fn_factory(i), hubert, cfg, a, ps, pfs, sid_t, c, index, big_npy, index_rate, version, protect)
        parts.append(out[c.t_pad_tgt: -c.t_pad_tgt])
    audio_opt = np.concatenate(parts)                                                      # :182
    res = to_int16(audio_opt)                                                              # :188-189
    return (res, parts, opt_ts) if return_parts else res
