"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU fp32 functional restatement of the RMVPE f0 estimator: /root/reference/lib/rmvpe.py
  * `MelSpectrogram.forward` (:489-556) with the reference's conv1d STFT (`STFT.__init__` :86-115, `.transform` :117-152);
  * `E2E.forward` (:431-472) = `DeepUnet` (:395-428: `Encoder` :270-305, `ResEncoderBlock` :308-327, `ConvBlockRes` :232-267,
    `Intermediate` :330-347, `ResDecoderBlock` :350-377, `Decoder` :380-392) -> `cnn` -> `BiGRU` (:217-229) -> Linear -> Sigmoid;
  * `RMVPE.mel2hidden` (:591-608), `.decode` (:610-615), `.to_local_average_cents` (:658-684), `.infer_from_audio` (:617-624),
    `.infer_from_audio_with_pitch` (:646-656).

Third-party arithmetic that is NOT under /root/reference and not installed in this image: `librosa.filters.mel(htk=True)`
(librosa is unpinned in the reference's requirements).  `mel_filterbank` below restates its published algorithm (HTK mel
scale, triangular filters built from `np.subtract.outer`, Slaney area normalisation, float32 result);
`tests/test_rmvpe_oracle.py` cross-checks it against the independent implementation in `transformers.audio_utils`.

Parity pin: `tests/golden/make_rmvpe_golden.py` runs the reference's own classes (lib/rmvpe.py imported read-only, `librosa`
stubbed with the three helpers it imports) on seeded weights and audio in the build container; `tests/test_rmvpe_oracle.py`
replays the fixtures against this file.  Only `tests/` and `bench.py`'s incumbent leg (`rmvpe_incumbent`: this functional form on CUDA
tensors as the eager-PyTorch figure to compare with) may import this module.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

N_FFT, HOP, N_MELS, SR, FMIN, FMAX, CLAMP = 1024, 160, 128, 16000, 30.0, 8000.0, 1e-5
N_CLASS = 360
BN_EPS = 1e-5


# ---- mel front end ---------------------------------------------------------------------------------------------------------
def mel_filterbank(sr: int = SR, n_fft: int = N_FFT, n_mels: int = N_MELS, fmin: float = FMIN, fmax: float = FMAX) -> np.ndarray:
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=True) (norm="slaney", dtype float32) -> [n_mels, 1 + n_fft/2]."""
    hz_to_mel = lambda f: 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)
    mel_to_hz = lambda m: 700.0 * (10.0 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def stft_forward_basis(n_fft: int = N_FFT) -> torch.Tensor:
    """rmvpe.py:93-115: rows = real then imaginary part of the DFT matrix, times the periodic Hann window -> [n_fft + 2, 1, n_fft]."""
    from scipy.signal import get_window
    fourier_basis = np.fft.fft(np.eye(n_fft))
    cutoff = n_fft // 2 + 1
    fourier_basis = np.vstack([np.real(fourier_basis[:cutoff, :]), np.imag(fourier_basis[:cutoff, :])])
    forward_basis = torch.FloatTensor(fourier_basis[:, None, :])
    fft_window = torch.from_numpy(get_window("hann", n_fft, fftbins=True)).float()
    forward_basis *= fft_window
    return forward_basis.float()


_BASIS_CACHE: Dict[str, torch.Tensor] = {}


@torch.no_grad()
def log_mel(audio: torch.Tensor) -> torch.Tensor:
    """rmvpe.py:489-556 with keyshift 0, speed 1, center=True, is_half False: audio [B, n] -> log-mel [B, 128, 1 + n // 160]."""
    if "basis" not in _BASIS_CACHE:
        _BASIS_CACHE["basis"] = stft_forward_basis()
        _BASIS_CACHE["mel"] = torch.from_numpy(mel_filterbank()).float()
    x = audio.float()
    x = F.pad(x[:, None, None, :], (N_FFT // 2, N_FFT // 2, 0, 0, 0, 0), mode="reflect").squeeze(1)          # :132-136
    ft = F.conv1d(x, _BASIS_CACHE["basis"], stride=HOP, padding=0)                                                # :139-141
    cutoff = N_FFT // 2 + 1
    magnitude = torch.sqrt(ft[:, :cutoff, :] ** 2 + ft[:, cutoff:, :] ** 2)                                       # :147
    mel_output = torch.matmul(_BASIS_CACHE["mel"], magnitude)                                                     # :550
    return torch.log(torch.clamp(mel_output, min=CLAMP))                                                          # :553


# ---- DeepUnet --------------------------------------------------------------------------------------------------------------
def _bn(x, w, p):
    return F.batch_norm(x, w[p + "running_mean"], w[p + "running_var"], w[p + "weight"], w[p + "bias"], False, 0.0, BN_EPS)


def _conv_block_res(x, w, p):
    """rmvpe.py:232-267: relu(bn(conv3x3(relu(bn(conv3x3(x)))))) + (shortcut 1x1 conv(x) if channels change else x)."""
    h = F.relu(_bn(F.conv2d(x, w[p + "conv.0.weight"], None, padding=1), w, p + "conv.1."))
    h = F.relu(_bn(F.conv2d(h, w[p + "conv.3.weight"], None, padding=1), w, p + "conv.4."))
    if p + "shortcut.weight" in w:
        return h + F.conv2d(x, w[p + "shortcut.weight"], w[p + "shortcut.bias"])
    return h + x


@torch.no_grad()
def e2e_forward(sd: Dict[str, torch.Tensor], mel: torch.Tensor, n_blocks: int = 4, en_de_layers: int = 5, inter_layers: int = 4,
                taps=None, gru=None, dtype=torch.float32) -> torch.Tensor:
    """rmvpe.py:465-472: mel [B, 128, T] (T a multiple of 32) -> salience [B, T, 360].
    `gru` (optional): a callable replacing the explicit recurrence below, e.g. a `torch.nn.GRU` holding the same weights -- used by
    bench.py's `rmvpe_incumbent` to time the library (cuDNN) path the reference takes on a GPU; `dtype` float16 = the reference's is_half."""
    w = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    x = mel.to(dtype).transpose(-1, -2).unsqueeze(1)                                 # [B, 1, T, 128]
    x = _bn(x, w, "unet.encoder.bn.")                                                # :299
    skips = []
    for i in range(en_de_layers):                                                    # :300-303, :320-327
        for j in range(n_blocks):
            x = _conv_block_res(x, w, f"unet.encoder.layers.{i}.conv.{j}.")
        skips.append(x)
        x = F.avg_pool2d(x, kernel_size=(2, 2))
    if taps is not None:
        taps["enc"] = x
    for i in range(inter_layers):                                                    # :343-347
        for j in range(n_blocks):
            x = _conv_block_res(x, w, f"unet.intermediate.layers.{i}.conv.{j}.")
    if taps is not None:
        taps["inter"] = x
    for i in range(en_de_layers):                                                    # :389-392, :371-377
        p = f"unet.decoder.layers.{i}."
        x = F.conv_transpose2d(x, w[p + "conv1.0.weight"], None, stride=(2, 2), padding=(1, 1), output_padding=(1, 1))
        x = F.relu(_bn(x, w, p + "conv1.1."))
        x = torch.cat((x, skips[-1 - i]), dim=1)
        for j in range(n_blocks):
            x = _conv_block_res(x, w, p + f"conv2.{j}.")
    if taps is not None:
        taps["unet"] = x
    x = F.conv2d(x, w["cnn.weight"], w["cnn.bias"], padding=1)                       # :468
    x = x.transpose(1, 2).flatten(-2)                                                # [B, T, 3 * 128]
    if taps is not None:
        taps["gru_in"] = x
    x = gru(x) if gru is not None else bigru(x, w, "fc.0.gru.")
    if taps is not None:
        taps["gru_out"] = x
    return torch.sigmoid(F.linear(x, w["fc.1.weight"], w["fc.1.bias"]))              # :451-456 (Dropout is identity in eval)


def bigru(x: torch.Tensor, w: Dict[str, torch.Tensor], p: str) -> torch.Tensor:
    """torch.nn.GRU(384, 256, 1, batch_first=True, bidirectional=True) (rmvpe.py:217-229), gates ordered r | z | n:
    r = s(W_ir x + b_ir + W_hr h + b_hr), z likewise, n = tanh(W_in x + b_in + r * (W_hn h + b_hn)), h' = (1 - z) n + z h."""
    B, T, _ = x.shape
    outs = []
    for sfx, order in (("", range(T)), ("_reverse", range(T - 1, -1, -1))):
        w_ih, w_hh = w[p + "weight_ih_l0" + sfx], w[p + "weight_hh_l0" + sfx]
        b_ih, b_hh = w[p + "bias_ih_l0" + sfx], w[p + "bias_hh_l0" + sfx]
        H = w_hh.shape[1]
        gi = F.linear(x, w_ih, b_ih)
        h = torch.zeros(B, H, dtype=x.dtype, device=x.device)
        out = torch.empty(B, T, H, dtype=x.dtype, device=x.device)
        for t in order:
            gh = F.linear(h, w_hh, b_hh)
            r = torch.sigmoid(gi[:, t, :H] + gh[:, :H])
            z = torch.sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
            n = torch.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])
            h = (1 - z) * n + z * h
            out[:, t] = h
        outs.append(out)
    return torch.cat(outs, dim=-1)


@torch.no_grad()
def mel2hidden(sd, mel: torch.Tensor, taps=None) -> torch.Tensor:
    """rmvpe.py:591-608: reflect-pad the frame axis to a multiple of 32, run the model, cut back."""
    n_frames = mel.shape[-1]
    padding = min(32 * ((n_frames - 1) // 32 + 1) - n_frames, n_frames)
    mel = F.pad(mel, (0, padding), mode="reflect")
    return e2e_forward(sd, mel, taps=taps)[:, :n_frames]


# ---- decode ----------------------------------------------------------------------------------------------------------------
CENTS_MAPPING = np.pad(20 * np.arange(N_CLASS) + 1997.3794084376191, (4, 4))          # rmvpe.py:588-589


def to_local_average_cents(salience: np.ndarray, thred: float = 0.05) -> np.ndarray:
    """rmvpe.py:658-684: salience-weighted mean of the cents of the 9 bins around the arg-max, 0 where the max <= thred."""
    center = np.argmax(salience, axis=1)
    salience = np.pad(salience, ((0, 0), (4, 4)))
    center += 4
    idx = center[:, None] + np.arange(-4, 5)[None, :]
    todo_salience = np.take_along_axis(salience, idx, axis=1)
    todo_cents = CENTS_MAPPING[idx]
    product_sum = np.sum(todo_salience * todo_cents, 1)
    weight_sum = np.sum(todo_salience, 1)
    devided = product_sum / weight_sum
    maxx = np.max(salience, axis=1)
    devided[maxx <= thred] = 0
    return devided


def decode(hidden: np.ndarray, thred: float = 0.03) -> np.ndarray:
    """rmvpe.py:610-615."""
    cents_pred = to_local_average_cents(hidden, thred=thred)
    f0 = 10 * (2 ** (cents_pred / 1200))
    f0[f0 == 10] = 0
    return f0


@torch.no_grad()
def infer_from_audio(sd, audio: np.ndarray, thred: float = 0.03, taps=None) -> np.ndarray:
    """rmvpe.py:617-624 (is_half False)."""
    mel = log_mel(torch.from_numpy(np.asarray(audio)).float()[None])
    if taps is not None:
        taps["mel"] = mel
    hidden = mel2hidden(sd, mel, taps=taps)
    if taps is not None:
        taps["hidden"] = hidden
    return decode(hidden.squeeze(0).numpy(), thred=thred)


def infer_from_audio_with_pitch(sd, audio: np.ndarray, thred: float = 0.03, f0_min: float = 50, f0_max: float = 1100) -> np.ndarray:
    """rmvpe.py:646-656."""
    return np.clip(infer_from_audio(sd, audio, thred), a_min=f0_min, a_max=f0_max)
