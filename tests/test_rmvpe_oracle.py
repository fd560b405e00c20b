"""CPU: the RMVPE oracle (oracle/rmvpe_oracle.py) reproduces what the reference's own `RMVPE` class
(/root/reference/lib/rmvpe.py:559-684) produced for the same seeded weights and audio in the build container
(tests/golden/make_rmvpe_golden.py), and its restated third-party pieces agree with independent implementations."""
import os

import numpy as np
import pytest
import torch

from comfy_rvc_b200 import synthetic
from oracle import rmvpe_oracle
from tests._util import GOLDEN_DIR

RMVPE_CASES = ["r1_rmvpe_0p5s", "r2_rmvpe_5s", "r3_rmvpe_1024frames"]


def load_rmvpe_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n, wseed, aseed = int(z["n_samples"]), int(z["weight_seed"]), int(z["audio_seed"])
    audio = synthetic.make_speech(n / 16000.0, seed=aseed)[0].numpy()
    assert audio.shape[0] == n                          # (fixtures of long inputs keep every `frame_step`-th frame of mel / hidden)
    return synthetic.make_rmvpe_state_dict(wseed), audio, z


def logit(p):
    p = np.clip(np.asarray(p, dtype=np.float64), 1e-7, 1 - 1e-7)
    return np.log(p) - np.log1p(-p)


@pytest.mark.parametrize("name", RMVPE_CASES[:2])
def test_rmvpe_oracle_matches_reference_golden(name):
    torch.set_num_threads(1)
    sd, audio, gold = load_rmvpe_golden(name)
    taps = {}
    f0 = rmvpe_oracle.infer_from_audio(sd, audio, thred=0.03, taps=taps)
    mel, hidden = taps["mel"][0].numpy(), taps["hidden"][0].numpy()
    assert mel.shape == gold["mel"].shape and hidden.shape == gold["hidden"].shape and f0.shape == gold["f0"].shape
    assert mel.shape[1] == audio.shape[0] // 160 + 1
    e_mel, e_hid = np.abs(mel - gold["mel"]).max(), np.abs(hidden - gold["hidden"]).max()
    cents = 1200 * np.abs(np.log2(np.maximum(f0, 1e-3) / np.maximum(gold["f0"], 1e-3)))
    print(f"{name}: mel max |err| {e_mel:.2e}, hidden max |err| {e_hid:.2e}, f0 max {cents.max():.3f} cents")
    assert e_mel < 1e-4 and e_hid < 1e-4
    assert np.mean(cents < 0.5) > 0.99          # an arg-max between two near-equal bins may flip on a 1e-6 difference
    f0c = rmvpe_oracle.infer_from_audio_with_pitch(sd, audio, thred=0.03, f0_min=50, f0_max=1100)
    assert f0c.min() >= 50 and f0c.max() <= 1100
    assert np.mean(np.abs(f0c - gold["f0_with_pitch"]) < 0.05) > 0.99


def test_mel_filterbank_matches_independent_implementation():
    """`librosa.filters.mel(htk=True)` restated in the oracle vs `transformers.audio_utils.mel_filter_bank` (htk scale, slaney
    norm), which documents itself as librosa-compatible."""
    au = pytest.importorskip("transformers.audio_utils")
    ours = rmvpe_oracle.mel_filterbank()
    theirs = au.mel_filter_bank(num_frequency_bins=513, num_mel_filters=128, min_frequency=30.0, max_frequency=8000.0,
                                sampling_rate=16000, norm="slaney", mel_scale="htk").T
    assert ours.shape == theirs.shape == (128, 513) and ours.dtype == np.float32
    assert np.abs(ours - theirs).max() < 1e-6 * np.abs(theirs).max() + 1e-9
    assert (ours >= 0).all() and (ours.sum(1) > 0).all()


def test_stft_basis_is_a_windowed_dft():
    """The reference's conv1d STFT (rmvpe.py:86-152) against numpy's FFT of the same frames."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal(16000 // 4).astype(np.float32)
    mel = rmvpe_oracle.log_mel(torch.from_numpy(x)[None])[0].numpy()
    xp = np.pad(x.astype(np.float64), 512, mode="reflect")
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(1024) / 1024)
    frames = np.stack([xp[i * 160:i * 160 + 1024] * win for i in range(x.shape[0] // 160 + 1)])
    mag = np.abs(np.fft.rfft(frames, axis=1)).T
    ref = np.log(np.maximum(rmvpe_oracle.mel_filterbank().astype(np.float64) @ mag, 1e-5))
    assert np.abs(mel - ref).max() < 2e-4


def test_decode_restatement_matches_reference_loop():
    """`to_local_average_cents` vectorised vs the reference's per-frame loop (rmvpe.py:658-684), written out here."""
    rng = np.random.default_rng(3)
    T = 300
    sal = rng.random((T, 360)).astype(np.float32) * 0.02
    peaks = rng.integers(0, 360, T)
    for t in range(T):
        if t % 7:
            lo, hi = max(0, peaks[t] - 3), min(360, peaks[t] + 4)
            sal[t, lo:hi] += rng.random(hi - lo).astype(np.float32) * 0.9
    sal[5, 0] = 0.99
    sal[6, 359] = 0.98
    cm = np.pad(20 * np.arange(360) + 1997.3794084376191, (4, 4))
    want = np.zeros(T)
    for t in range(T):
        c = int(np.argmax(sal[t])) + 4
        row = np.pad(sal[t], (4, 4))
        s, m = row[c - 4:c + 5], cm[c - 4:c + 5]
        want[t] = np.sum(s * m) / np.sum(s) if row.max() > 0.03 else 0
    got = rmvpe_oracle.to_local_average_cents(sal.copy(), thred=0.03)
    assert np.array_equal(got, want)
    f0 = rmvpe_oracle.decode(sal.copy(), thred=0.03)
    assert ((f0 == 0) == (want == 0)).all() and (want == 0).sum() > 10


def test_synthetic_rmvpe_state_dict_is_deterministic():
    a, b = synthetic.make_rmvpe_state_dict(0), synthetic.make_rmvpe_state_dict(0)
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    assert len(a) == 741 and a["fc.0.gru.weight_hh_l0_reverse"].shape == (768, 256)
