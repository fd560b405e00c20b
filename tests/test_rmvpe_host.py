"""CPU: host logic of comfy_rvc_b200/rmvpe.py -- the BatchNorm folding, tap / phase weight layouts, image row maps and the
front-end tables reproduce the reference-minted fixtures when the kernels' arithmetic is emulated in fp32 (tests/_emulate_rmvpe.py)."""
import numpy as np
import pytest
import torch

from comfy_rvc_b200 import rmvpe, synthetic
from oracle import rmvpe_oracle
from tests._emulate_rmvpe import hidden_logits
from tests.test_rmvpe_oracle import load_rmvpe_golden


def test_rmvpe_tables_match_oracle():
    assert np.array_equal(rmvpe.mel_filterbank(), rmvpe_oracle.mel_filterbank())
    m = rmvpe.RMVPE(synthetic.make_rmvpe_state_dict(0), device="cuda:0")
    m._build_host()
    rng, fb = m._host["mel.range"].numpy(), m._host["mel.basis"].numpy()
    for i in range(128):
        assert fb[i, :rng[i, 0]].sum() == 0 and fb[i, rng[i, 1]:].sum() == 0 and rng[i, 1] > rng[i, 0]
    basis = rmvpe_oracle.stft_forward_basis()
    assert np.array_equal(m._host["mel.window"].numpy(), basis[0, 0].numpy())          # row 0 of the real part = the window itself
    assert rmvpe.RMVPE._padded_frames(51) == 64 and rmvpe.RMVPE._padded_frames(1024) == 1024 and rmvpe.RMVPE._padded_frames(33) == 64


def test_rmvpe_host_layout_reproduces_reference_fixture():
    torch.set_num_threads(4)
    sd, audio, gold = load_rmvpe_golden("r1_rmvpe_0p5s")
    m = rmvpe.RMVPE(sd, is_half=False, device="cuda:0")
    m._build_host()
    mel = torch.from_numpy(gold["mel"])                                                 # [128][51]
    n_frames = mel.shape[1]
    Tp = m._padded_frames(n_frames)
    src = [j if j < n_frames else 2 * n_frames - 2 - j for j in range(Tp)]
    img = torch.zeros(Tp, 128 + m._pack(0), m._img_c)
    img[:, :128, 0] = (mel[:, src] * m._bn_scale + m._bn_shift).t()
    assert m._pack(0) == 4 and m._pack(1) == 2 and m._pack(2) == 1 and m._img_c == 16
    assert m._geom(0, 64) == (16, 64, 128, 4, 132, 33, 64 * 33) and m._geom(1, 64) == (32, 32, 64, 2, 66, 33, 32 * 33)
    hidden = torch.sigmoid(hidden_logits(m, img, Tp)[:n_frames, :360]).numpy()
    err = np.abs(hidden - gold["hidden"]).max()
    print(f"emulated device path vs reference: max |err| {err:.2e}")
    assert err < 2e-5


def test_rmvpe_has_no_cpu_fallback():
    m = rmvpe.RMVPE(synthetic.make_rmvpe_state_dict(0), device="cuda:0")
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    for call in (lambda: m.infer_from_audio(np.zeros(16000, np.float32)), lambda: m.decode(np.zeros((4, 360), np.float32)),
                 lambda: m.mel2hidden(torch.zeros(1, 128, 64)), lambda: m.mel_extractor(torch.zeros(1, 16000))):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()
    with pytest.raises(NotImplementedError):
        rmvpe.RMVPE(synthetic.make_rmvpe_state_dict(0), onnx=True)
    bad = {k: v for k, v in synthetic.make_rmvpe_state_dict(0).items() if not k.startswith("cnn.")}
    with pytest.raises(RuntimeError, match="missing keys"):
        rmvpe.RMVPE(bad)
