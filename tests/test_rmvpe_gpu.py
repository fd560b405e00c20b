"""GPU: the RMVPE f0 estimator on the B200 kernels (comfy_rvc_b200/rmvpe.py, through the C ABI) against what the reference's own
`RMVPE` class produced (tests/golden/make_rmvpe_golden.py) and, per kernel, against the CPU oracle (oracle/rmvpe_oracle.py).

Tolerances (floating point; the contractions run on fp16 operands with fp32 accumulation, the residual stream, the GRU recurrence
and the mel front end in fp32, the decode in fp64):
  * log-mel: max |err| <= 5e-3, mean |err| <= 2e-5 in the log domain.  The reference's fp32 DFT-matrix convolution is itself
    1.5e-3 (max) / 5e-6 (mean) away from the exact transform on these signals (bins near the 1e-5 clamp floor); the FFT here is
    ~10x closer to the exact values than the reference is, so the bound is the reference's own rounding noise;
  * salience (`hidden`): SNR of logit(hidden) >= 40 dB and max |err| of hidden <= 0.02 against the reference's fp32 CPU output;
  * f0 given the SAME salience (decode kernel): relative 1e-12 (float64 like numpy);
  * f0 end to end: >= 97 % of the frames within 5 cents -- with seeded random weights the salience has near-tied maxima, and an
    arg-max flip moves the 9-bin window; the remaining frames must still be a valid decode of our own salience.
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from comfy_rvc_b200 import _lib, synthetic
from comfy_rvc_b200.rmvpe import RMVPE
from comfy_rvc_b200.weights import pack_tc
from oracle import rmvpe_oracle
from tests.test_rmvpe_oracle import RMVPE_CASES, load_rmvpe_golden, logit

pytestmark = pytest.mark.gpu

_MODELS = {}


def model_for(wseed):
    if wseed not in _MODELS:
        _MODELS[wseed] = RMVPE(synthetic.make_rmvpe_state_dict(wseed), is_half=False, device="cuda:0")
    return _MODELS[wseed]


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def mel_close(got, ref):
    e = np.abs(got - ref)
    return e.max() <= 5e-3 and e.mean() <= 2e-5


def cents_between(a, b):
    return 1200 * np.abs(np.log2(np.maximum(a, 1e-3) / np.maximum(b, 1e-3)))


@pytest.mark.parametrize("name", RMVPE_CASES + ["r4_rmvpe_60s"])
def test_rmvpe_matches_reference(name):
    """r4 is the benchmark's segment length (60 s, 6016 recurrence steps); its fixture keeps every 8th frame of mel / hidden."""
    sd, audio, gold = load_rmvpe_golden(name)
    model = model_for(int(gold["weight_seed"]))
    taps = {}
    f0 = model.infer_from_audio(audio, thred=0.03, taps=taps)
    torch.cuda.synchronize()
    step = int(gold["frame_step"]) if "frame_step" in gold.files else 1
    mel_full, hidden_full = taps["mel"].cpu().numpy(), taps["hidden"].cpu().numpy()
    assert mel_full.shape == (128, f0.shape[0]) and hidden_full.shape == (f0.shape[0], 360)
    mel, hidden = mel_full[:, ::step], hidden_full[::step]
    assert f0.dtype == np.float64 and f0.shape == gold["f0"].shape and mel.shape == gold["mel"].shape and hidden.shape == gold["hidden"].shape
    e_mel = np.abs(mel - gold["mel"]).max()
    e_hid = np.abs(hidden - gold["hidden"]).max()
    snr = synthetic.snr_db(logit(gold["hidden"]), logit(hidden))
    cents = cents_between(f0, gold["f0"])
    frac = float(np.mean(cents < 5.0))
    print(f"{name}: mel max |err| {e_mel:.2e}; hidden max |err| {e_hid:.2e}, logit SNR {snr:.1f} dB; f0 within 5 cents: {100 * frac:.1f} % "
          f"(median {np.median(cents):.4f} cents); {model.last_launches} launches")
    assert mel_close(mel, gold["mel"])
    assert snr >= 40.0 and e_hid <= 0.02
    assert frac >= 0.97
    # every frame is the reference decode of OUR salience (so a differing frame is an arg-max flip, not a decode error)
    assert np.allclose(f0, rmvpe_oracle.decode(hidden_full.copy(), thred=0.03), rtol=1e-12, atol=0)


def test_rmvpe_public_api_matches_reference_semantics():
    sd, audio, gold = load_rmvpe_golden("r1_rmvpe_0p5s")
    model = model_for(0)
    mel = model.mel_extractor(torch.from_numpy(audio)[None].cuda(), center=True)
    assert tuple(mel.shape) == (1, 128, audio.shape[0] // 160 + 1) and mel.is_cuda
    assert mel_close(mel[0].cpu().numpy(), gold["mel"])
    hidden = model.mel2hidden(torch.from_numpy(gold["mel"])[None])                   # the reference's own mel as input
    assert tuple(hidden.shape) == (1, gold["hidden"].shape[0], 360)
    h = hidden[0].cpu().numpy()
    assert synthetic.snr_db(logit(gold["hidden"]), logit(h)) >= 40.0
    f0 = model.decode(gold["hidden"], thred=0.03)                                    # numpy in, numpy out like rmvpe.py:610-615
    assert np.allclose(f0, gold["f0"], rtol=1e-12, atol=0)
    cents = model.to_local_average_cents(gold["hidden"], thred=0.03)
    assert np.allclose(cents, rmvpe_oracle.to_local_average_cents(gold["hidden"].copy(), thred=0.03), rtol=1e-13, atol=0)
    f0c = model.infer_from_audio_with_pitch(audio, thred=0.03, f0_min=50, f0_max=1100)
    assert f0c.min() >= 50 and f0c.max() <= 1100
    a = model.infer_from_audio(audio)
    b = model.infer_from_audio(torch.from_numpy(audio))
    assert np.array_equal(a, b)                                                     # deterministic, tensor or numpy input


def test_rmvpe_decode_thresholds_ties_and_edges():
    rng = np.random.default_rng(5)
    T = 999
    sal = (rng.random((T, 360)) * 0.02).astype(np.float32)
    for t in range(T):
        if t % 5:
            c = int(rng.integers(0, 360))
            lo, hi = max(0, c - 3), min(360, c + 4)
            sal[t, lo:hi] += (rng.random(hi - lo) * 0.9).astype(np.float32)
    sal[1, 0], sal[2, 359], sal[3, 2], sal[4, 357] = 0.9, 0.95, 0.8, 0.85         # windows cut by the zero padding
    sal[6, 100] = sal[6, 200] = 0.7                                                # tie -> first maximum
    sal[7] = 0.03                                                                   # max == thred -> unvoiced
    model = model_for(0)
    want = rmvpe_oracle.decode(sal.copy(), thred=0.03)
    got = model.decode(sal, thred=0.03)
    assert (want == 0).sum() > 100 and ((got == 0) == (want == 0)).all()
    assert np.allclose(got, want, rtol=1e-12, atol=0)


def test_rmvpe_logmel_edges():
    """Shortest input the reference accepts (> 512 samples), lengths around a hop boundary, and a loud / silent signal."""
    model = model_for(0)
    rng = np.random.default_rng(2)
    for n in (513, 1599, 1600, 1601, 16000):
        x = (rng.standard_normal(n) * 0.3).astype(np.float32)
        if n == 1600:
            x[:] = 0                                                                # clamp floor: log(1e-5)
        ref = rmvpe_oracle.log_mel(torch.from_numpy(x)[None])[0].numpy()
        got = model.mel_extractor(torch.from_numpy(x)[None].cuda())[0].cpu().numpy()
        assert got.shape == ref.shape == (128, n // 160 + 1)
        assert mel_close(got, ref), (n, np.abs(got - ref).max(), np.abs(got - ref).mean())
    with pytest.raises(RuntimeError):
        model.mel_extractor(torch.zeros(1, 512).cuda())


@pytest.mark.parametrize("T", [1, 33, 256])
def test_rmvpe_gru_matches_oracle(T):
    lib = _lib.load()
    g = torch.Generator().manual_seed(T)
    w = {"weight_ih_l0": None}
    Hn = 256
    p = "g."
    sd = {}
    for sfx in ("", "_reverse"):
        sd[p + "weight_ih_l0" + sfx] = (torch.rand(768, 384, generator=g) * 2 - 1) / 16
        sd[p + "weight_hh_l0" + sfx] = (torch.rand(768, 256, generator=g) * 2 - 1) / 8
        sd[p + "bias_ih_l0" + sfx] = (torch.rand(768, generator=g) * 2 - 1) / 16
        sd[p + "bias_hh_l0" + sfx] = (torch.rand(768, generator=g) * 2 - 1) / 16
    x = torch.randn(1, T, 384, generator=g)
    want = rmvpe_oracle.bigru(x, sd, p)[0]
    gi = torch.cat([F.linear(x[0], sd[p + "weight_ih_l0"], sd[p + "bias_ih_l0"]),
                    F.linear(x[0], sd[p + "weight_ih_l0_reverse"], sd[p + "bias_ih_l0_reverse"])], dim=1).cuda().contiguous()
    whh = torch.stack([sd[p + "weight_hh_l0"], sd[p + "weight_hh_l0_reverse"]]).cuda().contiguous()
    bhh = torch.stack([sd[p + "bias_hh_l0"], sd[p + "bias_hh_l0_reverse"]]).cuda().contiguous()
    o16 = torch.empty(T, 512, dtype=torch.float16, device="cuda")
    o32 = torch.empty(T, 512, dtype=torch.float32, device="cuda")
    st = lib.rvcb200_op_rmvpe_gru(C.c_void_p(gi.data_ptr()), C.c_void_p(whh.data_ptr()), C.c_void_p(bhh.data_ptr()),
                                  C.c_void_p(o16.data_ptr()), C.c_void_p(o32.data_ptr()), T, stream())
    assert st == 0
    torch.cuda.synchronize()
    err = (o32.cpu() - want).abs().max().item()
    print(f"GRU T={T}: max |err| {err:.2e}")
    assert err <= 2e-5
    assert (o16.float().cpu() - want).abs().max().item() <= 1e-3


@pytest.mark.parametrize("H,W,cin,cout,taps,a_mode", [
    (8, 128, 16, 16, 9, 2), (8, 128, 16, 16, 9, 1), (64, 128, 8, 16, 9, 2), (32, 64, 32, 32, 9, 2), (32, 64, 32, 32, 9, 1),
    (700, 128, 16, 16, 9, 2),                                         # > 2 waves of tiles: resident weights + row slabs
    (16, 16, 128, 128, 9, 0), (16, 16, 128, 128, 9, 2), (6, 4, 256, 512, 9, 0), (6, 4, 256, 512, 9, 2),
    (12, 8, 512, 1024, 4, 0), (12, 8, 512, 1024, 4, 2), (64, 128, 32, 16, 1, 0)])
def test_rmvpe_image_convolution_matches_torch(H, W, cin, cout, taps, a_mode):
    """The generic tcgen05 kernel with 2-D taps -- one activation box per k-block (a_mode 0, narrow images), per tap (1) or per
    kernel row (2, what the model uses for wide images) -- against F.conv2d / F.conv_transpose2d on the same fp16-rounded operands."""
    lib = _lib.load()
    g = torch.Generator().manual_seed(H * 1000 + W + cin)
    Wp, rows = W + 1, H * (W + 1)
    img = torch.zeros(H, Wp, cin)
    img[:, :W] = torch.randn(H, W, cin, generator=g)
    x16 = img.reshape(rows, cin).half().cuda().contiguous()
    xin = x16.float().cpu().reshape(H, Wp, cin)[:, :W].permute(2, 0, 1)[None]                       # [1][cin][H][W]
    bias = torch.randn(cout, generator=g)
    res = torch.randn(rows, cout, generator=g)
    if taps == 9:
        w = (torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5).half().float()
        t = w.permute(2, 3, 1, 0).reshape(9, cin, cout)
        ref = F.relu(F.conv2d(xin, w, bias, padding=1))
    elif taps == 1:
        w = (torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5).half().float()
        t = w[:, :, 0, 0].t()[None]
        ref = F.relu(F.conv2d(xin, w, bias))
    else:
        co = cout // 4
        wt = (torch.randn(cin, co, 3, 3, generator=g) / (cin * 2.25) ** 0.5).half().float()
        t = torch.zeros(4, cin, cout)
        for dy in range(2):
            for dx in range(2):
                for py in range(2):
                    for px in range(2):
                        ky, kx = py + 1 - 2 * dy, px + 1 - 2 * dx
                        if 0 <= ky <= 2 and 0 <= kx <= 2:
                            t[dy * 2 + dx, :, (py * 2 + px) * co:(py * 2 + px + 1) * co] = wt[:, :, ky, kx]
        bias = bias[:co].repeat(4)
        ref = F.relu(F.conv_transpose2d(xin, wt, bias[:co], stride=2, padding=1, output_padding=1))
    n_tile = RMVPE._n_tile(rows, cout)
    w16 = pack_tc(t.contiguous(), torch.float16, n_tile).cuda()
    bias_d, res_d = bias.cuda(), res.cuda()
    y32 = torch.full((rows, cout), 7.0, device="cuda")
    y16 = torch.full((rows, cout), 7.0, device="cuda", dtype=torch.float16)
    d = _lib.TcConvDesc()
    d.x16, d.L_in, d.padf = x16.data_ptr(), rows, 32
    d.w16, d.bias = w16.data_ptr(), bias_d.data_ptr()
    d.Cin, d.ntaps, d.dil, d.G = cin, taps, 1, 1
    if taps == 9:
        d.tap_w, d.dil2, d.g_off[0] = 3, Wp, -(Wp + 1)
    elif taps == 4:
        d.tap_w, d.dil2 = 2, Wp
    d.a_mode = a_mode
    d.N, d.Cout_total = n_tile, cout
    d.Lj, d.out_stride, d.Lp_out = rows, 1, ((rows + 127) // 128) * 128 + 128
    d.div, d.out_slope, d.alpha, d.pre_slope = 1.0, 1.0, 1.0, 0.0
    d.generic, d.f32_cl, d.ldx16 = 1, 1, cin
    d.y32, d.ldy32, d.y16, d.ldy16 = y32.data_ptr(), cout, y16.data_ptr(), cout
    if taps != 4:
        d.pad_period, d.pad_valid, d.mask_post = Wp, W, 1
        d.res32, d.ldr32, d.res_mode = res_d.data_ptr(), cout, 1
    assert lib.rvcb200_op_conv_tc(C.byref(d), 1, stream()) == 0
    torch.cuda.synchronize()
    got = y32.cpu().reshape(H, Wp, cout)
    if taps == 4:
        co = cout // 4
        up = torch.zeros(2 * H, 2 * W, co)
        for ph in range(4):
            up[(ph >> 1)::2, (ph & 1)::2] = got[:, :W, ph * co:(ph + 1) * co]
        want = ref[0].permute(1, 2, 0)
        err = (up - want).abs().max().item()
        # the shuffle kernel scatters the same values into the decoder's concat buffer
        cat = torch.zeros(2 * H * (2 * W + 1), 2 * co, dtype=torch.float16, device="cuda")
        assert lib.rvcb200_op_rmvpe_shuffle(C.c_void_p(y16.data_ptr()), C.c_void_p(cat.data_ptr()), H, W, co, 2 * co, Wp, 2 * W + 1, 1,
                                            stream()) == 0
        torch.cuda.synchronize()
        c = cat.float().cpu().reshape(2 * H, 2 * W + 1, 2 * co)
        assert (c[:, :2 * W, :co] - up.half().float()).abs().max().item() == 0 and c[:, 2 * W].abs().max() == 0 and c[:, :, co:].abs().max() == 0
        # ... and with `pack` pixels per row (the wide levels): row = [pack x co up-sampled | pack x co skip], W / pack + 1 rows per line
        for pack in (2, 4):
            fp = 2 * W // pack + 1
            catp = torch.zeros(2 * H * fp, 2 * pack * co, dtype=torch.float16, device="cuda")
            assert lib.rvcb200_op_rmvpe_shuffle(C.c_void_p(y16.data_ptr()), C.c_void_p(catp.data_ptr()), H, W, co, 2 * pack * co, Wp, fp,
                                                pack, stream()) == 0
            torch.cuda.synchronize()
            cp = catp.float().cpu().reshape(2 * H, fp, 2 * pack * co)
            got_up = cp[:, :2 * W // pack, :pack * co].reshape(2 * H, 2 * W, co)
            assert (got_up - up.half().float()).abs().max().item() == 0 and cp[:, -1].abs().max() == 0 and cp[:, :, pack * co:].abs().max() == 0
    else:
        want = ref[0].permute(1, 2, 0) + res.reshape(H, Wp, cout)[:, :W]
        err = (got[:, :W] - want).abs().max().item()
        assert got[:, W].abs().max().item() == 0 and y16.float().cpu().reshape(H, Wp, cout)[:, W].abs().max().item() == 0
        assert (y16.float().cpu() - y32.cpu()).abs().max().item() <= 2e-3 * max(1.0, y32.abs().max().item())
    print(f"H={H} W={W} {cin}->{cout} taps={taps} a_mode={d.a_mode} N={n_tile}: max |err| {err:.2e}")
    assert err <= 2e-3


def test_rmvpe_pool_matches_torch():
    lib = _lib.load()
    H, W, Cn = 32, 64, 32
    g = torch.Generator().manual_seed(9)
    x = torch.randn(H, W + 1, Cn, generator=g)
    xd = x.cuda().contiguous()
    want = F.avg_pool2d(x[:, :W].permute(2, 0, 1)[None], 2)[0].permute(1, 2, 0)
    for p_out in (W // 2 + 1, W // 2 + 2, W // 2 + 4):                    # output line pitch: 1, 2 or 4 zero pixels behind the data
        y = torch.full((H // 2, p_out, Cn), 5.0, dtype=torch.float16, device="cuda")
        assert lib.rvcb200_op_rmvpe_pool(C.c_void_p(xd.data_ptr()), Cn, C.c_void_p(y.data_ptr()), H // 2, W // 2, Cn, W + 1, p_out,
                                         stream()) == 0
        torch.cuda.synchronize()
        got = y.float().cpu()
        assert (got[:, :W // 2] - want.half().float()).abs().max().item() <= 1e-3 and got[:, W // 2:].abs().max().item() == 0


def test_rmvpe_feeds_the_pitch_pipeline():
    """`FeatureExtractor.get_f0(..., f0_method="rmvpe")` (pitch_extraction.py:191-195, 252-304) with the B200 model attached."""
    from comfy_rvc_b200 import pipeline as pl
    vc = pl.VC(48000, pl.PipelineConfig(1, 6, 38, 41, is_half=True, device="cuda:0"))
    vc.model_rmvpe = model_for(0)
    audio = synthetic.make_speech(2.0, seed=21)[0].numpy()
    coarse, f0 = vc.get_f0(audio, 0, "rmvpe")
    assert coarse.dtype == np.int16 and coarse.shape == f0.shape == (audio.shape[0] // 160 + 1,)
    assert coarse.min() >= 1 and coarse.max() <= 255
    coarse2, f02 = vc.get_f0(audio, 12, "rmvpe+", f0_min=50, f0_max=1100)
    assert f02.min() >= 50 * 2 - 1e-9 and f02.max() <= 1100 * 2 + 1e-9
