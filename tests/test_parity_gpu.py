"""GPU: end-to-end parity of the CUDA `infer()` (through the drop-in class and the C ABI) against
(a) the reference-minted golden fixtures and (b) the CPU oracle, stage by stage.

Gate (BASELINE.json north_star): fp32 path int16 PCM within +-1 LSB of the reference after the
reference's own peak-normalising conversion (vc_infer_pipeline.py:188-189)."""
import numpy as np
import pytest
import torch

import comfy_rvc_b200 as rvc
from comfy_rvc_b200 import synthetic
from comfy_rvc_b200.config import NAMED_CONFIGS
from oracle import rvc_oracle
from tests._util import GOLDEN_CASES, int16_lsb_diff, load_golden, net_infer, oracle_infer

pytestmark = pytest.mark.gpu


def build_net(cfg, sd, precision="fp32"):
    if cfg.f0:
        cls = rvc.SynthesizerTrnMs256NSFsid if cfg.feat_dim == 256 else rvc.SynthesizerTrnMs768NSFsid
    else:
        cls = rvc.SynthesizerTrnMs256NSFsid_nono if cfg.feat_dim == 256 else rvc.SynthesizerTrnMs768NSFsid_nono
    net = cls(*cfg.to_positional(), is_half=False)
    del net.enc_q
    net.load_state_dict({k: v.half() for k, v in sd.items()}, strict=False)   # checkpoints are fp16 on disk
    net.eval().to("cuda:0")
    return net.float().set_precision(precision)


def stage_report(taps_gpu, taps_ref, names):
    rows = []
    for n in names:
        a = taps_gpu[n].cpu().double()
        r = taps_ref[n].double()
        if r.dim() == 3 and a.dim() == 3 and r.shape != a.shape:
            r = r.transpose(1, 2)                                       # oracle is channels-first
        if r.dim() == 3 and a.dim() == 2:
            r = r[:, 0]
        err = (a - r).abs().max().item()
        rows.append((n, err, r.abs().max().item()))
    return rows


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_infer_matches_golden_fp32(name):
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    cfg, sd, (phone, lens, pitch, pitchf, sid), noise, gold = load_golden(name)
    net = build_net(cfg, sd)
    tap_names = ["x_enc", "stats", "z_p", "z", "har_source", "dec.pre"] + \
                [f"dec.ups.{i}" for i in range(cfg.num_upsamples)] + [f"dec.stage.{i}" for i in range(cfg.num_upsamples)]
    taps = {n: None for n in tap_names}
    o, x_mask, (z, z_p, m_p, logs_p) = net_infer(net, cfg, (phone, lens, pitch, pitchf, sid), noise, taps)
    torch.cuda.synchronize()
    assert net.last_launches > 0
    # oracle with stage taps (CPU) for a readable per-stage report
    w = rvc_oracle.fold_weight_norm(sd)
    ref_taps = {}
    oracle_infer(cfg, w, (phone, lens, pitch, pitchf, sid), noise, taps=ref_taps)
    ref_taps["stats"] = torch.cat([ref_taps["m_p"], ref_taps["logs_p"]], dim=1)
    for i in range(cfg.num_upsamples):
        if f"dec.ups_plus_noise.{i}" in ref_taps:
            ref_taps[f"dec.ups.{i}"] = ref_taps[f"dec.ups_plus_noise.{i}"]
    names = [n for n in tap_names if n in ref_taps]
    T = phone.shape[1]
    valid_only = bool((lens < T).any())
    for n, err, mag in stage_report(taps, ref_taps, names):
        print(f"  {name} {n:14s} max|err| {err:.3e}  (|ref|max {mag:.3f})")
        if not valid_only or n in ("stats", "z_p", "z", "har_source"):
            assert err < 5e-4 * max(mag, 1.0), n
    assert np.array_equal(x_mask.cpu().numpy(), gold["x_mask"])
    for got, key in ((m_p, "m_p"), (logs_p, "logs_p"), (z_p, "z_p"), (z, "z")):
        assert got.shape == gold[key].shape
        np.testing.assert_allclose(got.cpu().numpy(), gold[key], rtol=0, atol=1e-4, err_msg=key)
    o_np = o[:, 0].cpu().numpy()
    assert o_np.shape == gold["o_f32"].shape
    worst = 0
    for b in range(o_np.shape[0]):
        n = int(lens[b]) * cfg.upp
        if int(lens[b]) < T:
            n -= 12 * cfg.upp          # H7: the decoder is unmasked; the last ~10 frames of a short item see padding
            continue_cmp = n > 0
        else:
            continue_cmp = True
        if continue_cmp:
            # compare with a shared peak so that truncated spans normalise identically
            ref_seg, got_seg = gold["o_f32"][b, :n], o_np[b, :n]
            peak = np.abs(gold["o_f32"][b, : int(lens[b]) * cfg.upp]).max() / 0.99
            d = np.abs((ref_seg * 32768 / peak).astype(np.int16).astype(np.int32)
                       - (got_seg * 32768 / peak).astype(np.int16).astype(np.int32)).max()
            worst = max(worst, int(d))
            print(f"  {name} item {b}: float max|err| {np.abs(ref_seg - got_seg).max():.3e}, int16 LSB diff {d}")
    assert worst <= 1, f"int16 PCM differs by {worst} LSB"
    if not valid_only:
        assert int16_lsb_diff(gold["o_f32"][0], o_np[0]) <= 1


def test_batch_items_are_independent_and_deterministic():
    cfg = NAMED_CONFIGS["32k_v2"]
    sd = synthetic.make_state_dict(cfg)
    net = build_net(cfg, sd)
    T = 80
    phone, lens, pitch, pitchf, sid = synthetic.make_inputs(cfg, 3, T, seed=5)
    noise = synthetic.draw_noise(cfg, 3, T, seed=6)
    o3 = net.infer(phone.cuda(), lens.cuda(), pitch.cuda(), pitchf.cuda(), sid.cuda(), noise=noise)[0]
    o3b = net.infer(phone.cuda(), lens.cuda(), pitch.cuda(), pitchf.cuda(), sid.cuda(), noise=noise)[0]
    assert torch.equal(o3, o3b)                                         # bitwise deterministic
    for b in range(3):
        nb = tuple(t[b:b + 1] for t in noise)
        o1 = net.infer(phone[b:b + 1].cuda(), lens[b:b + 1].cuda(), pitch[b:b + 1].cuda(), pitchf[b:b + 1].cuda(),
                       sid[b:b + 1].cuda(), noise=nb)[0]
        assert torch.equal(o1[0], o3[b])                                # batching does not change an item


def test_no_fallback_without_weights():
    cfg = NAMED_CONFIGS["40k"]
    net = rvc.SynthesizerTrnMs256NSFsid(*cfg.to_positional(), is_half=False)
    with pytest.raises(RuntimeError):
        net.infer(*[t.cuda() for t in synthetic.make_inputs(cfg, 1, 8)])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two CUDA devices in one process")
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_same_process_two_devices(precision):
    """The per-kernel shared-memory opt-in and the SM count are per-device state (csrc/common.cuh opt_in_smem):
    a process that synthesises on cuda:0 and then on cuda:1 must get the same waveform on both."""
    cfg, sd, inputs, noise, gold = load_golden("c2_48k_v2")
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        net = build_net(cfg, sd, precision).to(dev).set_precision(precision)
        o = net_infer(net, cfg, inputs, noise, device=dev)[0]
        torch.cuda.synchronize(dev)
        outs.append(o.cpu())
    assert torch.equal(outs[0], outs[1])
    assert int16_lsb_diff(gold["o_f32"][0], outs[1][0, 0].numpy()) <= (1 if precision == "fp32" else 400)


@pytest.mark.parametrize("name,rate", [("c2_48k_v2", 0.4), ("c3_32k_v2_ragged", 0.75), ("c7_40k_v1_nono", 0.5), ("c1_40k_v1", 0.001)])
def test_infer_rate_matches_oracle(name, rate):
    """`infer(..., rate=r)` (models.py:802-806): encoder and prior over the whole input, flow + decoder over the last
    int(T * r) frames; r so small that head = 0 means the whole tensor (`z_p[:, :, -0:]`)."""
    cfg, sd, (phone, lens, pitch, pitchf, sid), noise, gold = load_golden(name)
    net = build_net(cfg, sd)
    w = rvc_oracle.fold_weight_norm(sd)
    ins = [t.cuda() for t in (phone, lens, pitch, pitchf, sid)]
    if cfg.f0:
        ref = rvc_oracle.infer(w, cfg, phone, lens, pitch, pitchf, sid, *noise, rate=rate)
        got = net.infer(*ins, rate=rate, noise=noise)
    else:
        ref = rvc_oracle.infer_nono(w, cfg, phone, lens, sid, noise[0], rate=rate)
        got = net.infer(ins[0], ins[1], ins[4], rate=rate, noise=noise[:1])
    torch.cuda.synchronize()
    T = phone.shape[1]
    head = int(T * rate) or T
    assert got[0].shape == ref[0].shape == (phone.shape[0], 1, head * cfg.upp)
    assert torch.equal(got[1].cpu(), ref[1]) and got[2][0].shape == ref[2][0].shape == got[2][1].shape
    for g, r in zip(got[2], ref[2]):
        np.testing.assert_allclose(g.cpu().numpy(), r.numpy(), rtol=0, atol=1e-4)
    for b in range(phone.shape[0]):
        n_valid = max(int(lens[b]) - (T - head), 0)
        n = (n_valid - 12) * cfg.upp if n_valid < head else head * cfg.upp
        if n > 0:
            a, e = ref[0][b, 0, :n].numpy(), got[0][b, 0, :n].cpu().numpy()
            peak = np.abs(ref[0][b, 0, : max(n_valid, 1) * cfg.upp].numpy()).max() / 0.99
            d = np.abs((a * 32768 / peak).astype(np.int16).astype(np.int32) - (e * 32768 / peak).astype(np.int16).astype(np.int32)).max()
            assert d <= 1, (b, int(d))
