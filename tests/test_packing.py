"""CPU: the host-side weight packing (fold, flip absorption, gate interleave, transposed-conv
phase decomposition, stacked conditioning) is equivalent to the oracle's formulation."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from comfy_rvc_b200 import synthetic, weights
from comfy_rvc_b200.config import NAMED_CONFIGS
from oracle import rvc_oracle
from tests._emulate import conv_cl


@pytest.mark.parametrize("k,u", [(24, 12), (20, 10), (16, 10), (16, 8), (4, 2), (16, 6), (16, 4)])
def test_conv_transpose_phase_decomposition(k, u):
    g = torch.Generator().manual_seed(k * 100 + u)
    cin, cout, L = 16, 8, 37
    w = torch.randn(cin, cout, k, generator=g, dtype=torch.float64)
    b = torch.randn(cout, generator=g, dtype=torch.float64)
    x = torch.randn(2, cin, L, generator=g, dtype=torch.float64)
    ref = F.conv_transpose1d(x, w, b, stride=u, padding=(k - u) // 2)         # [B][cout][L*u]
    assert ref.shape[-1] == L * u
    pad, ntaps, g_off = weights.up_geometry(k, u)
    y = conv_cl(x.transpose(1, 2), weights.pack_conv_transpose(w, u), b, g_off=g_off, out_stride=u)
    torch.testing.assert_close(y.transpose(1, 2), ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("cfg_name", ["48k_v2", "40k"])
def test_flow_packing_matches_oracle(cfg_name):
    cfg = NAMED_CONFIGS[cfg_name]
    sd = {k: v.double() for k, v in synthetic.make_state_dict(cfg).items()}
    P, S = weights.pack(cfg, sd)
    P = {k: v.double() for k, v in P.items()}
    w = {k: v.double() for k, v in rvc_oracle.fold_weight_norm(sd).items()}
    B, T, H, C = 2, 40, cfg.hidden_channels, cfg.inter_channels
    half = C // 2
    g = torch.Generator().manual_seed(3)
    lens = torch.tensor([T, 29])
    mask = (torch.arange(T)[None] < lens[:, None]).double()                   # [B][T]
    z_p = torch.randn(B, C, T, generator=g, dtype=torch.float64) * mask[:, None]
    sid = torch.tensor([0, 5])
    gvec = w["emb_g.weight"][sid]                                             # [B][gin]
    ref = rvc_oracle.flow_reverse(w, cfg, z_p, mask[:, None], gvec.unsqueeze(-1))
    # --- emulate engine.cu's flow section with the packed tensors ---
    cond = gvec @ P["cond.w"].t() + P["cond.b"]                               # [B][n_cond]
    z = z_p.transpose(1, 2).clone()                                           # channels-last
    m3 = mask[:, :, None]
    flipped = False
    for i in reversed(range(cfg.n_flows)):
        flipped = not flipped
        in_off = half if flipped else 0
        out_off = half - in_off
        h = conv_cl(z[:, :, in_off:in_off + half], P[f"flow.{i}.pre.w"][None, None], P[f"flow.{i}.pre.b"]) * m3
        skip = None
        for j in range(cfg.flow_wn_layers):
            kf = cfg.flow_kernel
            xin = conv_cl(h, P[f"flow.{i}.in.{j}.w"][None], P[f"flow.{i}.in.{j}.b"], g_off=[-(kf - 1) // 2])
            # the tensor path's folded form (engine.cu): the same in_layer over [x0 m | m | 0.. | acts_0 m | ..] -- `pre` and the res
            # convolutions live in the weights, h is never formed
            xb = weights.flow_x0_block(cfg)
            if j == 0:
                x0 = z[:, :, in_off:in_off + half] * m3
                fold_in = torch.cat([x0, m3.expand(-1, -1, 1), x0.new_zeros(B, T, xb - half - 1)], dim=-1)
            assert P[f"flow.{i}.inf.{j}.w"].shape == (kf, xb + j * H, 2 * H)
            xin_fold = conv_cl(fold_in, P[f"flow.{i}.inf.{j}.w"][None], P[f"flow.{i}.in.{j}.b"], g_off=[-(kf - 1) // 2])
            torch.testing.assert_close(xin_fold, xin, rtol=0, atol=5e-6)      # folded weights are stored in float32
            off = cfg.upsample_initial_channel + (i * cfg.flow_wn_layers + j) * 2 * H
            xin = xin + cond[:, None, off:off + 2 * H]
            acts = torch.tanh(xin[..., 0::2]) * torch.sigmoid(xin[..., 1::2])
            if j < cfg.flow_wn_layers - 1:
                h = (h + conv_cl(acts, P[f"flow.{i}.rs.{j}.res.w"][None, None], P[f"flow.{i}.rs.{j}.res.b"])) * m3
            sk = conv_cl(acts, P[f"flow.{i}.rs.{j}.skip.w"][None, None], P[f"flow.{i}.rs.{j}.skip.b"])
            skip = sk if skip is None else skip + sk
            acts_all = acts if j == 0 else torch.cat([acts_all, acts], dim=-1)
            fold_in = torch.cat([fold_in, acts * m3], dim=-1)
        m = conv_cl(skip, P[f"flow.{i}.post.w"][None, None], P[f"flow.{i}.post.b"], in_len=lens) * m3
        # the tensor path's folded form (engine.cu "flow.skip+post"): one contraction of [acts_0 | acts_1 | ...] with W_skip_j W_post
        m_fold = conv_cl(acts_all, P[f"flow.{i}.sp.w"][None, None], P[f"flow.{i}.sp.b"], in_len=lens) * m3
        assert P[f"flow.{i}.sp.w"].shape == (cfg.flow_wn_layers * H, half)
        torch.testing.assert_close(m_fold, m, rtol=0, atol=2e-7)          # the folded weights are stored in float32
        z[:, :, out_off:out_off + half] = (z[:, :, out_off:out_off + half] - m) * m3
    torch.testing.assert_close(z.transpose(1, 2), ref, rtol=1e-10, atol=1e-10)
    # dec.cond is the first block of the stacked conditioning matrix
    ref_c = F.conv1d(gvec.unsqueeze(-1), w["dec.cond.weight"], w["dec.cond.bias"])[:, :, 0]
    torch.testing.assert_close(cond[:, :cfg.upsample_initial_channel], ref_c, rtol=1e-12, atol=1e-12)


def test_encoder_and_decoder_packing_shapes():
    cfg = NAMED_CONFIGS["48k"]                                                # 5-stage ladder
    sd = synthetic.make_state_dict(cfg)
    P, S = weights.pack(cfg, sd)
    w = rvc_oracle.fold_weight_norm(sd)
    x = torch.randn(1, 11, 192)
    # qkv fused 1x1 == three separate convs
    qkv = conv_cl(x, P["enc.0.qkv.w"][None, None], P["enc.0.qkv.b"])
    a = "enc_p.encoder.attn_layers.0"
    for n, name in enumerate("qkv"):
        ref = F.conv1d(x.transpose(1, 2), w[f"{a}.conv_{name}.weight"], w[f"{a}.conv_{name}.bias"]).transpose(1, 2)
        torch.testing.assert_close(qkv[..., n * 192:(n + 1) * 192], ref, rtol=1e-5, atol=1e-5)
    # dilated resblock conv through the generic contract
    xx = torch.randn(1, 50, 256)
    k, d = 11, 5
    ref = F.conv1d(xx.transpose(1, 2), w["dec.resblocks.2.convs1.2.weight"], w["dec.resblocks.2.convs1.2.bias"],
                   dilation=d, padding=(k * d - d) // 2).transpose(1, 2)
    y = conv_cl(xx, P["dec.rb.2.c1.2.w"][None], P["dec.rb.2.c1.2.b"], g_off=[-((k - 1) // 2) * d], dil=d)
    torch.testing.assert_close(y, ref, rtol=1e-4, atol=1e-4)
    assert P["dec.post.w"].shape == (7, 16) and P["dec.noise.0.w"].shape == (96, 256)
    missing, unexpected, mismatched = weights.validate_state_dict(cfg, sd)
    assert not missing and not unexpected and not mismatched


def test_dense_transposed_conv_packing_matches_torch():
    """Stride-2 transposed convs run as ONE ordinary 3-tap conv C_in -> 2*C_out with the phases side by side
    (weights.pack_conv_transpose_dense): check against F.conv_transpose1d (models.py:498-511 geometry)."""
    import torch.nn.functional as F
    from comfy_rvc_b200 import weights
    g = torch.Generator().manual_seed(3)
    cin, cout, u, k, L = 8, 4, 2, 4, 11
    w = torch.randn(cin, cout, k, generator=g)
    x = torch.randn(2, cin, L, generator=g)
    ref = F.conv_transpose1d(x, w, stride=u, padding=(k - u) // 2)
    w3 = weights.pack_conv_transpose_dense(w, u)
    xp = F.pad(x, (1, 1)).transpose(1, 2)
    out = sum(xp[:, tau:tau + L] @ w3[tau] for tau in range(3))
    out = out.reshape(2, L * u, cout).transpose(1, 2)
    assert torch.allclose(out, ref, atol=1e-6)
    from comfy_rvc_b200.config import NAMED_CONFIGS
    assert [weights.ups_is_dense(NAMED_CONFIGS["48k_v2"], i) for i in range(4)] == [False, False, True, True]
    assert [weights.ups_is_dense(NAMED_CONFIGS["48k"], i) for i in range(5)] == [False, False, True, True, True]


def test_conv_post_tensor_core_image():
    """conv_post as a C -> C convolution for the tcgen05 resblock kernel: output channel 0 carries the k x C taps, every
    other output channel and the bias are zero; only where C_last is one of the kernel's channel counts."""
    for name, cfg in NAMED_CONFIGS.items():
        c_last = cfg.upsample_initial_channel >> cfg.num_upsamples
        assert weights.post_is_tc(cfg) == (c_last in (32, 64, 128) and cfg.resblock == "1")
        P, _ = weights.pack(cfg, synthetic.make_state_dict(cfg))
        if weights.post_is_tc(cfg):
            wt = P["dec.post.wt"]
            assert tuple(wt.shape) == (7, c_last, c_last)
            assert torch.equal(wt[:, :, 0], P["dec.post.w"]) and float(wt[:, :, 1:].abs().max()) == 0.0
            assert float(P["dec.post.bt"].abs().max()) == 0.0 and P["dec.post.bt"].numel() == c_last
            assert "dec.post.wt" in weights.tc_weight_names(cfg)
        else:
            assert "dec.post.wt" not in P and "dec.post.wt" not in weights.tc_weight_names(cfg)
