"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A CPU, fp32, functional restatement of the reference synthesis hot path
(`SynthesizerTrnMs{256,768}NSFsid.infer`, /root/reference/lib/infer_pack/models.py:682-693 and
:798-809) written against a plain state_dict, with the three RNG draws made explicit inputs.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import this module; the CUDA product (`comfy_rvc_b200`) never does.

Parity pin: the reference has no tests or golden vectors for this path, so
this oracle is pinned against outputs of the reference itself, imported read-only in the build
container by `tests/golden/make_golden.py`; the resulting fixtures live in `tests/golden/*.npz`
and `tests/test_oracle_golden.py` replays them (no /root/reference needed at test time).

Each function cites the reference lines it restates.  The attention uses the banded
relative-position form instead of the reference's pad/reshape skewing; the
sine source follows the bit-level recipe of SURVEY.md App. C by simply calling the same torch
CPU ops in the same order.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # modules.py:13


# ------------------------------------------------------------------------------------------
# weights
# ------------------------------------------------------------------------------------------
def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """`weight_g`/`weight_v` -> `weight` with the same primitive the reference's hook uses
    (torch.nn.utils.weight_norm, dim=0 -> torch._weight_norm(v, g, 0)); SURVEY App. B."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if k.endswith("weight_v"):
            g = sd[k[:-1] + "g"]
            out[k[:-8] + "weight"] = torch._weight_norm(v.float(), g.float(), 0)
        elif k.endswith("weight_g"):
            continue
        else:
            out[k] = v.float()
    return out


# ------------------------------------------------------------------------------------------
# TextEncoder (models.py:43-58 / 90-105) and attentions.Encoder (attentions.py:57-69)
# ------------------------------------------------------------------------------------------
def _layer_norm_c(x, gamma, beta, eps=1e-5):
    """modules.py:25-28 — LayerNorm over the channel axis of [B,C,T]."""
    return F.layer_norm(x.transpose(1, -1), (x.shape[1],), gamma, beta, eps).transpose(1, -1)


def _attention(w, pfx, x, attn_mask, n_heads, window):
    """attentions.py:212-270 in banded form (App. D)."""
    B, C, T = x.shape
    dk = C // n_heads
    q = F.conv1d(x, w[pfx + ".conv_q.weight"], w[pfx + ".conv_q.bias"])
    k = F.conv1d(x, w[pfx + ".conv_k.weight"], w[pfx + ".conv_k.bias"])
    v = F.conv1d(x, w[pfx + ".conv_v.weight"], w[pfx + ".conv_v.bias"])
    q = q.view(B, n_heads, dk, T).transpose(2, 3)
    k = k.view(B, n_heads, dk, T).transpose(2, 3)
    v = v.view(B, n_heads, dk, T).transpose(2, 3)
    qs = q / math.sqrt(dk)                                     # :229, :235-237
    scores = torch.matmul(qs, k.transpose(-2, -1))             # [B,h,T,T]
    rel_k = w[pfx + ".emb_rel_k"]                              # [1, 2w+1, dk] (heads_share)
    rel_v = w[pfx + ".emb_rel_v"]
    rel_logits = torch.matmul(qs, rel_k.unsqueeze(0).transpose(-2, -1))   # [B,h,T,2w+1]
    idx = torch.arange(T, device=x.device)
    for r in range(2 * window + 1):
        off = r - window                                        # key j = i + off
        i0, i1 = max(0, -off), min(T, T - off)
        if i1 <= i0:
            continue
        ii = idx[i0:i1]
        scores[:, :, ii, ii + off] += rel_logits[:, :, i0:i1, r]
    scores = scores.masked_fill(attn_mask == 0, -1e4)          # :245-246
    p = F.softmax(scores, dim=-1)
    out = torch.matmul(p, v)
    band = torch.zeros(B, n_heads, T, 2 * window + 1, dtype=p.dtype, device=p.device)
    for r in range(2 * window + 1):
        off = r - window
        i0, i1 = max(0, -off), min(T, T - off)
        if i1 <= i0:
            continue
        ii = idx[i0:i1]
        band[:, :, i0:i1, r] = p[:, :, ii, ii + off]
    out = out + torch.matmul(band, rel_v.unsqueeze(0))         # :260-267
    out = out.transpose(2, 3).contiguous().view(B, C, T)
    return F.conv1d(out, w[pfx + ".conv_o.weight"], w[pfx + ".conv_o.bias"])


def _ffn(w, pfx, x, x_mask, ks):
    """attentions.py:387-413 (same padding, ReLU)."""
    pl, pr = (ks - 1) // 2, ks // 2
    h = F.conv1d(F.pad(x * x_mask, (pl, pr)), w[pfx + ".conv_1.weight"], w[pfx + ".conv_1.bias"])
    h = torch.relu(h)
    h = F.conv1d(F.pad(h * x_mask, (pl, pr)), w[pfx + ".conv_2.weight"], w[pfx + ".conv_2.bias"])
    return h * x_mask


def text_encoder(w, cfg, phone, pitch, lengths):
    H = cfg.hidden_channels
    x = F.linear(phone, w["enc_p.emb_phone.weight"], w["enc_p.emb_phone.bias"])
    if pitch is not None:
        x = x + F.embedding(pitch, w["enc_p.emb_pitch.weight"])
    x = x * math.sqrt(H)
    x = F.leaky_relu(x, 0.1)
    x = x.transpose(1, -1)
    T = x.shape[2]
    x_mask = (torch.arange(T, device=x.device).unsqueeze(0) < lengths.unsqueeze(1)).unsqueeze(1).to(x.dtype)  # commons.py:232-236
    attn_mask = x_mask.unsqueeze(2) * x_mask.unsqueeze(-1)
    x = x * x_mask
    x = x * x_mask
    for l in range(cfg.n_layers):
        y = _attention(w, f"enc_p.encoder.attn_layers.{l}", x, attn_mask, cfg.n_heads, cfg.window_size)
        x = _layer_norm_c(x + y, w[f"enc_p.encoder.norm_layers_1.{l}.gamma"], w[f"enc_p.encoder.norm_layers_1.{l}.beta"])
        y = _ffn(w, f"enc_p.encoder.ffn_layers.{l}", x, x_mask, cfg.kernel_size)
        x = _layer_norm_c(x + y, w[f"enc_p.encoder.norm_layers_2.{l}.gamma"], w[f"enc_p.encoder.norm_layers_2.{l}.beta"])
    x = x * x_mask
    stats = F.conv1d(x, w["enc_p.proj.weight"], w["enc_p.proj.bias"]) * x_mask
    m, logs = torch.split(stats, cfg.inter_channels, dim=1)
    return m, logs, x_mask


# ------------------------------------------------------------------------------------------
# reverse flow (models.py:185-192; modules.py:436-455, 184-209, 373-380)
# ------------------------------------------------------------------------------------------
def _wn(w, pfx, x, x_mask, g, H, n_layers, ks):
    out = torch.zeros_like(x)
    gc = F.conv1d(g, w[pfx + ".cond_layer.weight"], w[pfx + ".cond_layer.bias"])
    for i in range(n_layers):
        x_in = F.conv1d(x, w[f"{pfx}.in_layers.{i}.weight"], w[f"{pfx}.in_layers.{i}.bias"], padding=(ks - 1) // 2)
        a = x_in + gc[:, i * 2 * H:(i + 1) * 2 * H, :]
        acts = torch.tanh(a[:, :H]) * torch.sigmoid(a[:, H:])       # commons.py:211-218
        rs = F.conv1d(acts, w[f"{pfx}.res_skip_layers.{i}.weight"], w[f"{pfx}.res_skip_layers.{i}.bias"])
        if i < n_layers - 1:
            x = (x + rs[:, :H]) * x_mask
            out = out + rs[:, H:]
        else:
            out = out + rs
    return out * x_mask


def flow_reverse(w, cfg, z_p, x_mask, g):
    x = z_p
    half = cfg.inter_channels // 2
    for i in reversed(range(cfg.n_flows)):
        x = torch.flip(x, [1])                                       # Flip comes after RCL_i in forward order
        pfx = f"flow.flows.{2 * i}"
        x0, x1 = x[:, :half], x[:, half:]
        h = F.conv1d(x0, w[pfx + ".pre.weight"], w[pfx + ".pre.bias"]) * x_mask
        h = _wn(w, pfx + ".enc", h, x_mask, g, cfg.hidden_channels, cfg.flow_wn_layers, cfg.flow_kernel)
        m = F.conv1d(h, w[pfx + ".post.weight"], w[pfx + ".post.bias"]) * x_mask
        x1 = (x1 - m) * torch.exp(-torch.zeros_like(m)) * x_mask    # mean_only: logs = 0 (modules.py:446-447,453)
        x = torch.cat([x0, x1], 1)
    return x


# ------------------------------------------------------------------------------------------
# NSF source (models.py:361-411, 455-467)
# ------------------------------------------------------------------------------------------
def sine_source(w, cfg, f0, rand_ini, noise_sine):
    """f0 [B,T] (Hz, 0 = unvoiced) -> har_source [B,1,L]; all fp32 on CPU (cumsum accumulates in fp64)."""
    upp = cfg.upp
    f0 = f0.float()[:, :, None]                                       # [B,T,1]
    rad = (f0 / cfg.sr) % 1                                           # :377
    ri = rand_ini.clone().float()
    ri[:, 0] = 0                                                      # :381
    rad[:, 0, :] = rad[:, 0, :] + ri                                  # :382
    tmp = torch.cumsum(rad, 1)                                        # :383
    tmp = tmp * upp                                                   # :384
    tmp = F.interpolate(tmp.transpose(2, 1), scale_factor=float(upp), mode="linear", align_corners=True).transpose(2, 1)
    rad_up = F.interpolate(rad.transpose(2, 1), scale_factor=float(upp), mode="nearest").transpose(2, 1)
    tmp = tmp % 1                                                     # :396
    wrap = (tmp[:, 1:, :] - tmp[:, :-1, :]) < 0                       # :397
    shift = torch.zeros_like(rad_up)
    shift[:, 1:, :] = wrap * -1.0                                     # :398-399
    sine = torch.sin(torch.cumsum(rad_up + shift, dim=1) * 2 * np.pi) # :400-402
    sine = sine * 0.1                                                 # sine_amp :403
    uv = torch.ones_like(f0) * (f0 > 0)                               # :353-359
    uv = F.interpolate(uv.transpose(2, 1), scale_factor=float(upp), mode="nearest").transpose(2, 1)
    noise_amp = uv * 0.003 + (1 - uv) * 0.1 / 3                       # :408
    noise = noise_amp * noise_sine.float()                            # :409
    sine = sine * uv + noise                                          # :410
    sine = sine.to(w["dec.m_source.l_linear.weight"].dtype)           # :464-465 (`.half()` when is_half; no-op in fp32)
    merged = torch.tanh(F.linear(sine, w["dec.m_source.l_linear.weight"], w["dec.m_source.l_linear.bias"]))  # :466
    return merged.transpose(1, 2)                                     # [B,1,L]


# ------------------------------------------------------------------------------------------
# GeneratorNSF (models.py:542-564); ResBlock1/2 (modules.py:295-308 / 346-355)
# ------------------------------------------------------------------------------------------
def _resblock(w, pfx, x, ks, dils, kind):
    if kind == "1":
        for d_i, d in enumerate(dils):
            xt = F.leaky_relu(x, LRELU_SLOPE)
            xt = F.conv1d(xt, w[f"{pfx}.convs1.{d_i}.weight"], w[f"{pfx}.convs1.{d_i}.bias"], dilation=d, padding=(ks * d - d) // 2)
            xt = F.leaky_relu(xt, LRELU_SLOPE)
            xt = F.conv1d(xt, w[f"{pfx}.convs2.{d_i}.weight"], w[f"{pfx}.convs2.{d_i}.bias"], padding=(ks - 1) // 2)
            x = xt + x
    else:
        for d_i, d in enumerate(dils):
            xt = F.leaky_relu(x, LRELU_SLOPE)
            xt = F.conv1d(xt, w[f"{pfx}.convs.{d_i}.weight"], w[f"{pfx}.convs.{d_i}.bias"], dilation=d, padding=(ks * d - d) // 2)
            x = xt + x
    return x


def generator_nsf(w, cfg, x, har_source, g, taps: Optional[dict] = None):
    """GeneratorNSF.forward (models.py:542-564); with har_source=None the plain `Generator.forward`
    (models.py:293-311) of the no-f0 synthesizers: the same ladder without the noise_convs injection."""
    x = F.conv1d(x, w["dec.conv_pre.weight"], w["dec.conv_pre.bias"], padding=3)
    x = x + F.conv1d(g, w["dec.cond.weight"], w["dec.cond.bias"])
    nk = cfg.num_kernels
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, w[f"dec.ups.{i}.weight"], w[f"dec.ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        if har_source is not None:
            kn, sn, pn = cfg.noise_conv_geometry(i)
            x = x + F.conv1d(har_source, w[f"dec.noise_convs.{i}.weight"], w[f"dec.noise_convs.{i}.bias"], stride=sn, padding=pn)
        if taps is not None:
            taps[f"dec.ups_plus_noise.{i}"] = x
        xs = None
        for j in range(nk):
            r = _resblock(w, f"dec.resblocks.{i * nk + j}", x, cfg.resblock_kernel_sizes[j],
                          cfg.resblock_dilation_sizes[j], cfg.resblock)
            xs = r if xs is None else xs + r
        x = xs / nk
        if taps is not None:
            taps[f"dec.stage.{i}"] = x
    x = F.leaky_relu(x)                                               # default slope 0.01 (models.py:561)
    x = F.conv1d(x, w["dec.conv_post.weight"], None, padding=3)
    return torch.tanh(x)


# ------------------------------------------------------------------------------------------
# infer (models.py:682-693 / 798-809)
# ------------------------------------------------------------------------------------------
@torch.no_grad()
def infer(sd_folded, cfg, phone, phone_lengths, pitch, nsff0, sid, noise_zp, rand_ini, noise_sine,
          rate=None, taps: Optional[dict] = None):
    """Returns (o[B,1,L], x_mask[B,1,T], (z, z_p, m_p, logs_p)) exactly like the reference.

    `sd_folded` is `fold_weight_norm(cpt["weight"])` in fp32.  `taps`, if given, receives
    intermediate tensors for stage-level parity checks.
    """
    w = sd_folded
    g = F.embedding(sid, w["emb_g.weight"]).unsqueeze(-1)            # [B,256,1]
    m_p, logs_p, x_mask = text_encoder(w, cfg, phone.to(w["emb_g.weight"].dtype), pitch, phone_lengths)
    z_p = (m_p + torch.exp(logs_p) * noise_zp * 0.66666) * x_mask
    if rate:
        head = int(z_p.shape[2] * rate)
        z_p = z_p[:, :, -head:]
        x_mask = x_mask[:, :, -head:]
        nsff0 = nsff0[:, -head:]
        noise_sine = noise_sine[:, -head * cfg.upp:]
    z = flow_reverse(w, cfg, z_p, x_mask, g)
    har = sine_source(w, cfg, nsff0, rand_ini, noise_sine)
    if taps is not None:
        taps.update({"m_p": m_p, "logs_p": logs_p, "z_p": z_p, "z": z, "har_source": har})
    o = generator_nsf(w, cfg, z * x_mask, har, g, taps)
    return o, x_mask, (z, z_p, m_p, logs_p)


@torch.no_grad()
def infer_nono(sd_folded, cfg, phone, phone_lengths, sid, noise_zp, rate=None, taps: Optional[dict] = None):
    """`SynthesizerTrnMs{256,768}NSFsid_nono.infer` (models.py:905-915 / :1011-1021): no pitch embedding in the
    text encoder (`enc_p(phone, None, lengths)`, models.py:50-53), plain `Generator` decoder, one RNG draw."""
    w = sd_folded
    g = F.embedding(sid, w["emb_g.weight"]).unsqueeze(-1)
    m_p, logs_p, x_mask = text_encoder(w, cfg, phone.to(w["emb_g.weight"].dtype), None, phone_lengths)
    z_p = (m_p + torch.exp(logs_p) * noise_zp * 0.66666) * x_mask
    if rate:
        head = int(z_p.shape[2] * rate)
        z_p = z_p[:, :, -head:]
        x_mask = x_mask[:, :, -head:]
    z = flow_reverse(w, cfg, z_p, x_mask, g)
    if taps is not None:
        taps.update({"m_p": m_p, "logs_p": logs_p, "z_p": z_p, "z": z})
    o = generator_nsf(w, cfg, z * x_mask, None, g, taps)
    return o, x_mask, (z, z_p, m_p, logs_p)
