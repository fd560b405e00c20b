"""Checkpoint -> kernel layouts ("weight-norm folded at load").

Input is the reference state_dict `cpt["weight"]` (fp16 on disk, `weight_g`/`weight_v` pairs,
/root/reference/training_cli.py:41-45; key layout SURVEY.md §8b).  Output is a dict of fp32
tensors in the layouts the CUDA kernels consume, registered by name with the C ABI:

  conv weights      [G][taps][C_in][C_out]   (C_out contiguous; channels-last implicit GEMM)
  WN in_layers      output channels interleaved (2c = tanh half c, 2c+1 = sigmoid half c) so the
                    gate of commons.py:211-218 is applied on register pairs in the epilogue
  flow pre/post     channel order pre-reversed for the layers that run in "flipped" state, which
                    removes modules.Flip (modules.py:373-380) from the run time entirely
  dec.ups           ConvTranspose1d split into `stride` phase groups of ceil(k/stride) taps
                    (SURVEY.md App. E), tap order = increasing input frame
  cond.*            dec.cond and all WN cond_layers stacked into one [n_cond][gin] matrix
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from .config import SynthConfig, state_dict_shapes


def fold_weight_norm(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """`*.weight_g` + `*.weight_v` -> `*.weight` using the primitive the reference's
    torch.nn.utils.weight_norm hook calls (bit-exact)."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if k.endswith("weight_v"):
            out[k[:-8] + "weight"] = torch._weight_norm(v.float(), sd[k[:-1] + "g"].float(), 0)
        elif not k.endswith("weight_g"):
            out[k] = v.float()
    return out


def validate_state_dict(cfg: SynthConfig, sd: Dict[str, torch.Tensor]):
    """Returns (missing, unexpected, mismatched) against the reference key layout."""
    want = state_dict_shapes(cfg)
    missing = [k for k in want if k not in sd]
    unexpected = [k for k in sd if k not in want and not k.startswith("enc_q.")]
    mismatched = [(k, tuple(sd[k].shape), want[k]) for k in want if k in sd and tuple(sd[k].shape) != want[k]]
    return missing, unexpected, mismatched


def conv_w(w: torch.Tensor) -> torch.Tensor:
    """Conv1d weight [C_out, C_in, k] -> [k][C_in][C_out]."""
    return w.permute(2, 1, 0).contiguous()


def up_geometry(k: int, u: int) -> Tuple[int, int, list]:
    """(pad, ntaps, g_off[p]) — must match `up_geom` in csrc/engine.cu."""
    pad = (k - u) // 2
    ntaps = (k + u - 1) // u
    g_off = [(p + pad) // u - (ntaps - 1) for p in range(u)]
    return pad, ntaps, g_off


def pack_conv_transpose(w: torch.Tensor, u: int) -> torch.Tensor:
    """ConvTranspose1d weight [C_in, C_out, k] -> [u][ntaps][C_in][C_out].

    out[n = j*u + p] = sum_m x[j + q_p - m] W[:, :, kappa_0 + m*u],  kappa_0 = (p+pad) % u,
    q_p = (p+pad) // u (models.py:498-511 with padding=(k-u)//2).  Tap t of
    the packed kernel reads input frame j + g_off[p] + t, i.e. m = ntaps-1-t.
    """
    cin, cout, k = w.shape
    pad, ntaps, _ = up_geometry(k, u)
    out = torch.zeros(u, ntaps, cin, cout, dtype=w.dtype)
    for p in range(u):
        k0 = (p + pad) % u
        for t in range(ntaps):
            kap = k0 + (ntaps - 1 - t) * u
            if kap < k:
                out[p, t] = w[:, :, kap]
    return out


def ups_is_dense(cfg: SynthConfig, i: int) -> bool:
    """Stage i's transposed conv can run as ONE ordinary 3-tap convolution C_in -> u*C_out = C_in (all phases side by
    side in N) on the specialised resblock kernel: k = 2u (two taps per phase) and u = 2 (channels halve, so u*C_out =
    C_in).  Same rule as `ups_dense` in csrc/engine.cu."""
    u, k = cfg.upsample_rates[i], cfg.upsample_kernel_sizes[i]
    cin = cfg.upsample_initial_channel >> i
    return u == 2 and k == 2 * u and cin in (32, 64, 128, 256)


def post_is_tc(cfg: SynthConfig) -> bool:
    """conv_post (k = 7, C_last -> 1) runs on the specialised tcgen05 resblock kernel when C_last is one of its channel
    counts.  Same rule as `post_tc` in csrc/engine.cu."""
    c_last = cfg.upsample_initial_channel >> cfg.num_upsamples
    return c_last in (32, 64, 128) and cfg.resblock == "1"


def pack_conv_transpose_dense(w: torch.Tensor, u: int) -> torch.Tensor:
    """ConvTranspose1d weight [C_in, C_out, k = 2u] -> ordinary conv weight [3 taps][C_in][u*C_out]: output row j of
    the dense conv holds the u phases j*u .. j*u+u-1 side by side (= the channels-last tensor [L*u][C_out] itself);
    tap tau reads input frame j - 1 + tau, taps a phase does not use are zero."""
    cin, cout, k = w.shape
    pad, ntaps, g_off = up_geometry(k, u)
    assert ntaps == 2 and all(-1 <= o <= 0 for o in g_off)
    ph = pack_conv_transpose(w, u)                      # [u][2][C_in][C_out]
    out = torch.zeros(3, cin, u * cout, dtype=w.dtype)
    for p in range(u):
        for t in range(ntaps):
            out[g_off[p] + t + 1, :, p * cout:(p + 1) * cout] = ph[p, t]
    return out


def pack(cfg: SynthConfig, sd: Dict[str, torch.Tensor]) -> Tuple[Dict[str, torch.Tensor], Dict[str, float]]:
    """Reference state_dict -> (packed fp32 tensors, scalars)."""
    w = fold_weight_norm(sd)
    H, C = cfg.hidden_channels, cfg.inter_channels
    half = C // 2
    P: Dict[str, torch.Tensor] = {}
    S: Dict[str, float] = {}
    P["emb_g"] = w["emb_g.weight"].contiguous()
    # ---- TextEncoder ----
    P["enc.emb.w"] = w["enc_p.emb_phone.weight"].t().contiguous()            # [C_f][H]
    P["enc.emb.b"] = w["enc_p.emb_phone.bias"].contiguous()
    if cfg.f0:
        P["enc.emb_pitch"] = w["enc_p.emb_pitch.weight"].contiguous()        # [256][H]
    for l in range(cfg.n_layers):
        a = f"enc_p.encoder.attn_layers.{l}"
        P[f"enc.{l}.qkv.w"] = torch.cat([w[f"{a}.conv_{n}.weight"][:, :, 0].t() for n in "qkv"], dim=1).contiguous()
        P[f"enc.{l}.qkv.b"] = torch.cat([w[f"{a}.conv_{n}.bias"] for n in "qkv"]).contiguous()
        # tensor-core attention operands: q|k|v with 128 channels per head (96 + zero pad), q pre-scaled by
        # 1/sqrt(dk) (attentions.py:229); relative tables as K-major fp16 matrices
        nh, dk = cfg.n_heads, H // cfg.n_heads
        wp = torch.zeros(H, 3 * nh * 128)
        bp = torch.zeros(3 * nh * 128)
        for pi, n in enumerate("qkv"):
            wn = w[f"{a}.conv_{n}.weight"][:, :, 0].t()                       # [H_in][H_out]
            bn = w[f"{a}.conv_{n}.bias"]
            sc = 1.0 / math.sqrt(dk) if n == "q" else 1.0
            for h in range(nh):
                c0 = (pi * nh + h) * 128
                wp[:, c0:c0 + dk] = wn[:, h * dk:(h + 1) * dk] * sc
                bp[c0:c0 + dk] = bn[h * dk:(h + 1) * dk] * sc
        P[f"enc.{l}.qkvp.w"] = wp.contiguous()
        P[f"enc.{l}.qkvp.b"] = bp.contiguous()
        nrel = 2 * cfg.window_size + 1
        ek = torch.zeros(32, 128)
        ek[:nrel, :dk] = w[f"{a}.emb_rel_k"][0]
        evt = torch.zeros(128, 64)
        evt[:dk, :nrel] = w[f"{a}.emb_rel_v"][0].t()
        P[f"enc.{l}.ek16"] = ek
        P[f"enc.{l}.evt16"] = evt
        P[f"enc.{l}.rel_k"] = w[f"{a}.emb_rel_k"][0].contiguous()            # heads_share -> [2w+1][dk]
        P[f"enc.{l}.rel_v"] = w[f"{a}.emb_rel_v"][0].contiguous()
        P[f"enc.{l}.o.w"] = w[f"{a}.conv_o.weight"][:, :, 0].t().contiguous()
        P[f"enc.{l}.o.b"] = w[f"{a}.conv_o.bias"].contiguous()
        for n, m in (("ln1", "norm_layers_1"), ("ln2", "norm_layers_2")):
            P[f"enc.{l}.{n}.g"] = w[f"enc_p.encoder.{m}.{l}.gamma"].contiguous()
            P[f"enc.{l}.{n}.b"] = w[f"enc_p.encoder.{m}.{l}.beta"].contiguous()
        f = f"enc_p.encoder.ffn_layers.{l}"
        P[f"enc.{l}.ffn1.w"] = conv_w(w[f"{f}.conv_1.weight"])
        P[f"enc.{l}.ffn1.b"] = w[f"{f}.conv_1.bias"].contiguous()
        P[f"enc.{l}.ffn2.w"] = conv_w(w[f"{f}.conv_2.weight"])
        P[f"enc.{l}.ffn2.b"] = w[f"{f}.conv_2.bias"].contiguous()
    P["enc.proj.w"] = w["enc_p.proj.weight"][:, :, 0].t().contiguous()       # [H][2C]
    P["enc.proj.b"] = w["enc_p.proj.bias"].contiguous()
    # ---- conditioning: dec.cond then (flow i, layer j) blocks, each interleaved like in_layers ----
    cond_w = [w["dec.cond.weight"][:, :, 0]]
    cond_b = [w["dec.cond.bias"]]
    # ---- flow: reversed(flows) = Flip, RCL3, Flip, RCL2, ... (models.py:189-191) ----
    flipped = False
    state = {}
    for i in reversed(range(cfg.n_flows)):
        flipped = not flipped
        state[i] = flipped
    for i in range(cfg.n_flows):
        p = f"flow.flows.{2 * i}"
        fl = state[i]
        pre = w[f"{p}.pre.weight"][:, :, 0]                                   # [H][half] over logical x0
        post_w = w[f"{p}.post.weight"][:, :, 0]                               # [half][H] -> logical x1
        post_b = w[f"{p}.post.bias"]
        if fl:  # physical channel p holds logical channel C-1-p
            pre = pre.flip(1)
            post_w = post_w.flip(0)
            post_b = post_b.flip(0)
        P[f"flow.{i}.pre.w"] = pre.t().contiguous()                           # [half][H]
        P[f"flow.{i}.pre.b"] = w[f"{p}.pre.bias"].contiguous()
        P[f"flow.{i}.post.w"] = post_w.t().contiguous()                       # [H][half]
        P[f"flow.{i}.post.b"] = post_b.contiguous()
        cw = w[f"{p}.enc.cond_layer.weight"][:, :, 0]                         # [2H*n][gin]
        cb = w[f"{p}.enc.cond_layer.bias"]
        for j in range(cfg.flow_wn_layers):
            wi = conv_w(w[f"{p}.enc.in_layers.{j}.weight"])                   # [k][H][2H]
            P[f"flow.{i}.in.{j}.w"] = torch.stack([wi[..., :H], wi[..., H:]], dim=-1).flatten(-2).contiguous()
            bi = w[f"{p}.enc.in_layers.{j}.bias"]
            P[f"flow.{i}.in.{j}.b"] = torch.stack([bi[:H], bi[H:]], dim=-1).flatten().contiguous()
            cwj, cbj = cw[j * 2 * H:(j + 1) * 2 * H], cb[j * 2 * H:(j + 1) * 2 * H]
            cond_w.append(torch.stack([cwj[:H], cwj[H:]], dim=1).flatten(0, 1))
            cond_b.append(torch.stack([cbj[:H], cbj[H:]], dim=1).flatten())
            rs_w = w[f"{p}.enc.res_skip_layers.{j}.weight"][:, :, 0]          # [2H or H][H]
            rs_b = w[f"{p}.enc.res_skip_layers.{j}.bias"]
            if j < cfg.flow_wn_layers - 1:
                P[f"flow.{i}.rs.{j}.res.w"] = rs_w[:H].t().contiguous()
                P[f"flow.{i}.rs.{j}.res.b"] = rs_b[:H].contiguous()
                P[f"flow.{i}.rs.{j}.skip.w"] = rs_w[H:].t().contiguous()
                P[f"flow.{i}.rs.{j}.skip.b"] = rs_b[H:].contiguous()
            else:
                P[f"flow.{i}.rs.{j}.skip.w"] = rs_w.t().contiguous()
                P[f"flow.{i}.rs.{j}.skip.b"] = rs_b.contiguous()
        # Tensor path: `post(sum_j skip_j(acts_j))` (modules.py:196-209 then :497-505) is a chain of 1x1 convolutions with
        # nothing non-linear between them, so the n skip launches and the post launch of a flow are ONE contraction of
        # the concatenated gate outputs [acts_0 | acts_1 | ...] (K = n H) with the folded weights W_skip_j W_post
        # (products in float64); masked frames are zeroed by the output mask either way.
        # ... and `pre` and the res 1x1 convolutions fold into the NEXT in_layer: with m the frame mask,
        #     h_j = (W_pre x0 + b_pre) m + sum_{l<j} (W_res_l acts_l + b_res_l) m        (modules.py:497, :199-205)
        # so in_layer_j(h_j) is ONE k-tap convolution over the channels [x0 | m | 0.. | acts_0 | .. | acts_{j-1}] (x0 and acts
        # arrive masked, m is a channel of its own that carries the bias terms through the zero padding and the mask edge
        # exactly) with weights W_in_j[tap] applied after W_pre / (b_pre + sum b_res_l) / W_res_l.  The x0 block is padded
        # to a multiple of 64 channels (one MMA k-block granule).  h is never materialised on the tensor path.
        xb = flow_x0_block(cfg)
        pre64, preb64 = P[f"flow.{i}.pre.w"].double(), P[f"flow.{i}.pre.b"].double()
        for j in range(cfg.flow_wn_layers):
            win = P[f"flow.{i}.in.{j}.w"].double()                               # [k][H][2H], gate-interleaved columns
            rows = torch.zeros(win.shape[0], xb + j * H, 2 * H, dtype=torch.float64)
            rows[:, :half] = torch.einsum("ch,khn->kcn", pre64, win)
            bsum = preb64 + sum(P[f"flow.{i}.rs.{l}.res.b"].double() for l in range(j))
            rows[:, half] = torch.einsum("h,khn->kn", bsum, win)
            for l in range(j):
                rows[:, xb + l * H: xb + (l + 1) * H] = torch.einsum("ah,khn->kan", P[f"flow.{i}.rs.{l}.res.w"].double(), win)
            P[f"flow.{i}.inf.{j}.w"] = rows.float().contiguous()
        wp64 = P[f"flow.{i}.post.w"].double()
        P[f"flow.{i}.sp.w"] = torch.cat([P[f"flow.{i}.rs.{j}.skip.w"].double() @ wp64 for j in range(cfg.flow_wn_layers)],
                                        dim=0).float().contiguous()            # [n H][half]
        P[f"flow.{i}.sp.b"] = (sum(P[f"flow.{i}.rs.{j}.skip.b"].double() for j in range(cfg.flow_wn_layers)) @ wp64
                               + P[f"flow.{i}.post.b"].double()).float().contiguous()
    P["cond.w"] = torch.cat(cond_w, dim=0).contiguous()
    P["cond.b"] = torch.cat(cond_b, dim=0).contiguous()
    # ---- GeneratorNSF ----
    if cfg.f0:
        S["dec.src.lin_w"] = float(w["dec.m_source.l_linear.weight"].reshape(-1)[0])
        S["dec.src.lin_b"] = float(w["dec.m_source.l_linear.bias"].reshape(-1)[0])
    P["dec.pre.w"] = conv_w(w["dec.conv_pre.weight"])
    P["dec.pre.b"] = w["dec.conv_pre.bias"].contiguous()
    nk = cfg.num_kernels
    for i, u in enumerate(cfg.upsample_rates):
        P[f"dec.ups.{i}.w"] = pack_conv_transpose(w[f"dec.ups.{i}.weight"], u).contiguous()
        P[f"dec.ups.{i}.b"] = w[f"dec.ups.{i}.bias"].contiguous()
        if ups_is_dense(cfg, i):
            P[f"dec.ups.{i}.w3"] = pack_conv_transpose_dense(w[f"dec.ups.{i}.weight"], u).contiguous()
            P[f"dec.ups.{i}.b3"] = w[f"dec.ups.{i}.bias"].repeat(u).contiguous()
        if cfg.f0:
            P[f"dec.noise.{i}.w"] = w[f"dec.noise_convs.{i}.weight"][:, 0, :].t().contiguous()   # [k][C]
            P[f"dec.noise.{i}.b"] = w[f"dec.noise_convs.{i}.bias"].contiguous()
        for j in range(nk):
            n = i * nk + j
            for d in range(len(cfg.resblock_dilation_sizes[j])):
                if cfg.resblock == "1":
                    for src, dst in (("convs1", "c1"), ("convs2", "c2")):
                        P[f"dec.rb.{n}.{dst}.{d}.w"] = conv_w(w[f"dec.resblocks.{n}.{src}.{d}.weight"])
                        P[f"dec.rb.{n}.{dst}.{d}.b"] = w[f"dec.resblocks.{n}.{src}.{d}.bias"].contiguous()
                else:
                    P[f"dec.rb.{n}.c.{d}.w"] = conv_w(w[f"dec.resblocks.{n}.convs.{d}.weight"])
                    P[f"dec.rb.{n}.c.{d}.b"] = w[f"dec.resblocks.{n}.convs.{d}.bias"].contiguous()
    P["dec.post.w"] = w["dec.conv_post.weight"][0].t().contiguous()           # [k][C]
    if post_is_tc(cfg):
        # conv_post on the tensor core: a C -> C convolution whose output channel 0 is conv_post (no bias) and whose other
        # output channels are zero, so it runs on the specialised resblock kernel with a tanh-of-column-0 epilogue
        wt = P["dec.post.w"].new_zeros(P["dec.post.w"].shape[0], P["dec.post.w"].shape[1], P["dec.post.w"].shape[1])
        wt[:, :, 0] = P["dec.post.w"]
        P["dec.post.wt"] = wt                                                 # [k][C_in][C_out]
        P["dec.post.bt"] = P["dec.post.w"].new_zeros(P["dec.post.w"].shape[1])
    return P, S


# ---- tensor-core path: 16-bit smem images of the decoder weights ------------------------------------
TC_KB, TC_N_MAX = 64, 256


TC_N_MAX_SMALL_M = 64   # text encoder / flow: few 128-row tiles per launch (T/128), so split C_out over more CTAs


def flow_x0_block(cfg: SynthConfig) -> int:
    """Columns of the [x0 | mask | zeros] block in front of the gate outputs in the flow's activation buffer (engine.cu)."""
    half = cfg.inter_channels // 2
    return (half + 1 + TC_KB - 1) // TC_KB * TC_KB


def tc_n_max_for_name(name: str) -> int:
    """N-tile cap per packed tensor (same rule as csrc/engine.cu): the decoder's long time axis fills the GPU with
    M tiles and wants wide N; the text encoder and flow have T/128 (~47) M tiles per launch, so C_out is split into
    64-column tiles to occupy 3-12x more SMs and to cut the weight bytes each CTA has to pull from L2."""
    return TC_N_MAX if name.startswith("dec.") else TC_N_MAX_SMALL_M


def tc_n_for(cout: int, n_max: int = TC_N_MAX) -> int:
    """MMA N tile: the largest multiple of 16 <= n_max dividing C_out (same rule as csrc/engine.cu)."""
    for n in range(n_max, 15, -16):
        if cout % n == 0:
            return n
    raise ValueError(cout)


def pack_tc(w: torch.Tensor, dtype: torch.dtype, n_max: int = TC_N_MAX) -> torch.Tensor:
    """[G][taps][C_in][C_out] (or [taps][C_in][C_out]) fp32 -> [G][C_out/N][taps][ceil(C_in/64)][N][64] 16-bit.

    One (tap, k-block) slice is an [N][64] K-major matrix (input channels contiguous, zero-padded to
    64) that TMA loads into SWIZZLE_128B shared memory as the tcgen05 B operand (csrc/conv_tc.cu)."""
    if w.dim() == 2:
        w = w[None, None]
    elif w.dim() == 3:
        w = w.unsqueeze(0)
    G, taps, cin, cout = w.shape
    N = tc_n_for(cout, n_max)
    nkb = (cin + TC_KB - 1) // TC_KB
    assert cout % N == 0 and N % 16 == 0, (cin, cout)
    if nkb * TC_KB != cin:
        w = torch.cat([w, w.new_zeros(G, taps, nkb * TC_KB - cin, cout)], dim=2)
    t = w.reshape(G, taps, nkb, TC_KB, cout // N, N).permute(0, 4, 1, 2, 5, 3)
    return t.contiguous().to(dtype)


def tc_weight_names(cfg: SynthConfig):
    """Packed fp32 tensors that also get a 16-bit `.tc` image (the decoder's dense convolutions)."""
    names = ["dec.pre.w", "enc.emb.w", "enc.proj.w"]
    for l in range(cfg.n_layers):
        names += [f"enc.{l}.qkv.w", f"enc.{l}.qkvp.w", f"enc.{l}.o.w", f"enc.{l}.ffn1.w", f"enc.{l}.ffn2.w"]
    for i in range(cfg.n_flows):
        names.append(f"flow.{i}.sp.w")
        for j in range(cfg.flow_wn_layers):
            names.append(f"flow.{i}.inf.{j}.w")
    nk = cfg.num_kernels
    for i in range(cfg.num_upsamples):
        names.append(f"dec.ups.{i}.w")
        if ups_is_dense(cfg, i):
            names.append(f"dec.ups.{i}.w3")
        for j in range(nk):
            n = i * nk + j
            for d in range(len(cfg.resblock_dilation_sizes[j])):
                if cfg.resblock == "1":
                    names += [f"dec.rb.{n}.c1.{d}.w", f"dec.rb.{n}.c2.{d}.w"]
                else:
                    names.append(f"dec.rb.{n}.c.{d}.w")
    if post_is_tc(cfg):
        names.append("dec.post.wt")
    return names
