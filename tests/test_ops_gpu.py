"""GPU: op-level parity of each CUDA kernel (through the C ABI) against fp64 CPU restatements."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from comfy_rvc_b200 import _lib, synthetic
from comfy_rvc_b200.config import NAMED_CONFIGS
from tests._emulate import conv_cl

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    return torch.device("cuda", 0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run_conv(x, w, bias, g_off, dil=1, out_stride=1, in_slope=1.0, in_len=None, **epi):
    """x [B][L][Cin] (cuda), w [G][taps][Cin][Cout] -> y; epi: cond, gate, res, res_mode, relu, out_slope, accum(y0),
    div, mask_pre, mask_post, out_len, alpha, gather, gidx."""
    lib = _lib.load()
    B, L, Cin = x.shape
    G, ntaps, _, Cout = w.shape
    gate = int(epi.get("gate", 0))
    Co = Cout // 2 if gate else Cout
    y = epi["y0"].clone() if "y0" in epi else torch.zeros(B, L * out_stride, Co, device=x.device)
    d = _lib.ConvDesc()
    d.x, d.x_bstride, d.ldx, d.L_in = x.data_ptr(), L * Cin, Cin, L
    d.in_len = in_len.data_ptr() if in_len is not None else None
    d.in_slope = in_slope
    d.w, d.bias = w.data_ptr(), (bias.data_ptr() if bias is not None else None)
    d.Cin, d.Cout, d.ntaps, d.dil, d.G = Cin, Cout, ntaps, dil, G
    for i, o in enumerate(g_off):
        d.g_off[i] = o
    d.Lj, d.out_stride = L, out_stride
    d.y, d.y_bstride, d.ldy = y.data_ptr(), L * out_stride * Co, Co
    keep = []
    if "cond" in epi:
        d.cond, d.cond_bstride = epi["cond"].data_ptr(), epi["cond"].stride(0)
    if "gather" in epi:
        d.gather, d.gidx, d.gidx_bstride = epi["gather"].data_ptr(), epi["gidx"].data_ptr(), epi["gidx"].shape[1]
    d.alpha = epi.get("alpha", 1.0)
    d.gate = gate
    d.mask_pre, d.mask_post = int(epi.get("mask_pre", 0)), int(epi.get("mask_post", 0))
    if "out_len" in epi:
        d.out_len = epi["out_len"].data_ptr()
    if "res" in epi:
        r = epi["res"]
        d.res, d.res_bstride, d.ldr, d.res_mode = r.data_ptr(), r.shape[1] * r.shape[2], r.shape[2], epi.get("res_mode", 1)
    d.out_slope, d.relu = epi.get("out_slope", 1.0), int(epi.get("relu", 0))
    d.accum, d.div = int("y0" in epi), epi.get("div", 1.0)
    st = lib.rvcb200_op_conv_f32(C.byref(d), B, _stream())
    assert st == 0, f"rvcb200_op_conv_f32 status {st}"
    torch.cuda.synchronize()
    return y


CONV_CASES = [
    # name, B, L, Cin, Cout, ntaps, dil, G(out_stride)
    ("1x1_192_576", 2, 300, 192, 576, 1, 1, 1),
    ("k3_ffn", 1, 257, 192, 768, 3, 1, 1),
    ("k7_pre", 1, 130, 192, 512, 7, 1, 1),
    ("k11_d5_c256", 1, 700, 256, 256, 11, 5, 1),
    ("k7_d3_c64", 2, 1000, 64, 64, 7, 3, 1),
    ("k3_d1_c32", 1, 2100, 32, 32, 3, 1, 1),
    ("k11_d5_c32", 1, 1500, 32, 32, 11, 5, 1),
    ("post_96", 1, 200, 192, 96, 1, 1, 1),
    ("ups_12x", 1, 50, 512, 256, 2, 1, 12),
    ("ups_2x_c32", 2, 333, 64, 32, 2, 1, 2),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_f32_plain(case):
    name, B, L, Cin, Cout, ntaps, dil, G = case
    dev = _dev()
    g = torch.Generator().manual_seed(hash(name) & 0xFFFF)
    x = torch.randn(B, L, Cin, generator=g)
    w = torch.randn(G, ntaps, Cin, Cout, generator=g) / math.sqrt(Cin * ntaps)
    b = torch.randn(Cout, generator=g)
    if G == 1:
        g_off = [-((ntaps - 1) // 2) * dil]
    else:
        g_off = [(p + G // 2) // G - (ntaps - 1) for p in range(G)]
    ref = conv_cl(x.double(), w.double(), b.double(), g_off=g_off, dil=dil, out_stride=G, in_slope=0.1)
    y = run_conv(x.to(dev), w.to(dev), b.to(dev), g_off, dil=dil, out_stride=G, in_slope=0.1)
    err = (y.cpu().double() - ref).abs().max().item()
    print(f"{name}: max abs err {err:.3e} (ref max {ref.abs().max().item():.3f})")
    assert err < 2e-5


def test_conv_f32_epilogues():
    dev = _dev()
    g = torch.Generator().manual_seed(11)
    B, L, Cin, H = 2, 150, 192, 192
    lens = torch.tensor([150, 97], dtype=torch.int32)
    m3 = (torch.arange(L)[None] < lens[:, None]).double()[:, :, None]
    x = torch.randn(B, L, Cin, generator=g)
    # gate + cond (WN in_layer), k=5
    w = torch.randn(1, 5, Cin, 2 * H, generator=g) / math.sqrt(5 * Cin)
    b = torch.randn(2 * H, generator=g)
    cond = torch.randn(B, 2 * H + 64, generator=g)
    pre = conv_cl(x.double(), w.double(), b.double(), g_off=[-2]) + cond[:, None, 64:].double()
    ref = torch.tanh(pre[..., 0::2]) * torch.sigmoid(pre[..., 1::2])
    cd = cond.to(dev)
    y = run_conv(x.to(dev), w.to(dev), b.to(dev), [-2], gate=1, cond=cd[:, 64:])   # strided view, like engine.cu
    assert (y.cpu().double() - ref).abs().max().item() < 2e-5
    # residual add, in place, masked after (flow res): h = (h + conv(acts)) * mask
    w1 = torch.randn(1, 1, Cin, H, generator=g) / math.sqrt(Cin)
    h = torch.randn(B, L, H, generator=g)
    ref = (h.double() + conv_cl(x.double(), w1.double(), b[:H].double())) * m3
    hd = h.to(dev)
    y = run_conv(x.to(dev), w1.to(dev), b[:H].to(dev), [0], res=hd, res_mode=1, mask_post=1, out_len=lens.to(dev))
    assert (y.cpu().double() - ref).abs().max().item() < 2e-5
    # flow post: x1 = (x1 - conv(in*mask)*mask)*mask with Cout = 96 (BN=32 path), strided destination
    w2 = torch.randn(1, 1, Cin, 96, generator=g) / math.sqrt(Cin)
    z = torch.randn(B, L, 192, generator=g)
    ref = z.double().clone()
    ref[..., 96:] = (ref[..., 96:] - conv_cl(x.double(), w2.double(), b[:96].double(), in_len=lens.long()) * m3) * m3
    # emulate the strided in-place update through the raw descriptor
    lib = _lib.load()
    zd, xd, wd, bd, ld = z.to(dev), x.to(dev), w2.to(dev), b[:96].to(dev), lens.to(dev)
    d = _lib.ConvDesc()
    d.x, d.x_bstride, d.ldx, d.L_in, d.in_len, d.in_slope = xd.data_ptr(), L * Cin, Cin, L, ld.data_ptr(), 1.0
    d.w, d.bias, d.Cin, d.Cout, d.ntaps, d.dil, d.G = wd.data_ptr(), bd.data_ptr(), Cin, 96, 1, 1, 1
    d.Lj, d.out_stride = L, 1
    d.y, d.y_bstride, d.ldy = zd.data_ptr() + 96 * 4, L * 192, 192
    d.res, d.res_bstride, d.ldr, d.res_mode = zd.data_ptr() + 96 * 4, L * 192, 192, 2
    d.alpha, d.out_slope, d.div = 1.0, 1.0, 1.0
    d.mask_pre, d.mask_post, d.out_len = 1, 1, ld.data_ptr()
    assert lib.rvcb200_op_conv_f32(C.byref(d), B, _stream()) == 0
    torch.cuda.synchronize()
    assert (zd.cpu().double() - ref).abs().max().item() < 2e-5
    # embedding epilogue: (conv + bias + gather) * alpha -> lrelu -> mask   (TextEncoder front)
    table = torch.randn(256, H, generator=g)
    idx = torch.randint(1, 256, (B, L), generator=g)
    ref = (conv_cl(x.double(), w1.double(), b[:H].double()) + table.double()[idx]) * math.sqrt(192.0)
    ref = torch.where(ref > 0, ref, ref * 0.1) * m3
    y = run_conv(x.to(dev), w1.to(dev), b[:H].to(dev), [0], gather=table.to(dev), gidx=idx.to(dev),
                 alpha=math.sqrt(192.0), out_slope=0.1, mask_post=1, out_len=lens.to(dev))
    assert (y.cpu().double() - ref).abs().max().item() < 1e-4
    # accumulate + divide (resblock branch mean), relu, mask_pre + residual (FFN tail)
    y0 = torch.randn(B, L, H, generator=g)
    ref = (y0.double() + conv_cl(x.double(), w1.double(), b[:H].double()) + h.double()) / 3.0
    y = run_conv(x.to(dev), w1.to(dev), b[:H].to(dev), [0], res=hd, y0=y0.to(dev), div=3.0)
    assert (y.cpu().double() - ref).abs().max().item() < 2e-5
    ref = torch.relu(conv_cl(x.double(), w1.double(), b[:H].double(), in_len=lens.long()))
    y = run_conv(x.to(dev), w1.to(dev), b[:H].to(dev), [0], relu=1, in_len=lens.to(dev))
    assert (y.cpu().double() - ref).abs().max().item() < 2e-5
    ref = conv_cl(x.double(), w1.double(), b[:H].double()) * m3 + h.double()
    y = run_conv(x.to(dev), w1.to(dev), b[:H].to(dev), [0], mask_pre=1, out_len=lens.to(dev), res=hd)
    assert (y.cpu().double() - ref).abs().max().item() < 2e-5


def test_layernorm():
    dev = _dev()
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1000, 192, generator=g) * 3 + 1
    gm, bt = torch.randn(192, generator=g), torch.randn(192, generator=g)
    ref = torch.nn.functional.layer_norm(x.double(), (192,), gm.double(), bt.double(), 1e-5)
    xd, y = x.to(dev), torch.empty(1000, 192, device=dev)
    gd, bd = gm.to(dev), bt.to(dev)
    assert lib.rvcb200_op_layernorm(xd.data_ptr(), gd.data_ptr(), bd.data_ptr(), y.data_ptr(), 1000, 192, 1e-5, _stream()) == 0
    torch.cuda.synchronize()
    assert (y.cpu().double() - ref).abs().max().item() < 1e-5


def _attention_ref(qkv, rel_k, rel_v, lens, n_heads, window, device=None):
    """fp64 banded attention on [B][T][3H] -> [B][T][H]; rows >= len are zero.  `device`: where the
    plain-torch fp64 arithmetic runs (the result comes back on the CPU)."""
    if device is not None:
        return _attention_ref(qkv.to(device), rel_k.to(device), rel_v.to(device), lens, n_heads, window).cpu()
    B, T, H3 = qkv.shape
    H = H3 // 3
    dk = H // n_heads
    out = torch.zeros(B, T, H, dtype=torch.float64, device=qkv.device)
    for b in range(B):
        L = int(lens[b])
        for h in range(n_heads):
            q = qkv[b, :L, h * dk:(h + 1) * dk].double() / math.sqrt(dk)
            k = qkv[b, :L, H + h * dk:H + (h + 1) * dk].double()
            v = qkv[b, :L, 2 * H + h * dk:2 * H + (h + 1) * dk].double()
            s = q @ k.t()
            rl = q @ rel_k.double().t()                        # [L][2w+1]
            i = torch.arange(L, device=qkv.device)
            for r in range(2 * window + 1):
                j = i + r - window
                ok = (j >= 0) & (j < L)
                s[i[ok], j[ok]] += rl[i[ok], r]
            p = torch.softmax(s, dim=-1)
            o = p @ v
            for r in range(2 * window + 1):
                j = i + r - window
                ok = (j >= 0) & (j < L)
                o[i[ok]] += p[i[ok], j[ok]][:, None] * rel_v.double()[r][None]
            out[b, :L, h * dk:(h + 1) * dk] = o
    return out


@pytest.mark.parametrize("T,lens", [(7, [7]), (64, [64, 33]), (300, [300, 211]), (1000, [1000])])
def test_attention_f32(T, lens):
    dev = _dev()
    lib = _lib.load()
    g = torch.Generator().manual_seed(T)
    B = len(lens)
    qkv = torch.randn(B, T, 576, generator=g)
    rel_k, rel_v = torch.randn(21, 96, generator=g) * 0.1, torch.randn(21, 96, generator=g) * 0.1
    ref = _attention_ref(qkv, rel_k, rel_v, lens, 2, 10)
    qd, kd, vd = qkv.to(dev), rel_k.to(dev), rel_v.to(dev)
    ld = torch.tensor(lens, dtype=torch.int32, device=dev)
    out = torch.full((B, T, 192), float("nan"), device=dev)
    st = lib.rvcb200_op_attention_f32(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), ld.data_ptr(), out.data_ptr(), B, T, 2, 96, 10, _stream())
    assert st == 0
    torch.cuda.synchronize()
    err = (out.cpu().double() - ref).abs().max().item()
    print(f"attention T={T}: max abs err {err:.3e}")
    assert err < 2e-5


@pytest.mark.parametrize("cfg_name,T,variant", [("48k_v2", 300, "contour"), ("40k", 1000, "uniform"),
                                                ("32k_v2", 64, "unvoiced"), ("48k_v2", 6000, "contour")])
def test_sine_source(cfg_name, T, variant):
    from oracle import rvc_oracle
    dev = _dev()
    lib = _lib.load()
    cfg = NAMED_CONFIGS[cfg_name]
    B = 2
    f0 = torch.from_numpy(np.stack([synthetic.make_f0(T, variant=variant, seed=s) for s in (1, 2)]))
    g = torch.Generator().manual_seed(9)
    noise = torch.randn(B, T * cfg.upp, 1, generator=g)
    w = {"dec.m_source.l_linear.weight": torch.tensor([[0.8125]]), "dec.m_source.l_linear.bias": torch.tensor([0.046875])}
    ref = rvc_oracle.sine_source(w, cfg, f0, torch.zeros(B, 1), noise)[:, 0]   # [B][L]
    scratch = torch.empty(int(lib.rvcb200_op_sine_scratch_bytes(B, T, cfg.upp)), dtype=torch.uint8, device=dev)
    f0d, nd = f0.to(dev), noise.reshape(B, -1).contiguous().to(dev)
    har = torch.empty(B, T * cfg.upp, device=dev)
    st = lib.rvcb200_op_sine_source(f0d.data_ptr(), nd.data_ptr(), har.data_ptr(), B, T, cfg.upp, cfg.sr, 0.8125, 0.046875,
                                    scratch.data_ptr(), _stream())
    assert st == 0
    torch.cuda.synchronize()
    diff = (har.cpu() - ref).abs()
    print(f"sine {cfg_name} T={T} {variant}: max abs err {diff.max().item():.3e}, bit-equal {(har.cpu() == ref).float().mean().item():.4f}")
    assert diff.max().item() < 2e-6
