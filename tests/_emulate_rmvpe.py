"""CPU emulation (torch, fp32) of what the RMVPE device path computes, step for step: images as rows of W + 1 pixels with a zero
pad pixel, convolutions as row-offset tap contractions with the model's re-laid weights (`RMVPE._src` / `._host`; on the wide levels
`pack` pixels per 64-channel row and block-Toeplitz weights), the (phase, channel) transposed-convolution GEMM + shuffle, the
side-by-side `cat` buffer, the GRU operand packing.  Host-logic test aid only
(tests/test_rmvpe_host.py): it checks the weight transforms and index maps of comfy_rvc_b200/rmvpe.py against the oracle
without a GPU; the kernels themselves are checked on the GPU (tests/test_rmvpe_gpu.py)."""
import torch
import torch.nn.functional as F


def tap_conv(x, w, b, W, taps, relu=False, res=None, mask=True):
    """x [rows][cin], w [taps][cin][cout] -> y [rows][cout]; tap t reads row r + off(t), rows outside [0, rows) are zero."""
    rows, Wp = x.shape[0], W + 1
    if taps == 9:
        offs = [(t // 3) * Wp + (t % 3) - (Wp + 1) for t in range(9)]
    elif taps == 4:
        offs = [(t // 2) * Wp + (t % 2) for t in range(4)]
    else:
        offs = [0]
    lo, hi = max(0, -min(offs)), max(0, max(offs))
    xp = F.pad(x, (0, 0, lo, hi))
    y = b[None, :].clone().repeat(rows, 1)
    for t, o in enumerate(offs):
        y += xp[lo + o:lo + o + rows] @ w[t][:x.shape[1]]
    if relu:
        y = torch.relu(y)
    if res is not None:
        y = y + res
    if mask:
        y[torch.arange(rows) % Wp == W] = 0
    return y


def hidden_logits(model, img, Tp):
    """Mirror of RMVPE._hidden_from_img on CPU tensors.  img [Tp][P0][img_c] fp32 (pixel layout, pad pixels zero)."""
    S, Hb = model._src, model._host

    def block(p, x, x32, F):
        res = tap_conv(x, S[p + "sc.w"], Hb[p + "sc.b"], F, 1, mask=False) if (p + "sc.w") in S else x32
        h = tap_conv(x, S[p + "c1.w"], Hb[p + "c1.b"], F, 9, relu=True)
        return tap_conv(h, S[p + "c2.w"], Hb[p + "c2.b"], F, 9, relu=True, res=res)

    skips = []
    x = None
    cur = img                                                     # pixel layout [H][P][c]
    for i in range(model.n_levels):
        c, H, W, pk, P, FP, rows = model._geom(i, Tp)
        x = cur.reshape(rows, -1)                                 # the same memory as GEMM rows of pk pixels
        x32 = None
        for j in range(model.n_blocks):
            x32 = block(f"unet.encoder.layers.{i}.conv.{j}.", x, x32, FP - 1)
            x = x32
        skips.append(x32)
        im = x32.reshape(H, P, c)[:, :W]
        pooled = (((im[0::2, 0::2] + im[0::2, 1::2]) + im[1::2, 0::2]) + im[1::2, 1::2]) * 0.25
        _, H2, W2, _, P2, _, _ = model._geom(i + 1, Tp)
        cur = F.pad(pooled, (0, 0, 0, P2 - W2))                   # [H2][P2][c]
    c, H, W, pk, P, FP, rows = model._geom(model.n_levels, Tp)
    x = cur.reshape(rows, -1)
    x32 = None
    for i in range(model.n_inter):
        for j in range(model.n_blocks):
            x32 = block(f"unet.intermediate.layers.{i}.conv.{j}.", x, x32, FP - 1)
            x = x32
    for i in range(model.n_levels):
        p = f"unet.decoder.layers.{i}."
        _, Hi, Wi, _, Pi, _, _ = model._geom(model.n_levels - i, Tp)
        c, H, W, pk, P, FP, rows = model._geom(model.n_levels - 1 - i, Tp)
        xin = x.reshape(Hi * Pi, -1)                              # pixel rows of the level below
        g = tap_conv(xin, S[p + "up.w"], Hb[p + "up.b"], Pi - 1, 4, relu=True, mask=False).reshape(Hi, Pi, 4, c)
        up = torch.zeros(H, P, c)
        for ph in range(4):
            up[(ph >> 1)::2, (ph & 1):W:2] = g[:, :Wi, ph]
        skip = skips[model.n_levels - 1 - i].reshape(H, P, c)
        # concat buffer: per GEMM row [pk x c up | pk x c skip]
        cat = torch.cat([up.reshape(rows, pk * c), skip.reshape(rows, pk * c)], dim=1)
        x, x32 = cat, None
        for j in range(model.n_blocks):
            x32 = block(p + f"conv2.{j}.", x, x32, FP - 1)
            x = x32
    c, H, W, pk, P, FP, rows = model._geom(0, Tp)
    cnn = tap_conv(x, S["cnn.w"], Hb["cnn.b"], FP - 1, 9)
    gx = cnn.reshape(H, P, 16)[:, :W, :3].permute(0, 2, 1).reshape(H, 3 * W)
    gi = gx @ S["gru.ih.w"][0] + Hb["gru.ih.b"]
    hh, bh, Hn = Hb["gru.hh.w"], Hb["gru.hh.b"], model.gru_h
    out = torch.zeros(H, 2 * Hn)
    for d, order in ((0, range(H)), (1, range(H - 1, -1, -1))):
        h = torch.zeros(Hn)
        for t in order:
            g_i = gi[t, d * 3 * Hn:(d + 1) * 3 * Hn]
            g_h = hh[d] @ h + bh[d]
            r = torch.sigmoid(g_i[:Hn] + g_h[:Hn])
            z = torch.sigmoid(g_i[Hn:2 * Hn] + g_h[Hn:2 * Hn])
            n = torch.tanh(g_i[2 * Hn:] + r * g_h[2 * Hn:])
            h = (h - n) * z + n
            out[t, d * Hn:(d + 1) * Hn] = h
    return out @ S["fc.w"][0] + Hb["fc.b"]
