"""CPU emulation (torch, fp32) of what the RMVPE device path computes, step for step: images as rows of W + 1 pixels with a zero
pad pixel, convolutions as row-offset tap contractions with the model's re-laid weights (`RMVPE._src` / `._host`), the (phase,
channel) transposed-convolution GEMM + shuffle, the channel-offset `cat`, the GRU operand packing.  Host-logic test aid only
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
    """Mirror of RMVPE._hidden_from_img on CPU tensors.  img [Tp][129][8] fp32."""
    S, Hb = model._src, model._host

    def block(p, x, x32, cout, W):
        res = tap_conv(x, S[p + "sc.w"], Hb[p + "sc.b"], W, 1, mask=False) if (p + "sc.w") in S else x32
        h = tap_conv(x, S[p + "c1.w"], Hb[p + "c1.b"], W, 9, relu=True)
        return tap_conv(h, S[p + "c2.w"], Hb[p + "c2.b"], W, 9, relu=True, res=res)

    x = img.reshape(Tp * 129, 8)
    H, W = Tp, 128
    skips = []
    x32 = None
    for i in range(model.n_levels):
        c = model.c0 << i
        for j in range(model.n_blocks):
            x32 = block(f"unet.encoder.layers.{i}.conv.{j}.", x, x32, c, W)
            x = x32
        skips.append(x32)
        im = x32.reshape(H, W + 1, c)[:, :W]
        pooled = (((im[0::2, 0::2] + im[0::2, 1::2]) + im[1::2, 0::2]) + im[1::2, 1::2]) * 0.25
        H, W = H // 2, W // 2
        x = F.pad(pooled, (0, 0, 0, 1)).reshape(H * (W + 1), c)
        x32 = None
    c = model.c0 << model.n_levels
    for i in range(model.n_inter):
        for j in range(model.n_blocks):
            x32 = block(f"unet.intermediate.layers.{i}.conv.{j}.", x, x32, c, W)
            x = x32
    cin = c
    for i in range(model.n_levels):
        p = f"unet.decoder.layers.{i}."
        co = cin // 2
        g = tap_conv(x, S[p + "up.w"], Hb[p + "up.b"], W, 4, relu=True, mask=False).reshape(H, W + 1, 4, co)
        up = torch.zeros(2 * H, 2 * W + 1, co)
        for ph in range(4):
            up[(ph >> 1)::2, (ph & 1):2 * W:2] = g[:, :W, ph]
        H, W = 2 * H, 2 * W
        x = torch.cat([up.reshape(H * (W + 1), co), skips[model.n_levels - 1 - i]], dim=1)
        x32 = None
        for j in range(model.n_blocks):
            x32 = block(p + f"conv2.{j}.", x, x32, co, W)
            x = x32
        cin = co
    cnn = tap_conv(x, S["cnn.w"], Hb["cnn.b"], W, 9)
    gx = cnn.reshape(H, W + 1, 16)[:, :W, :3].permute(0, 2, 1).reshape(H, 3 * W)
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
