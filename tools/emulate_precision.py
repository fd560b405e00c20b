#!/usr/bin/env python
"""CPU what-if for the tensor-core decoder's storage formats (no GPU needed).

Replays GeneratorNSF with the rounding points of the CUDA tensor path -- 16-bit MMA operands (weights and
lrelu'd activations), fp32 accumulation -- and a selectable storage format for the residual stream `x`
between resblock convolutions, and prints the output SNR against the fp32 oracle
(formula of /root/reference/lib/karafan/compare.py:21-35).  Used to decide whether the residual stream can
live in HBM as fp16 (half the epilogue traffic of the resblock convolutions) without leaving the 45 dB gate.

    python tools/emulate_precision.py [--config 48k_v2] [--T 300]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from comfy_rvc_b200 import synthetic  # noqa: E402
from comfy_rvc_b200.config import NAMED_CONFIGS  # noqa: E402
from oracle import rvc_oracle  # noqa: E402  (analysis tool, not product code)


def q(x, dt):
    return x if dt is None else x.to(dt).float()


def snr_db(ref, est):
    ref, est = ref.astype(np.float64), est.astype(np.float64)
    return 10 * np.log10((ref ** 2).sum() / ((ref - est) ** 2).sum())


def decoder(w, cfg, x, har, g, op_rb, op_ladder, res_dt, acc_dt=None):
    lr = lambda t, s=0.1: F.leaky_relu(t, s)
    x = F.conv1d(q(x, op_ladder), q(w["dec.conv_pre.weight"], op_ladder), w["dec.conv_pre.bias"], padding=3)
    x = x + F.conv1d(g, w["dec.cond.weight"], w["dec.cond.bias"])
    nk = cfg.num_kernels
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        x = F.conv_transpose1d(q(lr(x), op_ladder), q(w[f"dec.ups.{i}.weight"], op_ladder), w[f"dec.ups.{i}.bias"], stride=u,
                               padding=(k - u) // 2)
        kn, sn, pn = cfg.noise_conv_geometry(i)
        x = x + F.conv1d(har, w[f"dec.noise_convs.{i}.weight"], w[f"dec.noise_convs.{i}.bias"], stride=sn, padding=pn)
        x = q(x, res_dt)                                   # stage input as stored
        xs = None
        for j in range(nk):
            pfx, ks = f"dec.resblocks.{i * nk + j}", cfg.resblock_kernel_sizes[j]
            r = x
            for d_i, d in enumerate(cfg.resblock_dilation_sizes[j]):
                xt = F.conv1d(q(lr(r), op_rb), q(w[f"{pfx}.convs1.{d_i}.weight"], op_rb), w[f"{pfx}.convs1.{d_i}.bias"],
                              dilation=d, padding=(ks * d - d) // 2)
                xt = F.conv1d(q(lr(xt), op_rb), q(w[f"{pfx}.convs2.{d_i}.weight"], op_rb), w[f"{pfx}.convs2.{d_i}.bias"],
                              padding=(ks - 1) // 2)
                r = xt + r
                if d_i < len(cfg.resblock_dilation_sizes[j]) - 1:
                    r = q(r, res_dt)
            xs = r if xs is None else xs + r
            if j < nk - 1:
                xs = q(xs, acc_dt)
        x = xs / nk
    x = F.conv1d(lr(x, 0.01), w["dec.conv_post.weight"], None, padding=3)
    return torch.tanh(x)


@torch.no_grad()
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="48k_v2")
    ap.add_argument("--T", type=int, default=300)
    args = ap.parse_args()
    cfg = NAMED_CONFIGS[args.config]
    sd = synthetic.make_state_dict(cfg)
    w = rvc_oracle.fold_weight_norm(sd)
    phone, lens, pitch, pitchf, sid = synthetic.make_inputs(cfg, 1, args.T)
    noise = synthetic.draw_noise(cfg, 1, args.T)
    taps = {}
    o, x_mask, (z, *_rest) = rvc_oracle.infer(w, cfg, phone, lens, pitch, pitchf, sid, *noise, taps=taps)
    g = F.embedding(sid, w["emb_g.weight"]).unsqueeze(-1)
    ref = o[0, 0].numpy()
    h, b = torch.float16, torch.bfloat16
    for name, op_rb, op_l, res, acc in [
        ("fp32 everywhere (sanity)", None, None, None, None),
        ("fp16 operands, fp32 residual (current fp16 mode)", h, h, None, None),
        ("fp16 operands, fp16 residual", h, h, h, None),
        ("fp16 operands, fp16 residual + fp16 branch sum", h, h, h, h),
        ("bf16 rb operands, fp32 residual (current bf16 mode)", b, h, None, None),
        ("bf16 rb operands, fp16 residual", b, h, h, None),
        ("bf16 rb operands, bf16 residual", b, h, b, None),
    ]:
        est = decoder(w, cfg, z * x_mask, taps["har_source"], g, op_rb, op_l, res, acc)[0, 0].numpy()
        print(f"{name:55s} SNR {snr_db(ref, est):6.2f} dB   max|x| stage acts n/a")


if __name__ == "__main__":
    main()
