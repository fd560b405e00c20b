"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU fp32 functional restatement of the feature extractor in front of the synthesis path:
`HubertModelWithFinalProj.extract_features` (/root/reference/lib/infer_pack/loaders.py:10-61), i.e. HuggingFace
`transformers.HubertModel` (base architecture: `feat_extract_norm="group"`, `do_stable_layer_norm=False`,
`conv_bias=False`) with `output_hidden_states=True`, element `output_layer - 1` of the hidden-state tuple
(loaders.py:56: index 8 for v1, 11 for v2) and the extra `final_proj` Linear for v1 (loaders.py:57).

The arithmetic lives in a third-party dependency that is not under /root/reference: `transformers` (unpinned in the
reference's requirements.txt / pyproject.toml; 5.5 in this image), `models/hubert/modeling_hubert.py`
(HubertFeatureEncoder, HubertFeatureProjection, HubertPositionalConvEmbedding + HubertSamePadLayer, HubertEncoder,
HubertEncoderLayer, HubertAttention).  Parity pin: `tests/golden/make_hubert_golden.py` runs the reference's own class
(loaders.py imported read-only) on seeded weights and audio in the build container; `tests/test_hubert_oracle.py`
replays the fixtures against this file.  Only `tests/` may import this module.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F


def fold_pos_conv_weight(sd: Dict[str, torch.Tensor]) -> torch.Tensor:
    """`weight_norm(conv, name="weight", dim=2)` of HubertPositionalConvEmbedding: g [1,1,k] * v / ||v||_(0,1), stored either
    as parametrizations (`original0` = g, `original1` = v; torch >= 2.1) or as `weight_g` / `weight_v` (older files)."""
    p = "encoder.pos_conv_embed.conv."
    if p + "parametrizations.weight.original0" in sd:
        g, v = sd[p + "parametrizations.weight.original0"], sd[p + "parametrizations.weight.original1"]
    elif p + "weight_g" in sd:
        g, v = sd[p + "weight_g"], sd[p + "weight_v"]
    else:
        return sd[p + "weight"].float()
    return torch._weight_norm(v.float(), g.float(), 2)


@torch.no_grad()
def extract_features(sd: Dict[str, torch.Tensor], hcfg, source: torch.Tensor, version: str = "v2") -> torch.Tensor:
    """source [B, n] (16 kHz) -> features [B, frames, 768] (v2) or [B, frames, 256] (v1), frames = (n - 400) // 320 + 1."""
    w = {k: v.float() for k, v in sd.items()}
    for _ in range(hcfg.num_hidden_layers):      # HubertEncoder's LayerDrop test draws torch.rand([]) per layer, in eval mode too
        torch.rand([])
    x = source.float()[:, None, :]
    # HubertFeatureEncoder: layer 0 = conv -> GroupNorm(C groups) -> GELU, layers 1.. = conv -> GELU (no bias)
    for i, (k, s) in enumerate(zip(hcfg.conv_kernel, hcfg.conv_stride)):
        x = F.conv1d(x, w[f"feature_extractor.conv_layers.{i}.conv.weight"], None, stride=s)
        if i == 0:
            c = x.shape[1]
            x = F.group_norm(x, c, w["feature_extractor.conv_layers.0.layer_norm.weight"],
                             w["feature_extractor.conv_layers.0.layer_norm.bias"], eps=1e-5)
        x = F.gelu(x)
    x = x.transpose(1, 2)                                                    # [B, T, 512]
    # HubertFeatureProjection (feat_proj_layer_norm=True)
    x = F.layer_norm(x, (x.shape[-1],), w["feature_projection.layer_norm.weight"], w["feature_projection.layer_norm.bias"],
                     hcfg.layer_norm_eps)
    h = F.linear(x, w["feature_projection.projection.weight"], w["feature_projection.projection.bias"])
    # HubertEncoder: h + GELU(SamePad(pos_conv(h))) -> LayerNorm -> layers (post-norm)
    kpos = hcfg.num_conv_pos_embeddings
    pos = F.conv1d(h.transpose(1, 2), fold_pos_conv_weight(sd), w["encoder.pos_conv_embed.conv.bias"], padding=kpos // 2,
                   groups=hcfg.num_conv_pos_embedding_groups)
    if kpos % 2 == 0:
        pos = pos[:, :, :-1]
    h = h + F.gelu(pos).transpose(1, 2)
    h = F.layer_norm(h, (h.shape[-1],), w["encoder.layer_norm.weight"], w["encoder.layer_norm.bias"], hcfg.layer_norm_eps)
    n_layers = 9 - 1 if version == "v1" else 12 - 1                           # hidden_states[output_layer - 1]
    nh = hcfg.num_attention_heads
    B, T, H = h.shape
    dk = H // nh
    for l in range(n_layers):
        p = f"encoder.layers.{l}."
        q = F.linear(h, w[p + "attention.q_proj.weight"], w[p + "attention.q_proj.bias"]) * dk ** -0.5
        k = F.linear(h, w[p + "attention.k_proj.weight"], w[p + "attention.k_proj.bias"])
        v = F.linear(h, w[p + "attention.v_proj.weight"], w[p + "attention.v_proj.bias"])
        q, k, v = (t.view(B, T, nh, dk).transpose(1, 2) for t in (q, k, v))
        a = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v
        a = a.transpose(1, 2).reshape(B, T, H)
        h = h + F.linear(a, w[p + "attention.out_proj.weight"], w[p + "attention.out_proj.bias"])
        h = F.layer_norm(h, (H,), w[p + "layer_norm.weight"], w[p + "layer_norm.bias"], hcfg.layer_norm_eps)
        f = F.gelu(F.linear(h, w[p + "feed_forward.intermediate_dense.weight"], w[p + "feed_forward.intermediate_dense.bias"]))
        h = h + F.linear(f, w[p + "feed_forward.output_dense.weight"], w[p + "feed_forward.output_dense.bias"])
        h = F.layer_norm(h, (H,), w[p + "final_layer_norm.weight"], w[p + "final_layer_norm.bias"], hcfg.layer_norm_eps)
    if version == "v1":
        h = F.linear(h, w["final_proj.weight"], w["final_proj.bias"])        # loaders.py:57
    return h


def frames_for(n_samples: int, hcfg) -> int:
    """Output length of the conv stack (valid convolutions)."""
    L = n_samples
    for k, s in zip(hcfg.conv_kernel, hcfg.conv_stride):
        L = (L - k) // s + 1
    return L

