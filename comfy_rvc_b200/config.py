"""Synthesizer hyper-parameters.

The reference builds its synthesizers from an 18-element positional list stored in the
checkpoint (`cpt["config"]`, written by /root/reference/training_cli.py:46-65, canonical
values in /root/reference/lib/train/process_ckpt.py:31-137, consumed by
/root/reference/vc_infer_pipeline.py:205-218).  `SynthConfig.from_positional` accepts that
exact list so the drop-in classes keep the reference constructor signature.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

SR2SR = {"32k": 32000, "40k": 40000, "48k": 48000}  # models.py:566-570


@dataclass(frozen=True)
class SynthConfig:
    spec_channels: int
    segment_size: int
    inter_channels: int
    hidden_channels: int
    filter_channels: int
    n_heads: int
    n_layers: int
    kernel_size: int
    p_dropout: float
    resblock: str
    resblock_kernel_sizes: Tuple[int, ...]
    resblock_dilation_sizes: Tuple[Tuple[int, ...], ...]
    upsample_rates: Tuple[int, ...]
    upsample_initial_channel: int
    upsample_kernel_sizes: Tuple[int, ...]
    spk_embed_dim: int
    gin_channels: int
    sr: int
    feat_dim: int = 768          # 256 for v1 (TextEncoder256), 768 for v2 (TextEncoder768)
    window_size: int = 10        # attentions.py:18 (Encoder default)
    flow_kernel: int = 5         # models.py:654-656 / 770-772: ResidualCouplingBlock(inter, hidden, 5, 1, 3)
    flow_wn_layers: int = 3
    n_flows: int = 4
    f0: bool = True              # False: the `_nono` classes (no pitch embedding, plain `Generator`, models.py:244-317, 812-1021)

    @staticmethod
    def from_positional(args: Sequence, feat_dim: int, f0: bool = True) -> "SynthConfig":
        if len(args) != 18:
            raise ValueError(f"expected the 18-element reference config list, got {len(args)}")
        a = list(args)
        sr = a[17]
        if isinstance(sr, str):
            sr = SR2SR[sr]
        dils = tuple(tuple(int(d) for d in ds) for ds in a[11])
        if str(a[9]) != "1":
            # ResBlock2 builds exactly two convs, from dilation[0] and dilation[1] (modules.py:319-339); further entries
            # of the config list are ignored by the reference, fewer than two raise there as well
            if any(len(ds) < 2 for ds in dils):
                raise IndexError("ResBlock2 needs two dilations per kernel (modules.py:327,337)")
            dils = tuple(ds[:2] for ds in dils)
        return SynthConfig(
            spec_channels=int(a[0]), segment_size=int(a[1]), inter_channels=int(a[2]),
            hidden_channels=int(a[3]), filter_channels=int(a[4]), n_heads=int(a[5]),
            n_layers=int(a[6]), kernel_size=int(a[7]), p_dropout=float(a[8]), resblock=str(a[9]),
            resblock_kernel_sizes=tuple(int(k) for k in a[10]),
            resblock_dilation_sizes=dils,
            upsample_rates=tuple(int(u) for u in a[12]), upsample_initial_channel=int(a[13]),
            upsample_kernel_sizes=tuple(int(k) for k in a[14]), spk_embed_dim=int(a[15]),
            gin_channels=int(a[16]), sr=int(sr), feat_dim=int(feat_dim), f0=bool(f0),
        )

    def to_positional(self) -> list:
        return [
            self.spec_channels, self.segment_size, self.inter_channels, self.hidden_channels,
            self.filter_channels, self.n_heads, self.n_layers, self.kernel_size, self.p_dropout,
            self.resblock, list(self.resblock_kernel_sizes),
            [list(d) for d in self.resblock_dilation_sizes], list(self.upsample_rates),
            self.upsample_initial_channel, list(self.upsample_kernel_sizes), self.spk_embed_dim,
            self.gin_channels, self.sr,
        ]

    # ---- derived quantities -------------------------------------------------------------
    @property
    def upp(self) -> int:
        p = 1
        for u in self.upsample_rates:
            p *= u
        return p

    @property
    def num_upsamples(self) -> int:
        return len(self.upsample_rates)

    @property
    def num_kernels(self) -> int:
        return len(self.resblock_kernel_sizes)

    def stage_channels(self, i: int) -> int:
        """Channels after upsample stage i (0-based): models.py:499."""
        return self.upsample_initial_channel // (2 ** (i + 1))

    def noise_conv_geometry(self, i: int) -> Tuple[int, int, int]:
        """(kernel, stride, padding) of dec.noise_convs[i]: models.py:512-524."""
        if i + 1 < self.num_upsamples:
            s = 1
            for u in self.upsample_rates[i + 1:]:
                s *= u
            return 2 * s, s, s // 2
        return 1, 1, 0


# The five shipped configurations (/root/reference/configs/*.json:28-45; process_ckpt.py:31-137).
_COMMON = dict(
    segment_size=32, inter_channels=192, hidden_channels=192, filter_channels=768, n_heads=2,
    n_layers=6, kernel_size=3, p_dropout=0.0, resblock="1", resblock_kernel_sizes=(3, 7, 11),
    resblock_dilation_sizes=((1, 3, 5), (1, 3, 5), (1, 3, 5)), upsample_initial_channel=512,
    spk_embed_dim=109, gin_channels=256,
)

NAMED_CONFIGS: Dict[str, SynthConfig] = {
    "32k": SynthConfig(spec_channels=513, upsample_rates=(10, 4, 2, 2, 2),
                       upsample_kernel_sizes=(16, 16, 4, 4, 4), sr=32000, feat_dim=256, **_COMMON),
    "40k": SynthConfig(spec_channels=1025, upsample_rates=(10, 10, 2, 2),
                       upsample_kernel_sizes=(16, 16, 4, 4), sr=40000, feat_dim=256, **_COMMON),
    "48k": SynthConfig(spec_channels=1025, upsample_rates=(10, 6, 2, 2, 2),
                       upsample_kernel_sizes=(16, 16, 4, 4, 4), sr=48000, feat_dim=256, **_COMMON),
    "32k_v2": SynthConfig(spec_channels=513, upsample_rates=(10, 8, 2, 2),
                          upsample_kernel_sizes=(20, 16, 4, 4), sr=32000, feat_dim=768, **_COMMON),
    "40k_v2": SynthConfig(spec_channels=1025, upsample_rates=(10, 10, 2, 2),
                          upsample_kernel_sizes=(16, 16, 4, 4), sr=40000, feat_dim=768, **_COMMON),
    "48k_v2": SynthConfig(spec_channels=1025, upsample_rates=(12, 10, 2, 2),
                          upsample_kernel_sizes=(24, 20, 4, 4), sr=48000, feat_dim=768, **_COMMON),
}


def nono(cfg: SynthConfig) -> SynthConfig:
    """The no-f0 variant of a configuration (`SynthesizerTrnMs{256,768}NSFsid_nono`)."""
    from dataclasses import replace
    return replace(cfg, f0=False)


def resblock2(cfg: SynthConfig, kernels=(3, 7, 11), dilations=((1, 3), (1, 3), (1, 3))) -> SynthConfig:
    """A configuration with `resblock="2"` (modules.ResBlock2, modules.py:311-355: one conv per dilation, no pair;
    selected at models.py:496).  No shipped config uses it; the defaults are ResBlock2's own `dilation=(1, 3)`."""
    from dataclasses import replace
    return replace(cfg, resblock="2", resblock_kernel_sizes=tuple(kernels),
                   resblock_dilation_sizes=tuple(tuple(d)[:2] for d in dilations))


VARIANTS = {
    "nono": nono,
    "rb2": resblock2,
    # HiFi-GAN V3-like kernel/dilation table (values outside the shipped one: exercises the generic conv kernels)
    "rb2x": lambda cfg: resblock2(cfg, (3, 5, 7), ((1, 2), (2, 6), (3, 8))),
}


def resolve(name: str) -> SynthConfig:
    """`"48k_v2"`, `"40k:nono"`, `"40k:rb2"` ... -> SynthConfig (fixture metadata stores these names)."""
    base, *mods = name.split(":")
    cfg = NAMED_CONFIGS[base]
    for m in mods:
        cfg = VARIANTS[m](cfg)
    return cfg


def state_dict_shapes(cfg: SynthConfig) -> Dict[str, Tuple[int, ...]]:
    """Key -> shape of the reference `cpt["weight"]` state_dict (enc_q removed).

    Mirrors the module tree of /root/reference/lib/infer_pack/models.py:604-658 (256) and
    :720-773 (768); weight-normed layers appear as `weight_g`/`weight_v`
    (torch.nn.utils.weight_norm, dim=0).  457 tensors for the 4-stage configs.
    """
    H, F_, C = cfg.hidden_channels, cfg.filter_channels, cfg.inter_channels
    nh = cfg.n_heads
    dk = H // nh
    W = 2 * cfg.window_size + 1
    G = cfg.gin_channels
    s: Dict[str, Tuple[int, ...]] = {}
    # enc_p (models.py:33-41 / 80-88)
    s["enc_p.emb_phone.weight"] = (H, cfg.feat_dim)
    s["enc_p.emb_phone.bias"] = (H,)
    if cfg.f0:
        s["enc_p.emb_pitch.weight"] = (256, H)
    for l in range(cfg.n_layers):
        a = f"enc_p.encoder.attn_layers.{l}"
        s[f"{a}.emb_rel_k"] = (1, W, dk)
        s[f"{a}.emb_rel_v"] = (1, W, dk)
        for n in "qkvo":
            s[f"{a}.conv_{n}.weight"] = (H, H, 1)
            s[f"{a}.conv_{n}.bias"] = (H,)
        for n in ("norm_layers_1", "norm_layers_2"):
            s[f"enc_p.encoder.{n}.{l}.gamma"] = (H,)
            s[f"enc_p.encoder.{n}.{l}.beta"] = (H,)
        f = f"enc_p.encoder.ffn_layers.{l}"
        s[f"{f}.conv_1.weight"] = (F_, H, cfg.kernel_size)
        s[f"{f}.conv_1.bias"] = (F_,)
        s[f"{f}.conv_2.weight"] = (H, F_, cfg.kernel_size)
        s[f"{f}.conv_2.bias"] = (H,)
    s["enc_p.proj.weight"] = (2 * C, H, 1)
    s["enc_p.proj.bias"] = (2 * C,)
    # flow (models.py:163-183; modules.py:403-434, 137-182)
    half = C // 2
    for i in range(cfg.n_flows):
        p = f"flow.flows.{2 * i}"
        s[f"{p}.pre.weight"] = (H, half, 1)
        s[f"{p}.pre.bias"] = (H,)
        s[f"{p}.enc.cond_layer.bias"] = (2 * H * cfg.flow_wn_layers,)
        s[f"{p}.enc.cond_layer.weight_g"] = (2 * H * cfg.flow_wn_layers, 1, 1)
        s[f"{p}.enc.cond_layer.weight_v"] = (2 * H * cfg.flow_wn_layers, G, 1)
        for j in range(cfg.flow_wn_layers):
            s[f"{p}.enc.in_layers.{j}.bias"] = (2 * H,)
            s[f"{p}.enc.in_layers.{j}.weight_g"] = (2 * H, 1, 1)
            s[f"{p}.enc.in_layers.{j}.weight_v"] = (2 * H, H, cfg.flow_kernel)
            rs = 2 * H if j < cfg.flow_wn_layers - 1 else H
            s[f"{p}.enc.res_skip_layers.{j}.bias"] = (rs,)
            s[f"{p}.enc.res_skip_layers.{j}.weight_g"] = (rs, 1, 1)
            s[f"{p}.enc.res_skip_layers.{j}.weight_v"] = (rs, H, 1)
        s[f"{p}.post.weight"] = (half, H, 1)
        s[f"{p}.post.bias"] = (half,)
    # dec (models.py:470-540)
    if cfg.f0:
        s["dec.m_source.l_linear.weight"] = (1, 1)
        s["dec.m_source.l_linear.bias"] = (1,)
    U0 = cfg.upsample_initial_channel
    s["dec.conv_pre.weight"] = (U0, C, 7)
    s["dec.conv_pre.bias"] = (U0,)
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        cin, cout = U0 // (2 ** i), U0 // (2 ** (i + 1))
        s[f"dec.ups.{i}.bias"] = (cout,)
        s[f"dec.ups.{i}.weight_g"] = (cin, 1, 1)
        s[f"dec.ups.{i}.weight_v"] = (cin, cout, k)
        if cfg.f0:
            nk, _, _ = cfg.noise_conv_geometry(i)
            s[f"dec.noise_convs.{i}.weight"] = (cout, 1, nk)
            s[f"dec.noise_convs.{i}.bias"] = (cout,)
        for j, (k_r, ds) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            r = f"dec.resblocks.{i * cfg.num_kernels + j}"
            if cfg.resblock == "1":
                for grp in ("convs1", "convs2"):
                    for d in range(len(ds)):
                        s[f"{r}.{grp}.{d}.bias"] = (cout,)
                        s[f"{r}.{grp}.{d}.weight_g"] = (cout, 1, 1)
                        s[f"{r}.{grp}.{d}.weight_v"] = (cout, cout, k_r)
            else:  # ResBlock2: modules.py:311-355
                for d in range(len(ds)):
                    s[f"{r}.convs.{d}.bias"] = (cout,)
                    s[f"{r}.convs.{d}.weight_g"] = (cout, 1, 1)
                    s[f"{r}.convs.{d}.weight_v"] = (cout, cout, k_r)
    s["dec.conv_post.weight"] = (1, U0 // (2 ** cfg.num_upsamples), 7)
    s["dec.cond.weight"] = (U0, G, 1)
    s["dec.cond.bias"] = (U0,)
    s["emb_g.weight"] = (cfg.spk_embed_dim, G)
    return s
