"""CPU: the C-ABI library loads and exports every symbol include/rvcb200.h declares; the host
classes keep the reference's constructor/load contract; no compute calls here (no GPU)."""
import os
import re

import pytest
import torch

import comfy_rvc_b200 as rvc
from comfy_rvc_b200 import _lib, build, synthetic
from comfy_rvc_b200.config import NAMED_CONFIGS, state_dict_shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "rvcb200.h")).read()
    declared = set(re.findall(r"\b(rvcb200_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in rvcb200.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.rvcb200_abi_version() == 1


def test_struct_sizes_match_header_layout():
    import ctypes as C
    assert C.sizeof(_lib.RvcConfig) == 4 * (12 + 1 + 4 + 4 + 16 + 1 + 8 + 8 + 4 + 1)
    assert C.sizeof(_lib.RvcTap) == 24
    # the library reports sizeof() of every struct that crosses the boundary: the ctypes mirrors must agree
    lib = _lib.load()
    for which, mirror in enumerate((_lib.RvcConfig, _lib.RvcTap, _lib.ConvDesc, _lib.TcConvDesc)):
        assert lib.rvcb200_sizeof(which) == C.sizeof(mirror), mirror.__name__
    assert lib.rvcb200_sizeof(99) == -1


def test_reference_constructor_and_state_dict_contract():
    for name, cfg in NAMED_CONFIGS.items():
        shapes = state_dict_shapes(cfg)
        if cfg.num_upsamples == 4:
            assert len(shapes) == 457
        cls = rvc.SynthesizerTrnMs256NSFsid if cfg.feat_dim == 256 else rvc.SynthesizerTrnMs768NSFsid
        net = cls(*cfg.to_positional(), is_half=True)
        del net.enc_q                                                       # vc_infer_pipeline.py:219
        sd = {k: torch.zeros(s, dtype=torch.float16) for k, s in shapes.items()}
        res = net.load_state_dict(sd, strict=False)
        assert not res.missing_keys
        assert net.eval().to("cuda:0").half().float() is net
        # the no-f0 classes (models.py:812-1021): same constructor list, state_dict without pitch / source tensors
        from comfy_rvc_b200.config import nono
        shapes0 = state_dict_shapes(nono(cfg))
        assert set(shapes) - set(shapes0) == ({"enc_p.emb_pitch.weight", "dec.m_source.l_linear.weight", "dec.m_source.l_linear.bias"}
                                              | {f"dec.noise_convs.{i}.{n}" for i in range(cfg.num_upsamples) for n in ("weight", "bias")})
        cls0 = rvc.SynthesizerTrnMs256NSFsid_nono if cfg.feat_dim == 256 else rvc.SynthesizerTrnMs768NSFsid_nono
        net0 = cls0(*cfg.to_positional(), is_half=True)
        del net0.enc_q
        assert not net0.load_state_dict({k: torch.zeros(s, dtype=torch.float16) for k, s in shapes0.items()}, strict=False).missing_keys
        bad = dict(sd)
        bad.pop("dec.conv_post.weight")
        with pytest.raises(RuntimeError):
            net.load_state_dict(bad, strict=False)
    net = rvc.SynthesizerTrnMs256NSFsid(*NAMED_CONFIGS["40k"].to_positional()[:-1], "40k", is_half=False)   # sr2sr strings
    assert net.cfg.sr == 40000


def test_create_refuses_without_device_or_bad_config():
    import ctypes as C
    lib = _lib.load()
    ctx = C.c_void_p()
    cfg = _lib.RvcConfig()
    st = lib.rvcb200_create(C.byref(cfg), C.byref(ctx))
    assert st in (1, 5)          # bad config, or no CUDA device in the CPU container
    assert not ctx.value
