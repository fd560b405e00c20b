"""Drop-in replacements for the reference synthesizer classes.

`SynthesizerTrnMs256NSFsid` / `SynthesizerTrnMs768NSFsid` keep the reference contract
(/root/reference/lib/infer_pack/models.py:573-693, :696-809; call sites
/root/reference/vc_infer_pipeline.py:205-226, :100-101):

    net_g = Cls(*cpt["config"], is_half=config.is_half)     # 18 positional args + is_half
    del net_g.enc_q
    net_g.load_state_dict(cpt["weight"], strict=False)      # reference key layout, fp16 tensors
    net_g.eval().to(device); net_g = net_g.half() | net_g.float()
    o, x_mask, (z, z_p, m_p, logs_p) = net_g.infer(phone, phone_lengths, pitch, nsff0, sid)

but every FLOP of `infer` runs in librvcb200.so (hand-written sm_100a CUDA behind the C ABI of
include/rvcb200.h).  There is no PyTorch or CPU fallback: without the extension or without a CUDA
device `infer` raises.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch
from torch import nn

from . import _lib
from .config import SynthConfig
from .weights import pack, pack_tc, tc_n_max_for_name, tc_weight_names, validate_state_dict


class _IncompatibleKeys:
    def __init__(self, missing, unexpected):
        self.missing_keys, self.unexpected_keys = missing, unexpected

    def __repr__(self):
        return f"<missing={self.missing_keys} unexpected={self.unexpected_keys}>"


def graph_policy(key, frames: int, max_frames: int, repeat: bool, seen: "OrderedDict", graphs) -> bool:
    """Should this infer() call go through a CUDA graph?  Launch-bound sizes (`frames <= max_frames`) always do; larger calls
    do from the second time their key (B, T, precision) occurs (`repeat`), so a length that occurs once never pays for a
    capture; a key that already has a graph keeps using it.  `max_frames == 0` switches the mechanism off.  Records the key in
    `seen` (bounded, least recently used first out).  Host logic only: tested on the CPU (tests/test_pipeline_host.py)."""
    if max_frames <= 0:
        return False
    use = frames <= max_frames or (repeat and key in seen) or key in graphs
    seen[key] = None
    seen.move_to_end(key)
    while len(seen) > 256:
        seen.popitem(last=False)
    return use


class SynthesizerB200(nn.Module):
    """Common implementation; subclasses fix `feat_dim` (TextEncoder256 vs TextEncoder768)."""

    feat_dim = 768
    f0 = True            # False: the `_nono` classes

    def __init__(self, *args, is_half: bool = False, **kwargs):
        super().__init__()
        self.cfg = SynthConfig.from_positional(args, self.feat_dim, self.f0)
        if self.cfg.hidden_channels != 192 or self.cfg.n_heads != 2 or self.cfg.inter_channels != 192:
            raise ValueError("rvcb200 kernels are built for hidden=inter=192, 2 heads (every shipped RVC config)")
        self.enc_q = nn.Identity()          # `del net_g.enc_q` (vc_infer_pipeline.py:219) must work
        self.is_half = bool(is_half)
        # arithmetic of the heavy contractions: `is_half` / `.half()` select the tensor-core path with fp16 operands
        # (the reference's own GPU dtype), `.float()` the fp32 CUDA-core path; set_precision("bf16") is the third option
        self.precision = "fp16" if self.is_half else "fp32"
        self._device = torch.device("cuda", 0)
        self._ref_sd: Optional[Dict[str, torch.Tensor]] = None
        self._packed: Optional[Dict[str, torch.Tensor]] = None   # device tensors (kept alive for the ctx)
        self._ctx = None
        self._ws: Optional[torch.Tensor] = None
        self._tc_done = set()
        self.last_launches = 0
        # CUDA graphs: an infer() of at most `graph_max_frames` frames (B*T) is launch-bound (one C call enqueues ~170
        # kernels, ~13 us of host work each), so its launch sequence is captured once per (B, T, precision) into a CUDA
        # graph over static buffers and replayed (0 disables; RVCB200_GRAPH_FRAMES overrides the default)
        self.graph_max_frames = int(os.environ.get("RVCB200_GRAPH_FRAMES", "2500"))
        # Larger calls are replayed from a graph too once their (B, T, precision) has been seen before (a serving loop
        # with a fixed segment length, the benchmark's 60 s step: 9.80 -> 9.60 ms, the ~1 us gaps between its 165
        # kernels); a length that occurs once never pays for a capture.  RVCB200_GRAPH_REPEAT=0 disables.
        self.graph_repeat = os.environ.get("RVCB200_GRAPH_REPEAT", "1") != "0"
        self._seen_keys: "OrderedDict[tuple, None]" = OrderedDict()
        self.graph_cache_size = 16
        self._graphs: "OrderedDict[tuple, dict]" = OrderedDict()
        self._graph_ws: Optional[torch.Tensor] = None
        self._capture_stream = None
        self.last_graph_replay = False

    # ---- nn.Module protocol the reference callers use ------------------------------------------
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        missing, unexpected, mismatched = validate_state_dict(self.cfg, state_dict)
        if mismatched:
            raise RuntimeError(f"size mismatch for {mismatched[:3]} ...")
        if missing:
            # the reference would silently keep random-init values with strict=False; a synthesis
            # engine with missing weights is never what the caller wants
            raise RuntimeError(f"missing keys in state_dict: {missing[:5]} ...")
        if strict and unexpected:
            raise RuntimeError(f"unexpected keys in state_dict: {unexpected[:5]} ...")
        self._ref_sd = {k: v.detach().to("cpu") for k, v in state_dict.items() if not k.startswith("enc_q.")}
        self._release()
        return _IncompatibleKeys([], unexpected)

    def state_dict(self, *a, **k):
        return dict(self._ref_sd or {})

    def to(self, device=None, *args, **kwargs):
        if device is not None and not isinstance(device, torch.dtype):
            dev = torch.device(device)
            if dev.type == "cuda":
                dev = torch.device("cuda", dev.index if dev.index is not None else 0)
            if dev != self._device:
                self._device = dev
                self._release()
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if isinstance(device, int) else 0))

    def half(self):
        self.is_half = True
        return self.set_precision("fp16") if self.precision == "fp32" else self

    def float(self):
        self.is_half = False
        return self.set_precision("fp32")

    def remove_weight_norm(self):  # models.py:661-664: weight-norm is already folded at load
        return None

    def set_precision(self, precision: str):
        if precision not in _lib.PREC:
            raise ValueError(precision)
        if precision != self.precision:
            self._drop_graphs()                  # the captured launches point at the previous precision's weight images
        self.precision = precision
        if self._ctx is not None:
            self._ensure_tc()
        return self

    # ---- engine management ----------------------------------------------------------------------
    def _drop_graphs(self):
        self._graphs.clear()
        self._seen_keys.clear()
        self._graph_ws = None

    def _release(self):
        self._drop_graphs()
        if self._ctx is not None:
            _lib.load().rvcb200_destroy(self._ctx)
        self._ctx, self._packed, self._ws = None, None, None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _c_config(self) -> _lib.RvcConfig:
        cfg = self.cfg
        c = _lib.RvcConfig()
        c.feat_dim, c.inter_channels, c.hidden_channels = cfg.feat_dim, cfg.inter_channels, cfg.hidden_channels
        c.filter_channels, c.n_heads, c.n_layers = cfg.filter_channels, cfg.n_heads, cfg.n_layers
        c.enc_kernel, c.window_size, c.flow_kernel = cfg.kernel_size, cfg.window_size, cfg.flow_kernel
        c.flow_wn_layers, c.n_flows = cfg.flow_wn_layers, cfg.n_flows
        c.resblock_kind = 1 if cfg.resblock == "1" else 2
        c.n_res_kernels = cfg.num_kernels
        for j, (k, ds) in enumerate(zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)):
            c.res_kernels[j] = k
            c.n_res_dils[j] = len(ds)
            for d, v in enumerate(ds):
                c.res_dils[j][d] = v
        c.n_ups = cfg.num_upsamples
        for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
            c.up_rates[i], c.up_kernels[i] = u, k
        c.up_init_channels, c.gin_channels = cfg.upsample_initial_channel, cfg.gin_channels
        c.n_speakers = int(self._ref_sd["emb_g.weight"].shape[0])
        c.sr = cfg.sr
        c.no_f0 = 0 if cfg.f0 else 1
        return c

    def _materialize(self):
        if self._ctx is not None:
            return
        if self._ref_sd is None:
            raise RuntimeError("load_state_dict() must be called before infer()")
        if not torch.cuda.is_available():
            raise RuntimeError("comfy_rvc_b200 needs a CUDA (sm_100a) device; it has no CPU fallback")
        lib = _lib.load()
        packed, scalars = pack(self.cfg, self._ref_sd)
        with torch.cuda.device(self._device):
            self._packed = {k: v.to(self._device, dtype=torch.float32).contiguous() for k, v in packed.items()}
            ctx = C.c_void_p()
            cc = self._c_config()
            _lib.check(lib.rvcb200_create(C.byref(cc), C.byref(ctx)), None, "create")
            self._ctx = ctx
            for k, t in self._packed.items():
                _lib.check(lib.rvcb200_set_tensor(ctx, k.encode(), C.c_void_p(t.data_ptr()), t.numel(), 0), ctx, k)
            for k, v in scalars.items():
                _lib.check(lib.rvcb200_set_scalar(ctx, k.encode(), C.c_float(v)), ctx, k)
            _lib.check(lib.rvcb200_finalize(ctx), ctx, "finalize")
        self._tc_done = set()
        self._ensure_tc()

    def _ensure_tc(self):
        """Register the 16-bit tcgen05 weight images for the selected precision.

        The pair-closing resblock convolutions (`convs2`, half of the decoder FLOPs) use `precision` (fp16 | bf16)
        operands.  Everything that reads the fp16 activation stream -- `convs1`, conv_pre, the transposed-conv
        ladder, text encoder, flow -- uses fp16 operands: a bf16 stream costs ~5-9 dB of output SNR and tcgen05
        kind::f16 rejects mixed fp16 x bf16 operands (measured: illegal instruction)."""
        if self.precision == "fp32" or self.precision in self._tc_done:
            return
        lib = _lib.load()
        with torch.cuda.device(self._device):
            for name in tc_weight_names(self.cfg):
                closing = name.startswith("dec.rb.") and ".c2." in name
                prec = self.precision if closing else "fp16"
                dtype = torch.float16 if prec == "fp16" else torch.bfloat16
                t = pack_tc(self._packed[name].cpu(), dtype, tc_n_max_for_name(name)).to(self._device)
                key = f"{name}.tc"
                self._packed[f"{key}#{self.precision}"] = t          # keep alive; the engine holds the pointer
                _lib.check(lib.rvcb200_set_tensor(self._ctx, key.encode(), C.c_void_p(t.data_ptr()), t.numel(),
                                                  _lib.PREC[prec]), self._ctx, key)
            for l in range(self.cfg.n_layers):          # fp16 relative-position tables of the tcgen05 attention
                for nm in (f"enc.{l}.ek16", f"enc.{l}.evt16"):
                    t = self._packed[nm].to(torch.float16).contiguous()
                    self._packed[nm + "#h"] = t
                    _lib.check(lib.rvcb200_set_tensor(self._ctx, nm.encode(), C.c_void_p(t.data_ptr()), t.numel(), 1),
                               self._ctx, nm)
            _lib.check(lib.rvcb200_finalize(self._ctx), self._ctx, "finalize")
        self._tc_done = {self.precision}      # one `.tc` image per name: switching precision re-registers

    def _workspace(self, B: int, T: int, prec: int) -> torch.Tensor:
        need = int(_lib.load().rvcb200_workspace_bytes(self._ctx, B, T, prec))
        if need <= 0:
            raise RuntimeError("rvcb200_workspace_bytes failed")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self._device)
        return self._ws

    # ---- the hot path ---------------------------------------------------------------------------
    def draw_noise(self, B: int, T: int, device=None, dtype=torch.float32):
        """The reference's three RNG draws, same order and shapes (models.py:685/801, :378, :409), on
        `device`, so a seeded torch generator yields the stream the reference would consume there."""
        device = device or self._device
        L = T * self.cfg.upp
        nz = torch.randn(B, self.cfg.inter_channels, T, device=device, dtype=dtype)
        ri = torch.rand(B, 1, device=device)
        ns = torch.randn(B, L, 1, device=device, dtype=torch.float32)
        return nz, ri, ns

    @torch.no_grad()
    def infer(self, phone, phone_lengths, *rest, rate=None, noise=None, taps=None):
        """Same signature/return as the reference `infer`: `(phone, phone_lengths, pitch, nsff0, sid, rate=None)`
        for the f0 classes (models.py:682-693 / :798-809), `(phone, phone_lengths, sid, rate=None)` for the `_nono`
        classes (models.py:905-915 / :1011-1021).

        `noise=(noise_zp[B,192,T], rand_ini, noise_sine[B,L,1])` injects the RNG draws (parity tests);
        otherwise they are drawn with torch on the compute device in the reference's order.
        `taps` (dict name -> None) is filled with intermediate tensors (channels-last) for tests.
        """
        n_pos = 3 if self.f0 else 1
        if len(rest) == n_pos + 1 and rate is None:      # `rate` passed positionally, like the reference allows
            rest, rate = rest[:n_pos], rest[n_pos]
        if len(rest) != n_pos:
            raise TypeError(f"infer() takes {'(phone, phone_lengths, pitch, nsff0, sid)' if self.f0 else '(phone, phone_lengths, sid)'}")
        pitch, nsff0, sid = rest if self.f0 else (None, None, rest[0])
        if rate:
            return self._infer_rate(phone, phone_lengths, pitch, nsff0, sid, float(rate), noise)
        self._materialize()
        lib = _lib.load()
        dev = self._device
        cfg = self.cfg
        B, T, Cf = phone.shape
        if Cf != cfg.feat_dim:
            raise ValueError(f"phone has {Cf} features, model expects {cfg.feat_dim}")
        L = T * cfg.upp
        with torch.cuda.device(dev):
            phone_d = phone.to(dev, dtype=torch.float32).contiguous()
            len_d = phone_lengths.to(dev, dtype=torch.int64).contiguous()
            sid_d = sid.to(dev, dtype=torch.int64).reshape(-1).contiguous()
            if sid_d.numel() != B or len_d.numel() != B:
                raise ValueError("inconsistent batch/time dimensions")
            if self.f0:
                pitch_d = pitch.to(dev, dtype=torch.int64).contiguous()
                f0_d = nsff0.to(dev, dtype=torch.float32).contiguous()
                if pitch_d.shape != (B, T) or f0_d.shape != (B, T):
                    raise ValueError("inconsistent batch/time dimensions")
                if noise is None:
                    nz, _ri, ns = self.draw_noise(B, T, dev)
                else:
                    nz, _ri, ns = noise
                ns = ns.to(dev, dtype=torch.float32).reshape(B, L).contiguous()
            else:                                    # one RNG draw only (models.py:908)
                nz = noise[0] if noise is not None else torch.randn(B, cfg.inter_channels, T, device=dev)
            nz = nz.to(dev, dtype=torch.float32).contiguous()
            prec = _lib.PREC[self.precision]
            tap_arr, n_taps, tap_keep = None, 0, {}
            if taps is not None:
                shapes = self._tap_shapes(B, T)
                names = [n for n in taps if n in shapes]
                tap_arr = (_lib.RvcTap * max(len(names), 1))()
                for i, n in enumerate(names):
                    t = torch.zeros(shapes[n], device=dev, dtype=torch.float32)
                    tap_keep[n] = t
                    tap_arr[i].name = n.encode()
                    tap_arr[i].dst = t.data_ptr()
                    tap_arr[i].bytes = t.numel() * 4
                n_taps = len(names)
            ins = {"phone": phone_d, "len": len_d, "sid": sid_d, "nz": nz}
            if self.f0:
                ins.update(pitch=pitch_d, f0=f0_d, ns=ns)
            self.last_graph_replay = False
            use_graph = taps is None and graph_policy((B, T, prec), B * T, self.graph_max_frames, self.graph_repeat, self._seen_keys,
                                                       self._graphs)
            if use_graph:
                o, stats, z_p, z = self._infer_graphed(B, T, prec, ins)
            else:
                o, stats, z_p, z = self._outputs(B, T, dev)
                self._enqueue(B, T, prec, ins, (o, stats, z_p, z), self._workspace(B, T, prec), tap_arr, n_taps)
            if taps is not None:
                taps.update(tap_keep)
            x_mask = (torch.arange(T, device=dev).unsqueeze(0) < len_d.unsqueeze(1)).unsqueeze(1).to(torch.float32)
            Ci = cfg.inter_channels
            m_p = stats[:, :, :Ci].transpose(1, 2)
            logs_p = stats[:, :, Ci:].transpose(1, 2)
            return o, x_mask, (z.transpose(1, 2), z_p.transpose(1, 2), m_p, logs_p)

    def _infer_rate(self, phone, phone_lengths, pitch, nsff0, sid, rate, noise):
        """`infer(..., rate=r)` (models.py:802-806 / :908-912): the text encoder and the prior sample see the whole input,
        the flow and the decoder only the last `head = int(T * r)` frames.  Not on the pipeline's path (nothing in the
        reference passes `rate`), so it is built from two engine calls: a full `infer` for m_p / logs_p / z_p, then
        `rvcb200_infer_tail` on the tail of z_p.  RNG order as in the reference: prior noise for T frames first, then the
        source's draws for the tail."""
        dev, cfg = self._device, self.cfg
        B, T, _ = phone.shape
        head = int(T * rate)
        if head <= 0:
            head = T                         # `z_p[:, :, -0:]` is the whole tensor
        head = min(head, T)
        Lh = head * cfg.upp
        with torch.cuda.device(dev):
            if noise is None:
                nz = torch.randn(B, cfg.inter_channels, T, device=dev)
                ns_tail = None
                if self.f0:
                    torch.rand(B, 1, device=dev)                                             # models.py:378 (zeroed, but drawn)
                    ns_tail = torch.randn(B, Lh, 1, device=dev)
            else:
                nz = noise[0]
                ns_tail = None
                if self.f0:
                    ns = noise[2].reshape(B, -1)
                    ns_tail = ns if ns.shape[1] == Lh else ns[:, -Lh:]
            full_noise = (nz, None, torch.zeros(B, T * cfg.upp, device=dev)) if self.f0 else (nz,)
            keep = self.graph_max_frames
            self.graph_max_frames = 0
            try:
                if self.f0:
                    _, x_mask, (_, z_p, m_p, logs_p) = self.infer(phone, phone_lengths, pitch, nsff0, sid, noise=full_noise)
                else:
                    _, x_mask, (_, z_p, m_p, logs_p) = self.infer(phone, phone_lengths, sid, noise=full_noise)
            finally:
                self.graph_max_frames = keep
            lib = _lib.load()
            zp_tail = z_p[:, :, T - head:].transpose(1, 2).contiguous()                      # channels-last [B][head][C]
            len_tail = (phone_lengths.to(dev, torch.int64).reshape(-1) - (T - head)).clamp(0, head).contiguous()
            sid_d = sid.to(dev, torch.int64).reshape(-1).contiguous()
            f0_tail = nsff0.to(dev, torch.float32)[:, T - head:].contiguous() if self.f0 else None
            ns_d = ns_tail.to(dev, torch.float32).reshape(B, Lh).contiguous() if self.f0 else None
            prec = _lib.PREC[self.precision]
            ws = self._workspace(B, head, prec)
            o = torch.empty(B, 1, Lh, device=dev, dtype=torch.float32)
            z = torch.empty(B, head, cfg.inter_channels, device=dev, dtype=torch.float32)
            p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)
            st = lib.rvcb200_infer_tail(self._ctx, B, head, p(zp_tail), p(len_tail), p(f0_tail), p(sid_d), p(ns_d), p(o), p(z),
                                        p(ws), ws.numel(), prec, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            _lib.check(st, self._ctx, "infer_tail")
            self.last_launches += int(lib.rvcb200_last_launch_count(self._ctx))
            return o, x_mask[:, :, T - head:], (z.transpose(1, 2), z_p[:, :, T - head:], m_p, logs_p)

    def _outputs(self, B, T, dev):
        cfg = self.cfg
        return (torch.empty(B, 1, T * cfg.upp, device=dev, dtype=torch.float32),
                torch.empty(B, T, 2 * cfg.inter_channels, device=dev, dtype=torch.float32),
                torch.empty(B, T, cfg.inter_channels, device=dev, dtype=torch.float32),
                torch.empty(B, T, cfg.inter_channels, device=dev, dtype=torch.float32))

    def _enqueue(self, B, T, prec, ins, outs, ws, tap_arr=None, n_taps=0):
        """One `rvcb200_infer` call: enqueues every kernel of the step on the current stream of the device."""
        lib = _lib.load()
        o, stats, z_p, z = outs
        ptr = lambda name: C.c_void_p(ins[name].data_ptr()) if name in ins else C.c_void_p(None)
        stream = torch.cuda.current_stream(self._device).cuda_stream
        st = lib.rvcb200_infer(
            self._ctx, B, T, ptr("phone"), ptr("len"), ptr("pitch"), ptr("f0"), ptr("sid"), ptr("nz"), ptr("ns"),
            C.c_void_p(o.data_ptr()), C.c_void_p(stats.data_ptr()), C.c_void_p(z_p.data_ptr()), C.c_void_p(z.data_ptr()),
            C.c_void_p(ws.data_ptr()), ws.numel(), prec, tap_arr, n_taps, C.c_void_p(stream))
        _lib.check(st, self._ctx, "infer")
        self.last_launches = int(lib.rvcb200_last_launch_count(self._ctx))

    def _infer_graphed(self, B, T, prec, ins):
        """Launch-bound sizes: the step's launch sequence is captured once per (B, T, precision) over static input /
        output / workspace buffers and replayed.  The first call of a key runs eagerly on the static buffers (it is also
        the warm-up that a capture needs: per-device kernel attributes are set outside the capture) and then captures;
        later calls copy their inputs in, replay, and return fresh copies of the outputs (the caller owns what infer()
        returns, like the reference)."""
        dev = self._device
        key = (B, T, prec)
        need = int(_lib.load().rvcb200_workspace_bytes(self._ctx, B, T, prec))
        if need <= 0:
            raise RuntimeError("rvcb200_workspace_bytes failed")
        e = self._graphs.get(key)
        if e is None:
            while len(self._graphs) >= self.graph_cache_size:
                self._graphs.popitem(last=False)
            # graphs replay on one stream, so they share a workspace; when a new key needs a bigger one, a new buffer is
            # allocated for it and later keys -- graphs captured earlier keep (and keep alive) the buffer they point into
            if self._graph_ws is None or self._graph_ws.numel() < need:
                self._graph_ws = torch.empty(need, dtype=torch.uint8, device=dev)
            e = {"ins": {k: torch.empty_like(v) for k, v in ins.items()}, "outs": self._outputs(B, T, dev), "graph": None,
                 "ws": self._graph_ws}
            self._graphs[key] = e
        else:
            self._graphs.move_to_end(key)
        for k, v in ins.items():
            e["ins"][k].copy_(v)
        if e["graph"] is None:
            self._enqueue(B, T, prec, e["ins"], e["outs"], e["ws"])             # eager: result of this call + warm-up
            launches = self.last_launches
            g = torch.cuda.CUDAGraph()
            if self._capture_stream is None or self._capture_stream.device != dev:
                self._capture_stream = torch.cuda.Stream(device=dev)     # torch's default capture stream is per process, not per device
            with torch.cuda.graph(g, stream=self._capture_stream, capture_error_mode="thread_local"):
                self._enqueue(B, T, prec, e["ins"], e["outs"], e["ws"])
            e["graph"], e["launches"] = g, launches
        else:
            e["graph"].replay()
            self.last_launches = e["launches"]
            self.last_graph_replay = True
        return tuple(t.clone() for t in e["outs"])

    def _tap_shapes(self, B, T):
        cfg = self.cfg
        s = {"x_enc": (B, T, cfg.hidden_channels), "stats": (B, T, 2 * cfg.inter_channels),
             "z_p": (B, T, cfg.inter_channels), "z": (B, T, cfg.inter_channels), "har_source": (B, T * cfg.upp),
             "dec.pre": (B, T, cfg.upsample_initial_channel)}
        Lc = T
        for i, u in enumerate(cfg.upsample_rates):
            Lc *= u
            s[f"dec.ups.{i}"] = (B, Lc, cfg.stage_channels(i))
            s[f"dec.stage.{i}"] = (B, Lc, cfg.stage_channels(i))
        return s

    def forward(self, *a, **k):
        raise NotImplementedError("training forward() is out of scope; use infer()")


class SynthesizerTrnMs256NSFsid(SynthesizerB200):
    """v1 models: 256-d HuBERT features (reference models.py:573-693)."""
    feat_dim = 256


class SynthesizerTrnMs768NSFsid(SynthesizerB200):
    """v2 models: 768-d HuBERT features (reference models.py:696-809)."""
    feat_dim = 768


class SynthesizerTrnMs256NSFsid_nono(SynthesizerB200):
    """v1 models trained without pitch guidance (reference models.py:812-915): `infer(phone, phone_lengths, sid)`."""
    feat_dim = 256
    f0 = False


class SynthesizerTrnMs768NSFsid_nono(SynthesizerB200):
    """v2 models trained without pitch guidance (reference models.py:918-1021)."""
    feat_dim = 768
    f0 = False
