"""Segmented conversion driver: drop-in for the reference's `VC` (FeatureExtractor constants + `vc` + `pipeline`).

Reference: /root/reference/vc_infer_pipeline.py:23-196 (`VC.vc`, `VC.pipeline`),
/root/reference/pitch_extraction.py:14-45 (constants), :252-304 (`get_f0` post-processing and the coarse
pitch quantiser), /root/reference/config.py:124-141 (the x_pad/x_query/x_center/x_max tiers).

Same names, arguments and results as the reference:

    vc = VC(tgt_sr, config)                       # config: anything with x_pad/x_query/x_center/x_max/is_half/device
    pcm_f32 = vc.vc(hubert, net_g, sid, audio0, pitch, pitchf, times, index, big_npy, index_rate, version, protect)
    pcm_i16 = vc.pipeline(hubert, net_g, sid, audio, times, f0_up_key, f0_method, merge_type, file_index, index_rate,
                          if_f0, filter_radius, tgt_sr, resample_sr, rms_mix_rate, version, protect, crepe_hop_length,
                          f0_autotune, rmvpe_onnx, f0_file, f0_min, f0_max)

What is different is where the work happens (B200-first):
  * the reference loops over segments sequentially and, per segment, interpolates/blends with torch ops, copies the
    PCM to the host and `gc_collect()`s (a device sync + `empty_cache`).  Here a song is planned once on the host
    (the reference's own quiet-point search, bit-identical segmentation), every segment is enqueued on the stream
    without a host round trip (`rvcb200_op_prepare_feats` → `rvcb200_infer` → on-device trim), and the peak
    normalisation + int16 conversion run on the device (`rvcb200_op_absmax`, `rvcb200_op_to_int16`); one D2H of
    int16 PCM ends the call.
  * with `torch.distributed` initialised (one process per GPU) the segments are independent units: they are
    assigned longest-first to ranks, each rank converts its own, the only cross-segment reduction (`max|x|`, one
    float per rank) and the int16 pieces are gathered on the host; there is no collective on the data path.
  * segmentation is never changed for balance (it would change attention context and break parity).

HuBERT, the f0 estimators and faiss are upstream of this path: `model` is the caller's
HuBERT object, f0 estimators are looked up in `f0_method_dict` (register callables there) or passed as a
callable, and a preloaded `(index, big_npy)` tuple is honoured exactly like the reference does.
"""
from __future__ import annotations

import ctypes as C
import time
import traceback
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch
from scipy import signal

from . import _lib

MAX_INT16 = 32768                                    # lib/audio.py:14
_BH, _AH = (np.ascontiguousarray(c, dtype=np.float64) for c in signal.butter(N=5, Wn=48, btype="high", fs=16000))   # :21
_ZI = np.ascontiguousarray(signal.lfilter_zi(_BH, _AH), dtype=np.float64)   # what filtfilt computes on every call (a strided view)


@dataclass
class PipelineConfig:
    """The attributes of the reference `config` object that `VC` reads (config.py:24-141)."""
    x_pad: int = 3
    x_query: int = 10
    x_center: int = 60
    x_max: int = 64
    is_half: bool = True
    device: str = "cuda:0"
    rmvpe_model_path: Optional[str] = None     # BASE_MODELS_DIR/rmvpe.pt in the reference (pitch_extraction.py:193)

    @classmethod
    def for_device(cls, is_half: bool = True, gpu_mem_gb: Optional[int] = None, device: str = "cuda:0") -> "PipelineConfig":
        """config.py:124-141: half tier (3,10,60,64); fp32 tier (1,6,38,41); <= 4 GB tier (1,5,30,32)."""
        t = (3, 10, 60, 64) if is_half else (1, 6, 38, 41)
        if gpu_mem_gb is not None and gpu_mem_gb <= 4:
            t = (1, 5, 30, 32)
        return cls(*t, is_half=is_half, device=device)


def hz_to_mel(hz):
    """lib/audio.py:302-304."""
    return 2595 * np.log10(1 + hz / 700)


# ------------------------------------------------------------------------------------------------------------
# host-side planning (pure integer / numpy work; shared by every rank, testable without a GPU)
# ------------------------------------------------------------------------------------------------------------
def filtfilt_pad(audio: np.ndarray, t_pad: int, out: Optional[np.ndarray] = None) -> np.ndarray:
    """`np.pad(signal.filtfilt(bh, ah, audio), (t_pad, t_pad), mode="reflect")` (vc_infer_pipeline.py:122 and :141) in C
    (csrc/host_plan.cu: scipy's own recursion with the order fixed at compile time, bit-identical to scipy / numpy, ~2.3x
    faster).  `out`: optional float64 buffer of n + 2 * t_pad elements (e.g. pinned staging memory)."""
    x = np.ascontiguousarray(audio, dtype=np.float64)
    n = x.shape[0]
    if x.ndim != 1 or n <= 18 or n <= t_pad:            # scipy raises for n <= padlen; short clips keep the numpy path
        return np.pad(signal.filtfilt(_BH, _AH, audio), (t_pad, t_pad), mode="reflect")
    if out is None:
        out = np.empty(n + 2 * t_pad, dtype=np.float64)
    scratch = np.empty(n + 36, dtype=np.float64)
    st = _lib.load().rvcb200_host_filtfilt_pad(x.ctypes.data, n, _BH.ctypes.data, _AH.ctypes.data, _ZI.ctypes.data,
                                               len(_BH) - 1, t_pad, out.ctypes.data, scratch.ctypes.data)
    if st != 0:
        raise RuntimeError("rvcb200_host_filtfilt_pad: bad argument")
    return out[: n + 2 * t_pad]


def split_points(audio: np.ndarray, window: int, t_query: int, t_center: int, t_max: int,
                 padded: Optional[np.ndarray] = None) -> List[int]:
    """Quiet-point search of vc_infer_pipeline.py:123-135 on the high-passed audio.  `padded`: `audio` already
    reflect-padded by window // 2 on both sides (a view of the song's t_pad-padded buffer: the inner window // 2 samples
    of a longer reflect padding are the same values)."""
    audio_pad = padded if padded is not None else np.pad(audio, (window // 2, window // 2), mode="reflect")
    opt_ts: List[int] = []
    if audio_pad.shape[0] > t_max:
        # sliding |sum| over `window` samples: a cumulative-sum form of the reference's 160 shifted adds would round
        # differently, so keep the reference's accumulation order (float64, one add per shift) -- but only over the
        # 2*t_query samples around each centre that are ever looked at (each element's sum is independent of the others,
        # so this is bit-identical to summing the whole song: 3x less host work for a 10 min song)
        n = audio.shape[0]
        # float64 (what filtfilt returns): the same sums in the same order, in C over all host threads (csrc/host_plan.cu;
        # _lib.load() raises if the library is not built).  Other dtypes keep the reference's numpy loop.
        native = _lib.load().rvcb200_host_quiet_point if audio_pad.dtype == np.float64 and audio_pad.flags.c_contiguous else None
        for t in range(t_center, n, t_center):
            lo, hi = t - t_query, min(t + t_query, n)
            if native is not None:
                j = int(native(audio_pad.ctypes.data, lo, hi, window, 0))
                if j >= 0:
                    opt_ts.append(lo + j)
                    continue
            seg = np.zeros(hi - lo, dtype=audio.dtype)
            for i in range(window):
                seg += audio_pad[lo + i: hi + i]
            seg = np.abs(seg)
            opt_ts.append(lo + int(np.where(seg == seg.min())[0][0]))
    return opt_ts


@dataclass(frozen=True)
class Segment:
    index: int
    start: int                 # sample range [start, end) in the t_pad-padded 16 kHz audio
    end: int
    f0_start: int              # frame range in the song-level pitch arrays
    f0_end: int

    @property
    def n_samples(self) -> int:
        return self.end - self.start


def plan_segments(n_padded: int, opt_ts: Sequence[int], window: int, t_pad2: int) -> List[Segment]:
    """The (start, end) slices of vc_infer_pipeline.py:167-180 as an explicit list."""
    segs: List[Segment] = []
    s = 0
    t = None
    for i, t in enumerate(opt_ts):
        t = t // window * window
        end = t + t_pad2 + window
        segs.append(Segment(i, s, end, s // window, end // window))
        s = t
    start = t if t is not None else 0
    segs.append(Segment(len(segs), start, n_padded, start // window, 1 << 62))
    return segs


def hubert_frames(n_samples: int) -> int:
    """Frames a HuBERT/ContentVec front end emits for n 16 kHz samples (receptive field 400, hop 320)."""
    return (n_samples - 400) // 320 + 1


def assign_segments(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of segments to ranks (deterministic; ties by index)."""
    order = sorted(range(len(lengths)), key=lambda i: (-lengths[i], i))
    load = [0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += lengths[i]
    return [sorted(x) for x in out]


def makespan_bound(lengths: Sequence[int], world: int) -> float:
    """Best possible speed-up of this segment list on `world` ranks = Σlen / max rank load under LPT."""
    loads = [sum(lengths[i] for i in idx) for idx in assign_segments(lengths, world)]
    return sum(lengths) / max(max(loads), 1)


# ------------------------------------------------------------------------------------------------------------
class FeatureExtractor:
    """Constants and f0 post-processing of /root/reference/pitch_extraction.py:13-45, :252-304."""

    def __init__(self, tgt_sr, config, onnx: bool = False):
        self.x_pad, self.x_query, self.x_center, self.x_max, self.is_half = (
            config.x_pad, config.x_query, config.x_center, config.x_max, config.is_half)
        self.sr = 16000
        self.window = 160
        self.f0_bins = 256
        self.t_pad = self.sr * self.x_pad
        self.t_pad_tgt = tgt_sr * self.x_pad
        self.t_pad2 = self.t_pad * 2
        self.t_query = self.sr * self.x_query
        self.t_center = self.sr * self.x_center
        self.t_max = self.sr * self.x_max
        self.device = config.device
        self.onnx = onnx
        # the reference registers pm/harvest/dio/rmvpe/crepe here (pitch_extraction.py:28-45).  RMVPE -- the nodes' default
        # (custom_nodes/rvc_nodes.py:49) -- runs on the B200 kernels (comfy_rvc_b200/rmvpe.py); the CPU estimators (parselmouth,
        # pyworld, torchcrepe) are upstream of this path: register any callable `fn(x=, f0_up_key=, f0_min=, f0_max=, ...) -> f0[frames]`
        self.f0_method_dict = {"rmvpe": self.get_rmvpe, "rmvpe+": self.get_pitch_dependant_rmvpe}
        self.rmvpe_model_path = getattr(config, "rmvpe_model_path", None)

    def _rmvpe(self):
        """pitch_extraction.py:192-193: built on first use and kept; attach a ready model as `self.model_rmvpe` to share it."""
        if not hasattr(self, "model_rmvpe"):
            if not self.rmvpe_model_path:
                raise RuntimeError("f0_method 'rmvpe' needs `model_rmvpe` (a comfy_rvc_b200.RMVPE) or `config.rmvpe_model_path` (rmvpe.pt)")
            from .rmvpe import RMVPE
            self.model_rmvpe = RMVPE(self.rmvpe_model_path, is_half=self.is_half, device=self.device, onnx=self.onnx)
        return self.model_rmvpe

    def get_rmvpe(self, x, *args, **kwargs):
        """pitch_extraction.py:191-195."""
        return self._rmvpe().infer_from_audio(x, thred=0.03)

    def get_pitch_dependant_rmvpe(self, x, f0_min=0, f0_max=40000, *args, **kwargs):
        """pitch_extraction.py:197-201."""
        return self._rmvpe().infer_from_audio_with_pitch(x, thred=0.03, f0_min=f0_min, f0_max=f0_max)

    def load_index(self, file_index):
        """pitch_extraction.py:49-73: a preloaded `(index, big_npy)` tuple, "" for none, or a faiss file path."""
        index = big_npy = None
        try:
            if isinstance(file_index, tuple):
                index, big_npy = file_index
            elif file_index == "" or file_index is None:
                index = None
            else:
                import faiss  # not part of this image; same failure behaviour as the reference (prints, continues)
                index = faiss.read_index(file_index)
                big_npy = index.reconstruct_n(0, index.ntotal)
        except Exception as e:  # noqa: BLE001
            print(f"Could not open Faiss index file for reading. {e}")
        return index, big_npy

    def get_f0(self, x, f0_up_key, f0_method, merge_type="median", filter_radius=3, crepe_hop_length=160,
               f0_autotune=False, rmvpe_onnx=False, inp_f0=None, f0_min=50, f0_max=1100, **kwargs):
        """pitch_extraction.py:252-304 → (f0_coarse int16 [frames], f0 float [frames])."""
        time_step = self.window / self.sr * 1000
        f0_mel_min = hz_to_mel(f0_min)
        f0_mel_max = hz_to_mel(f0_max)
        params = {"x": x, "f0_up_key": f0_up_key, "f0_min": f0_min, "f0_max": f0_max, "time_step": time_step,
                  "filter_radius": filter_radius, "crepe_hop_length": crepe_hop_length, "model": "full", "onnx": rmvpe_onnx}
        if hasattr(f0_method, "pop") and len(f0_method) == 1:
            f0_method = f0_method.pop()
        if isinstance(f0_method, list):
            stack = [np.asarray(self._f0_fn(m)(**params), dtype=np.float64) for m in f0_method]
            n = max(len(s) for s in stack)
            stack = np.stack([np.pad(s, (0, n - len(s))) for s in stack])                 # lib/audio.py:257-262
            merge = {"min": np.nanmin, "max": np.nanmax, "median": np.nanmedian}.get(merge_type, np.nanmean)
            f0 = merge(stack, axis=0)                                                      # lib/utils.py:104-108
        else:
            f0 = self._f0_fn(f0_method)(**params)
        if f0_autotune:
            raise NotImplementedError("f0_autotune (lib/audio.py:274) is outside the synthesis hot path")
        f0 = np.array(f0, copy=True)
        f0 *= pow(2, f0_up_key / 12)
        tf0 = self.sr // self.window
        if inp_f0 is not None:                                                             # :282-292
            delta_t = np.round((inp_f0[:, 0].max() - inp_f0[:, 0].min()) * tf0 + 1).astype("int16")
            replace_f0 = np.interp(list(range(delta_t)), inp_f0[:, 0] * 100, inp_f0[:, 1])
            shape = f0[self.x_pad * tf0: self.x_pad * tf0 + len(replace_f0)].shape[0]
            f0[self.x_pad * tf0: self.x_pad * tf0 + len(replace_f0)] = replace_f0[:shape]
        f0_mel = hz_to_mel(f0)
        f0_mel = (f0_mel - f0_mel_min) * (self.f0_bins - 2) / (f0_mel_max - f0_mel_min) + 1
        f0_mel = np.clip(f0_mel, a_min=1, a_max=self.f0_bins - 1)
        f0_coarse = np.rint(f0_mel).astype(np.int16)
        return f0_coarse, f0

    def _f0_fn(self, method) -> Callable:
        if callable(method):
            return method
        if method not in self.f0_method_dict:
            raise Exception(f"Method {method} not found.")      # same message as pitch_extraction.py:229
        return self.f0_method_dict[method]


# ------------------------------------------------------------------------------------------------------------
class VC(FeatureExtractor):
    """Drop-in for /root/reference/vc_infer_pipeline.py:23 `class VC(FeatureExtractor)`.

    Extra keyword arguments (all optional, the reference call `VC(tgt_sr, config)` keeps working):
      noise : "device" (default) — per-segment CUDA generator seeded by (seed, segment index): results do not depend
              on how segments are sharded;  "reference" — the three draws of every segment come from the global
              torch CPU RNG in the reference's order (parity tests).
      group : torch.distributed process group to shard segments over (default: the world group when initialised).
    """

    def __init__(self, tgt_sr, config, onnx: bool = False, noise: str = "device", seed: int = 0, group=None):
        super().__init__(tgt_sr, config, onnx)
        if noise not in ("device", "reference"):
            raise ValueError(noise)
        self.noise_mode, self.seed, self.group = noise, int(seed), group
        self._host_group = None
        self._stage_buf: Optional[torch.Tensor] = None
        self._out_buf: Optional[torch.Tensor] = None
        self.last_plan: Optional[dict] = None

    # ---- one segment ------------------------------------------------------------------------------------
    def _torch_device(self) -> torch.device:
        dev = torch.device(self.device)
        if dev.type != "cuda":
            raise RuntimeError("comfy_rvc_b200.VC needs a CUDA device; the product path has no CPU fallback")
        return torch.device("cuda", dev.index if dev.index is not None else 0)

    def _segment_noise(self, net_g, i: int, T: int, staged):
        cfg = net_g.cfg
        dev = staged["dev"]
        if self.noise_mode == "reference":
            nz = torch.randn(1, cfg.inter_channels, T)          # models.py:685/801 (:908 for the no-f0 classes)
            if not cfg.f0:
                return (nz,)
            ri = torch.rand(1, 1)                               # models.py:378
            ns = torch.randn(1, T * cfg.upp, 1)                 # models.py:409
            return nz, ri, ns
        g = torch.Generator(device=dev).manual_seed(self.seed * 1000003 + i)
        nz = torch.randn(1, cfg.inter_channels, T, device=dev, generator=g)
        if not cfg.f0:
            return (nz,)
        ri = torch.rand(1, 1, device=dev, generator=g)
        ns = torch.randn(1, T * cfg.upp, 1, device=dev, generator=g)
        return nz, ri, ns

    def _vc_device(self, model, net_g, sid, audio0, n_samples: int, pitch, pitchf, index, big_npy, index_rate, version,
                   protect, noise=None) -> torch.Tensor:
        """`VC.vc` up to (not including) the D2H copy: returns the segment's PCM as a device tensor [L] fp32.
        `audio0` is a device tensor [n] (fp32 or fp16)."""
        dev = audio0.device
        lib = _lib.load()
        feats = audio0.view(1, -1)
        padding_mask = torch.zeros(feats.shape, dtype=torch.bool, device=dev)
        inputs = {"source": feats, "padding_mask": padding_mask, "output_layer": 9 if version == "v1" else 12}
        feats = model.extract_features(version=version, **inputs)                              # :48-55
        if callable(noise):      # "reference" noise: drawn AFTER the feature extractor, like models.py:801 after :48-55 -- HuggingFace's
            noise = noise()      # HuBERT consumes the global generator itself (one torch.rand([]) per encoder layer and forward)
        use_f0 = pitch is not None and pitchf is not None
        feats0 = feats if (protect < 0.5 and use_f0) else None                                 # :57-58 (no clone needed)
        if index is not None and big_npy is not None and index_rate > 0:                       # :59-75, host like the reference
            npy = feats[0].float().cpu().numpy()
            score, ix = index.search(npy, k=1)
            weight = np.square(1 / score)
            weight /= weight.sum(axis=1, keepdims=True)
            npy = np.sum(big_npy[ix] * np.expand_dims(weight, axis=2), axis=1)
            if self.is_half:
                npy = npy.astype("float16")
            feats = torch.from_numpy(npy).unsqueeze(0).to(dev) * index_rate + (1 - index_rate) * feats
        if feats.dtype not in (torch.float32, torch.float16):
            feats = feats.float()
            feats0 = feats0.float() if feats0 is not None else None
        feats = feats.contiguous()
        F_, Cf = int(feats.shape[1]), int(feats.shape[2])
        p_len = min(n_samples // self.window, 2 * F_)                                          # :83
        if use_f0:
            pitch = pitch[:, :p_len].contiguous()                                              # :86-87
            pitchf = pitchf[:, :p_len].contiguous()
            if pitch.shape[1] < p_len:
                raise ValueError(f"pitch has {pitch.shape[1]} frames, segment needs {p_len}")
        phone = torch.empty(1, p_len, Cf, device=dev, dtype=torch.float32)
        use_protect = 1 if feats0 is not None else 0
        f0p = feats0.contiguous() if feats0 is not None else feats
        if f0p.dtype != feats.dtype:
            f0p = f0p.to(feats.dtype)
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.rvcb200_op_prepare_feats(                                               # :77-95 in one kernel
            C.c_void_p(feats.data_ptr()), C.c_void_p(f0p.data_ptr()), 0 if feats.dtype == torch.float32 else 1,
            C.c_void_p(pitchf.data_ptr() if use_f0 else None), C.c_void_p(phone.data_ptr()), F_, p_len, Cf, float(protect),
            use_protect,
            C.c_void_p(stream)), None, "prepare_feats")
        p_len_t = torch.full((1,), p_len, device=dev, dtype=torch.int64)                       # :96 (no H2D copy)
        kw = {"noise": noise} if noise is not None else {}
        if use_f0:
            return net_g.infer(phone, p_len_t, pitch, pitchf, sid, **kw)[0][0, 0]              # :97-101
        return net_g.infer(phone, p_len_t, sid, **kw)[0][0, 0]                                 # :102-105

    def vc(self, model, net_g, sid, audio0, pitch, pitchf, times, index, big_npy, index_rate, version, protect):
        """Same contract as the reference `VC.vc` (vc_infer_pipeline.py:25-114): numpy in, float32 numpy PCM out."""
        dev = self._torch_device()
        feats = torch.from_numpy(np.ascontiguousarray(audio0))
        feats = feats.half() if self.is_half else feats.float()
        if feats.dim() == 2:
            feats = feats.mean(-1)
        assert feats.dim() == 1, feats.dim()
        with torch.cuda.device(dev):
            o = self._vc_device(model, net_g, sid, feats.to(dev), int(audio0.shape[0]), pitch, pitchf, index, big_npy,
                                index_rate, version, protect)
            return o.data.cpu().float().numpy()

    # ---- sharding helpers -----------------------------------------------------------------------------
    def _dist(self):
        import torch.distributed as dist
        if self.group is None and not (dist.is_available() and dist.is_initialized()):
            return None, 0, 1
        return dist, dist.get_rank(self.group), dist.get_world_size(self.group)

    def _gloo_group(self, dist):
        """Host-side gathers go through a gloo group (created once, collectively) when the data group is NCCL."""
        if self._host_group is None:
            if dist.get_backend(self.group) == "gloo":
                self._host_group = self.group if self.group is not None else dist.group.WORLD
            else:
                ranks = dist.get_process_group_ranks(self.group) if self.group is not None else None
                self._host_group = dist.new_group(ranks=ranks, backend="gloo")
        return self._host_group

    # ---- the song-level driver ---------------------------------------------------------------------------
    def plan(self, audio: np.ndarray):
        """Host-side planning shared by all ranks: high-pass, quiet points, padded audio, segment list."""
        audio = np.asarray(audio)
        n, h = audio.shape[0], self.window // 2
        if audio.ndim == 1 and n > max(self.t_pad, 18) and self.t_pad >= h:
            audio_pad = filtfilt_pad(audio, self.t_pad, out=self._staging(n + 2 * self.t_pad))  # :122 + :141, one pass
            audio = audio_pad[self.t_pad: self.t_pad + n]
            padded = audio_pad[self.t_pad - h: self.t_pad + n + h]
        else:
            audio = signal.filtfilt(_BH, _AH, audio)                                           # :122
            audio_pad = np.pad(audio, (self.t_pad, self.t_pad), mode="reflect")                # :141
            padded = None
        opt_ts = split_points(audio, self.window, self.t_query, self.t_center, self.t_max, padded)   # :123-135
        segs = plan_segments(audio_pad.shape[0], opt_ts, self.window, self.t_pad2)
        return audio, audio_pad, opt_ts, segs

    def _pinned_out(self, n: int) -> torch.Tensor:
        """Reusable pinned int16 buffer for the song's single D2H copy (cudaHostAlloc of ~60 MB costs ~10 ms per call)."""
        if self._out_buf is None or self._out_buf.numel() < n:
            self._out_buf = torch.empty(max(n, 1), dtype=torch.int16).pin_memory()
        return self._out_buf[:n]

    def _staging(self, n: int) -> np.ndarray:
        """Reusable float64 host buffer the filtered, padded song is written into: pinned when a CUDA device is present
        (it is the source of the song's one H2D copy), plain memory otherwise (host-only planning)."""
        if self._stage_buf is None or self._stage_buf.numel() < n:
            t = torch.empty(n, dtype=torch.float64)
            self._stage_buf = t.pin_memory() if torch.cuda.is_available() else t
        return self._stage_buf.numpy()[:n]

    def pipeline(self, model, net_g, sid, audio, times, f0_up_key, f0_method, merge_type, file_index, index_rate, if_f0,
                 filter_radius, tgt_sr, resample_sr, rms_mix_rate, version, protect, crepe_hop_length, f0_autotune,
                 rmvpe_onnx, f0_file=None, f0_min=50, f0_max=1600, all_ranks: bool = False):
        """Same contract as the reference `VC.pipeline` (vc_infer_pipeline.py:116-196) → int16 PCM at `tgt_sr`.

        Under `torch.distributed` every rank must call it with the same arguments; rank 0 (or every rank with
        `all_ranks=True`) returns the song, the others return None."""
        if rms_mix_rate < 1:
            raise NotImplementedError("rms_mix_rate < 1 (change_rms, lib/model_utils.py:39, needs librosa) is outside this path")
        if resample_sr >= 16000 and tgt_sr != resample_sr:
            raise NotImplementedError("resample_sr (librosa.resample, vc_infer_pipeline.py:185-186) is outside this path")
        dist, rank, world = self._dist()
        t0 = time.time()
        index, big_npy = self.load_index(file_index)
        audio, audio_pad, opt_ts, segs, staged_audio = self._plan_song(np.asarray(audio))
        t_plan = time.time()
        inp_f0 = None
        if f0_file is not None:                                                                # :144-149
            try:
                with open(f0_file.name, "r") as f:
                    inp_f0 = np.array([list(map(float, line.split(","))) for line in f.read().strip("\n").split("\n")],
                                      dtype="float32")
            except Exception:  # noqa: BLE001
                traceback.print_exc()
        pitch = pitchf = None
        if if_f0 == 1:                                                                         # :153-162
            pitch, pitchf = self.get_f0(audio_pad, f0_up_key, f0_method, merge_type, filter_radius, crepe_hop_length,
                                        f0_autotune, rmvpe_onnx, inp_f0, f0_min, f0_max)
            p_len = min(pitch.shape[0], pitchf.shape[0])
            pitch = pitch[:p_len].astype(np.int64)
            pitchf = pitchf[:p_len].astype(np.float32)
        if (pitch is None) != (not getattr(getattr(net_g, "cfg", None), "f0", True)):
            raise ValueError("if_f0 does not match the synthesizer class (f0 vs `_nono`)")
        lengths = [s.n_samples for s in segs]
        assignment = assign_segments(lengths, world)
        mine = assignment[rank]
        self.last_plan = {"opt_ts": list(opt_ts), "segments": segs, "assignment": assignment,
                          "makespan_bound": makespan_bound(lengths, world)}
        staged = self._stage(audio_pad, pitch, pitchf, sid, net_g, staged_audio)               # H2D once per song
        t1 = time.time()
        times[1] += t1 - t0                                                                    # :164-165
        ev0 = self._record_event(staged)
        parts = []
        for s in segs:
            T_formula = min(s.n_samples // self.window, 2 * hubert_frames(s.n_samples))
            run = s.index in mine
            noise = None
            if self.noise_mode == "reference":          # every rank advances the global stream past every segment, in the
                if run:                                 # reference's order: the front end's own draws, then the synthesizer's
                    noise = (lambda s=s, T=T_formula: self._segment_noise(net_g, s.index, T, staged))
                else:
                    for _ in range(int(getattr(model, "rng_draws_per_call", 0))):
                        torch.rand([])
                    self._segment_noise(net_g, s.index, T_formula, staged)
            elif run:
                noise = self._segment_noise(net_g, s.index, T_formula, staged)
            if run:
                parts.append((s.index, self._convert(staged, model, net_g, s, T_formula, index, big_npy, index_rate,
                                                     version, protect, noise)))
        t_enq = time.time()
        if world > 1 and self._device_collectives(dist, staged):
            # NCCL group: the peak is one 4-byte all-reduce and the int16 pieces travel GPU to GPU (NVLink); rank 0 orders
            # them on the device and does the song's single D2H -- no pickling, no host round trip per rank
            sizes = [min(s.n_samples // self.window, 2 * hubert_frames(s.n_samples)) * net_g.cfg.upp - 2 * self.t_pad_tgt for s in segs]
            out = self._finalize_gather_device(staged, parts, sizes, assignment, dist, rank, world, all_ranks)
            self._timing(staged, ev0, t0, t_plan, t1, t_enq)
            times[2] += time.time() - t1
            return out

        def exchange_peak(local_peak: float) -> float:                                         # the only reduction
            if world == 1:
                return local_peak
            peaks = [None] * world
            dist.all_gather_object(peaks, float(local_peak), group=self._gloo_group(dist))
            return max(peaks)

        mine_np = self._finalize(staged, parts, exchange_peak if world > 1 else None)          # :182-189
        self._timing(staged, ev0, t0, t_plan, t1, t_enq)
        if world > 1:
            gathered = [None] * world
            hg = self._gloo_group(dist)
            if all_ranks:
                dist.all_gather_object(gathered, mine_np, group=hg)
            else:
                dist.gather_object(mine_np, gathered if rank == 0 else None, dst=dist.get_global_rank(hg, 0), group=hg)
            if rank != 0 and not all_ranks:
                times[2] += time.time() - t1
                return None
            mine_np = {}
            for d in gathered:
                mine_np.update(d)
        out = np.concatenate([mine_np[i] for i in range(len(segs))])
        times[2] += time.time() - t1
        return out

    # ---- hooks with a host default (the CPU test double overrides the device ones below) ------------------------
    def _plan_song(self, audio: np.ndarray):
        """Planning for `pipeline`: like `plan`, but with a CUDA device the quiet-point search runs there on the staged
        float64 song (same sums in the same order, `rvcb200_op_quiet_point`), so the song crosses PCIe once, as float64,
        straight from the pinned buffer the filter wrote.  Returns (..., staged float64 device tensor or None)."""
        dev = torch.device(self.device)
        n, h = audio.shape[0], self.window // 2
        if (dev.type != "cuda" or not torch.cuda.is_available() or type(self)._stage is not VC._stage
                or not (audio.ndim == 1 and n > max(self.t_pad, 18) and self.t_pad >= h)):
            return (*self.plan(audio), None)
        dev = self._torch_device()
        lib = _lib.load()
        with torch.cuda.device(dev):
            audio_pad = filtfilt_pad(audio, self.t_pad, out=self._staging(n + 2 * self.t_pad))  # :122 + :141
            audio_d = torch.from_numpy(audio_pad).to(dev, non_blocking=True)
            opt_ts: List[int] = []
            if n + 2 * h > self.t_max:                                                         # :127
                centres = list(range(self.t_center, n, self.t_center))
                nb = 64
                bv = torch.empty(len(centres), nb, dtype=torch.float64, device=dev)
                bj = torch.empty(len(centres), nb, dtype=torch.int64, device=dev)
                stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                base = audio_d.data_ptr() + (self.t_pad - h) * 8                               # audio reflect-padded by window // 2
                for k, t in enumerate(centres):
                    lo, hi = t - self.t_query, min(t + self.t_query, n)
                    _lib.check(lib.rvcb200_op_quiet_point(C.c_void_p(base), lo, hi, self.window, C.c_void_p(bv[k].data_ptr()),
                                                          C.c_void_p(bj[k].data_ptr()), nb, stream), None, "quiet_point")
                bv_h, bj_h = bv.cpu().numpy(), bj.cpu().numpy()                                # one small D2H (sync)
                for k in range(len(centres)):
                    best_v, best_j = np.inf, -1
                    for b in range(nb):                                                        # ascending ranges: first minimum
                        if bj_h[k, b] >= 0 and bv_h[k, b] < best_v:
                            best_v, best_j = bv_h[k, b], int(bj_h[k, b])
                    if best_j < 0:                                                             # every sum NaN: numpy's argmin rule
                        best_j = centres[k] - self.t_query
                    opt_ts.append(best_j)
        segs = plan_segments(audio_pad.shape[0], opt_ts, self.window, self.t_pad2)
        return audio_pad[self.t_pad: self.t_pad + n], audio_pad, opt_ts, segs, audio_d

    # ---- device side of the song-level driver (everything below touches the GPU) -----------------------------
    def _record_event(self, staged):
        if staged["dev"].type != "cuda":
            return None
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(staged["dev"]))
        return ev

    def _timing(self, staged, ev0, t0, t_plan, t_stage, t_enq):
        """Where the call's time went: host phases by wall clock, the enqueued device work by CUDA events."""
        ev1 = self._record_event(staged)
        if ev1 is not None:
            ev1.synchronize()
        self.last_plan["host_s"] = {"plan": t_plan - t0, "f0_and_stage": t_stage - t_plan, "enqueue": t_enq - t_stage,
                                    "total": time.time() - t0}
        self.last_plan["device_ms"] = ev0.elapsed_time(ev1) if ev0 is not None else None

    def _device_collectives(self, dist, staged) -> bool:
        return staged["dev"].type == "cuda" and dist.get_backend(self.group) == "nccl"

    def _stage(self, audio_pad, pitch, pitchf, sid, net_g, staged_audio=None) -> dict:
        dev = self._torch_device()
        with torch.cuda.device(dev):
            # the whole song, once: float64 from the (pinned) staging buffer, rounded to the model dtype on the device --
            # the same single rounding as the reference's `torch.from_numpy(audio0).half() / .float()` (:40-44)
            audio_d = staged_audio if staged_audio is not None else torch.from_numpy(audio_pad).to(dev, non_blocking=True)
            return {
                "dev": dev,
                "audio": audio_d.half() if self.is_half else audio_d.float(),
                "sid": torch.as_tensor(sid).reshape(1).to(dev).long(),                         # :151
                "pitch": torch.from_numpy(pitch).to(dev).unsqueeze(0) if pitch is not None else None,
                "pitchf": torch.from_numpy(pitchf).to(dev).unsqueeze(0) if pitchf is not None else None,
            }

    def _convert(self, staged, model, net_g, s: Segment, T_formula: int, index, big_npy, index_rate, version, protect, noise):
        """One segment, enqueued without a host round trip; returns the trimmed PCM as a device tensor."""
        dev = staged["dev"]
        with torch.cuda.device(dev):
            o = self._vc_device(model, net_g, staged["sid"], staged["audio"][s.start:s.end], s.n_samples,
                                staged["pitch"][:, s.f0_start:s.f0_end] if staged["pitch"] is not None else None,
                                staged["pitchf"][:, s.f0_start:s.f0_end] if staged["pitchf"] is not None else None,
                                index, big_npy, index_rate, version, protect, noise=noise)
            if self.noise_mode == "reference" and o.shape[0] != T_formula * net_g.cfg.upp:
                raise RuntimeError("HuBERT front end does not follow the 400/320 frame formula; reference-order noise "
                                   "cannot be pre-drawn")
            return o[self.t_pad_tgt: o.shape[0] - self.t_pad_tgt]                              # :174/:180 trim, on device

    def _finalize_gather_device(self, staged, parts, sizes_h, assignment, dist, rank, world, all_ranks):
        """`_finalize` + the gather for an NCCL group, all on the device: peak = all-reduce(max) of one float; every rank
        converts its segments to int16 against the song-wide peak; the pieces go to rank 0 (to every rank with
        `all_ranks`) as bytes over NVLink; the destination orders them by segment index and copies the song to the host
        once.  No collective touches the synthesis itself."""
        dev = staged["dev"]
        lib = _lib.load()
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            local = torch.cat([p for _, p in parts]) if parts else torch.empty(0, device=dev)
            peak = torch.zeros(1, device=dev, dtype=torch.float32)
            _lib.check(lib.rvcb200_op_absmax(C.c_void_p(local.data_ptr()), local.numel(), C.c_void_p(peak.data_ptr()), 1,
                                             stream), None, "absmax")
            dist.all_reduce(peak, op=dist.ReduceOp.MAX, group=self.group)                      # :188, the only reduction
            for i, p in parts:           # every rank derives all piece sizes from the plan (400 / 320 frame formula of the front end)
                if int(p.numel()) != sizes_h[i]:
                    raise RuntimeError("HuBERT front end does not follow the 400/320 frame formula; segment sizes cannot be "
                                       "derived from the plan for the device-side gather")
            pcm = torch.empty(local.numel(), device=dev, dtype=torch.int16)
            _lib.check(lib.rvcb200_op_to_int16(C.c_void_p(local.data_ptr()), local.numel(), C.c_void_p(peak.data_ptr()),
                                               C.c_void_p(pcm.data_ptr()), stream), None, "to_int16")
            per_rank = [sum(sizes_h[i] for i in assignment[r]) for r in range(world)]
            dst_ranks = list(range(world)) if all_ranks else [0]
            bufs = {}
            ops = []
            grp = self.group
            g = (lambda r: dist.get_global_rank(grp, r)) if grp is not None else (lambda r: r)
            mine = pcm.view(torch.uint8)
            if rank in dst_ranks:
                for r in range(world):
                    if r == rank:
                        bufs[r] = mine
                    elif per_rank[r]:
                        bufs[r] = torch.empty(2 * per_rank[r], device=dev, dtype=torch.uint8)
                        ops.append(dist.P2POp(dist.irecv, bufs[r], g(r), group=grp))
            if mine.numel():
                for d in dst_ranks:
                    if d != rank:
                        ops.append(dist.P2POp(dist.isend, mine, g(d), group=grp))
            if ops:
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
            if rank not in dst_ranks:
                torch.cuda.current_stream(dev).synchronize()
                return None
            offs = np.concatenate([[0], np.cumsum(sizes_h)]).astype(np.int64)
            song = torch.empty(int(offs[-1]), device=dev, dtype=torch.int16)
            for r in range(world):
                o = 0
                for i in assignment[r]:
                    n = sizes_h[i]
                    song[offs[i]: offs[i] + n] = bufs[r].view(torch.int16)[o: o + n]
                    o += n
            out_h = self._pinned_out(song.numel())
            out_h.copy_(song, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        return out_h.numpy().copy()              # the caller owns the result; the pinned buffer is reused by the next song

    def _finalize(self, staged, parts, exchange_peak) -> dict:
        """Concatenate this rank's trimmed segments, peak-normalise against the song-wide max and convert to int16 on the
        device (vc_infer_pipeline.py:182-189); one D2H.  Returns {segment index: int16 array}."""
        dev = staged["dev"]
        lib = _lib.load()
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            local = torch.cat([p for _, p in parts]) if parts else torch.empty(0, device=dev)
            peak = torch.zeros(1, device=dev, dtype=torch.float32)
            _lib.check(lib.rvcb200_op_absmax(C.c_void_p(local.data_ptr()), local.numel(), C.c_void_p(peak.data_ptr()), 1,
                                             stream), None, "absmax")
            if exchange_peak is not None:
                peak.fill_(exchange_peak(float(peak.item())))
            pcm = torch.empty(local.numel(), device=dev, dtype=torch.int16)
            _lib.check(lib.rvcb200_op_to_int16(C.c_void_p(local.data_ptr()), local.numel(), C.c_void_p(peak.data_ptr()),
                                               C.c_void_p(pcm.data_ptr()), stream), None, "to_int16")
            out_h = self._pinned_out(pcm.numel())
            out_h.copy_(pcm, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        out = out_h.numpy()
        res, off = {}, 0
        for i, p in parts:
            n = int(p.numel())
            res[i] = out[off:off + n].copy()
            off += n
        return res


def get_vc(model_path, file_index=None, config=None, device=None):
    """Reference `get_vc` (vc_infer_pipeline.py:198-249) on the B200 classes: checkpoint → {vc, cpt, net_g, ...}."""
    import os
    from .synthesizer import (SynthesizerTrnMs256NSFsid, SynthesizerTrnMs768NSFsid, SynthesizerTrnMs256NSFsid_nono,
                              SynthesizerTrnMs768NSFsid_nono)
    config = config or PipelineConfig()
    cpt = torch.load(model_path, map_location="cpu")
    tgt_sr = cpt["config"][-1]
    cpt["config"][-3] = cpt["weight"]["emb_g.weight"].shape[0]       # n_spk
    if_f0 = cpt.get("f0", 1)
    version = cpt.get("version", "v1")
    if if_f0 == 1:                                                     # the 4-way dispatch of vc_infer_pipeline.py:205-218
        cls = SynthesizerTrnMs256NSFsid if version == "v1" else SynthesizerTrnMs768NSFsid
    else:
        cls = SynthesizerTrnMs256NSFsid_nono if version == "v1" else SynthesizerTrnMs768NSFsid_nono
    net_g = cls(*cpt["config"], is_half=config.is_half)
    del net_g.enc_q
    net_g.load_state_dict(cpt["weight"], strict=False)
    net_g.eval().to(device if device else config.device)
    net_g = net_g.half() if config.is_half else net_g.float()
    vc = VC(tgt_sr, config)
    model_name = os.path.basename(model_path).split(".")[0]
    if isinstance(file_index, str) and file_index and os.path.exists(file_index):
        file_index = vc.load_index(file_index)
        if file_index[0] is None:
            file_index = ""
    elif not isinstance(file_index, tuple):
        file_index = ""
    return {"vc": vc, "cpt": cpt, "net_g": net_g, "model_name": model_name, "file_index": file_index, "sr": cpt["config"][-1]}
