"""RMVPE f0 estimator on the B200 kernels (the other model in front of the synthesis path).

Drop-in for the reference's `RMVPE` class (/root/reference/lib/rmvpe.py:559-684) as `FeatureExtractor.get_rmvpe` /
`get_pitch_dependant_rmvpe` drive it (/root/reference/pitch_extraction.py:191-201):

    model = RMVPE(model_path, is_half=..., device="cuda:0")     # model_path: rmvpe.pt (a state_dict of E2E(4, 1, (2, 2))) or a dict
    f0 = model.infer_from_audio(audio16k, thred=0.03)           # numpy float64 [n // 160 + 1], Hz, 0 = unvoiced

Host code is Python like the reference's; every FLOP runs in librvcb200.so through the C ABI (include/rvcb200.h):
  * log-mel (`MelSpectrogram`, rmvpe.py:476-556) + the model's input BatchNorm + `mel2hidden`'s reflect padding of the frame
    axis: `rvcb200_op_rmvpe_logmel` (one FFT per frame in shared memory) -> fp16 image;
  * DeepUnet (rmvpe.py:232-428): images are channels-last fp16, a line = W pixels + zero padding, so a 3 x 3 convolution
    is a 9-tap row-offset contraction on the generic tcgen05 implicit-GEMM kernel (`rvcb200_op_conv_tc`, `tap_w` / `dil2`).  On the
    two wide levels (16 / 32 channels) `pack` = 4 / 2 neighbouring pixels form ONE 64-channel GEMM row and the weights become
    block-Toeplitz (out pixel j of a row reads in pixel i of the rows left / same / right with kernel column 4 fdx + i - j): a TMA box
    row then carries 128 useful bytes instead of 32 and a tile covers 512 instead of 128 pixels (L0 convolutions 97 -> ~35 us).  The
    eval-mode BatchNorms are folded into the convolution weights and biases at load, ReLU and the residual add sit in the epilogue,
    the residual stream stays fp32, `torch.cat` is a channel offset into one buffer, ConvTranspose2d(stride 2) is a 2 x 2-tap GEMM
    over (phase, channel) columns + `rvcb200_op_rmvpe_shuffle`, AvgPool2d is `rvcb200_op_rmvpe_pool`;
  * BiGRU (rmvpe.py:217-229): input projection on the tensor cores, recurrence in `rvcb200_op_rmvpe_gru` (8-CTA cluster per
    direction, fp32 W_hh in registers, state in distributed shared memory);
  * Linear + Sigmoid + `to_local_average_cents` + `decode` (rmvpe.py:451-456, 610-615, 658-684): tensor-core GEMM +
    `rvcb200_op_rmvpe_decode` (float64 like the reference's numpy).
There is no PyTorch or CPU fallback: without the extension or a CUDA device every entry point raises.
`is_half` is accepted and ignored: operands are fp16 and accumulation fp32 in both modes.  `onnx=True` is not supported.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Union

import numpy as np
import torch

from . import _lib
from .weights import pack_tc

PADF = 32
N_FFT, HOP, N_MELS, N_BINS, SR, FMIN, FMAX = 1024, 160, 128, 513, 16000, 30.0, 8000.0
N_CLASS, N_CLASS_PAD = 360, 384
BN_EPS = 1e-5
ROW_C = 64                    # channels of a GEMM row on the wide levels: `pack` = ROW_C / C pixels per row


def mel_filterbank(sr=SR, n_fft=N_FFT, n_mels=N_MELS, fmin=FMIN, fmax=FMAX) -> np.ndarray:
    """The table `librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=True)` returns (Slaney-normalised triangles on the HTK mel
    scale, float32 [n_mels, 1 + n_fft / 2]); rmvpe.py:504-511."""
    hz_to_mel = lambda f: 2595.0 * np.log10(1.0 + np.asarray(f, dtype=np.float64) / 700.0)
    mel_to_hz = lambda m: 700.0 * (10.0 ** (np.asarray(m, dtype=np.float64) / 2595.0) - 1.0)
    freqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sr)
    edges = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    width = np.diff(edges)
    slope = edges[:, None] - freqs[None, :]
    out = np.zeros((n_mels, freqs.shape[0]), dtype=np.float32)
    for m in range(n_mels):
        out[m] = np.maximum(0, np.minimum(-slope[m] / width[m], slope[m + 2] / width[m + 1]))
    out *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return out


class RMVPE:
    """See module docstring."""

    def __init__(self, model_path: Union[str, Dict[str, torch.Tensor]], is_half: bool = False, onnx: bool = False, device=None):
        if onnx:
            raise NotImplementedError("the onnxruntime / DirectML variant (rmvpe.py:571-577) is outside this path")
        self.is_half, self.onnx = is_half, False
        if device is None:
            device = "cuda"
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", 0)
        sd = torch.load(model_path, map_location="cpu") if isinstance(model_path, (str, bytes)) or hasattr(model_path, "__fspath__") \
            else model_path
        self._sd = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items() if not k.endswith("num_batches_tracked")}
        need = ["unet.encoder.bn.weight", "unet.encoder.layers.0.conv.0.conv.0.weight", "cnn.weight", "fc.0.gru.weight_hh_l0_reverse",
                "fc.1.weight"]
        missing = [k for k in need if k not in self._sd]
        if missing:
            raise RuntimeError(f"missing keys in state_dict: {missing}")
        # architecture as rmvpe.py:578 builds it -- E2E(4, 1, (2, 2)); sizes are read off the weights
        self.n_blocks = 1 + max(int(k.split(".")[5]) for k in self._sd if k.startswith("unet.encoder.layers.0.conv."))
        self.n_levels = 1 + max(int(k.split(".")[3]) for k in self._sd if k.startswith("unet.encoder.layers."))
        self.n_inter = 1 + max(int(k.split(".")[3]) for k in self._sd if k.startswith("unet.intermediate.layers."))
        self.c0 = self._sd["unet.encoder.layers.0.conv.0.conv.0.weight"].shape[0]
        self.gru_h = self._sd["fc.0.gru.weight_hh_l0"].shape[1]
        ok = (self.c0 % 16 == 0 and self.n_levels == 5 and self.gru_h == 256 and self._sd["fc.1.weight"].shape == (N_CLASS, 512)
              and self._sd["fc.0.gru.weight_ih_l0"].shape == (768, 3 * N_MELS) and self._sd["cnn.weight"].shape[0] == 3
              and self._sd["unet.encoder.layers.0.conv.0.conv.0.weight"].shape[1] == 1)
        if not ok:
            raise ValueError("RMVPE (B200) is built for E2E(4, 1, (2, 2)): 1-channel 128-bin input, 5 levels, BiGRU(384, 256), 360 classes")
        self.cents_mapping = np.pad(20 * np.arange(N_CLASS) + 1997.3794084376191, (4, 4))        # rmvpe.py:588-589
        self._w: Optional[Dict[str, torch.Tensor]] = None
        self._src: Dict[str, torch.Tensor] = {}
        self.last_launches = 0

    # ---- weights -> device images ------------------------------------------------------------------------------------
    def _fold_bn(self, conv_w: torch.Tensor, p_bn: str, transposed: bool = False):
        """conv (no bias) -> eval-mode BatchNorm2d: w' = w * s[co], b' = beta - mean * s[co], s = gamma / sqrt(var + eps)."""
        sd = self._sd
        s = sd[p_bn + "weight"].double() / torch.sqrt(sd[p_bn + "running_var"].double() + BN_EPS)
        b = sd[p_bn + "bias"].double() - sd[p_bn + "running_mean"].double() * s
        w = conv_w.double() * (s[None, :, None, None] if transposed else s[:, None, None, None])
        return w.float(), b.float()

    def _materialize(self):
        if self._w is not None:
            return
        if not torch.cuda.is_available() or self.device.type != "cuda":
            raise RuntimeError("comfy_rvc_b200.RMVPE needs a CUDA (sm_100a) device; it has no CPU fallback")
        _lib.load()
        if not self._src:
            self._build_host()
        self._w = {k: v.to(self.device).contiguous() for k, v in self._host.items()}

    # ---- image geometry ----------------------------------------------------------------------------------------------------
    def _pack(self, level: int) -> int:
        """Pixels per GEMM row at a level (levels 0.. = encoder / decoder, n_levels = intermediate)."""
        c = self.c0 << level
        return ROW_C // c if c < ROW_C else 1

    @property
    def _img_c(self) -> int:
        """Channels of the input image (1 real + zeros): one GEMM row of level 0 has ROW_C of them (>= 8: 16-byte TMA rows)."""
        return max(8, ROW_C // self._pack(0)) if self._pack(0) > 1 else 8

    def _geom(self, level: int, Tp: int):
        """(C, lines H, pixels per line W, pack, pixel pitch P = W + pack, rows per line FP = W / pack + 1, rows)."""
        c, Wd, H, pk = self.c0 << level, N_MELS >> level, Tp >> level, self._pack(level)
        return c, H, Wd, pk, Wd + pk, Wd // pk + 1, H * (Wd // pk + 1)

    def _build_host(self):
        """Weight-norm-free model: fold the BatchNorms, re-lay the kernels as tap matrices (CPU tensors; `_materialize` uploads)."""
        sd = self._sd
        W: Dict[str, torch.Tensor] = {}
        f32 = lambda t: t.to(torch.float32).contiguous()

        def krows(i, ci, pack, cat_c):
            """GEMM-row channel indices of pixel i's ci input channels.  Plain images: pixel-major.  The decoder's concat buffer
            keeps the up-sampled and the skip half of a row side by side ([pack x C up | pack x C skip]) so that the encoder's
            epilogue can write its half as one column range."""
            if cat_c is None:
                return i * ci + torch.arange(ci)
            c = torch.arange(ci)
            return (c // cat_c) * pack * cat_c + i * cat_c + (c % cat_c)

        def conv3(name, w, b, pack=1, cat_c=None):                # [Co][Ci][3][3] -> taps [9][pack Ci][pack Co]
            co, ci = w.shape[:2]
            if ci == 1:                                           # the 1-channel input image is stored with _img_c channels
                w = torch.cat([w, w.new_zeros(co, self._img_c - 1, 3, 3)], dim=1)
                ci = self._img_c
            if co % 16:                                           # cnn: 3 output channels -> 16 columns
                w = torch.cat([w, w.new_zeros(16 - co % 16, ci, 3, 3)], dim=0)
                b = torch.cat([b, b.new_zeros(16 - co % 16)])
                co = w.shape[0]
            t = torch.zeros(9, pack * ci, pack * co)
            for ky in range(3):
                for fdx in (-1, 0, 1):                            # GEMM row to the left / same / right
                    for i in range(pack):                         # input pixel inside that row
                        for j in range(pack):                     # output pixel inside this row
                            dx = pack * fdx + i - j
                            if -1 <= dx <= 1:
                                t[ky * 3 + fdx + 1, krows(i, ci, pack, cat_c), j * co:(j + 1) * co] = w[:, :, ky, dx + 1].t()
            self._src[name + ".w"], W[name + ".b"] = t.contiguous(), f32(b.repeat(pack))

        def conv1(name, w, b, pack=1, cat_c=None):                # 1 x 1 shortcut: [Co][Ci][1][1] -> [1][pack Ci][pack Co]
            co, ci = w.shape[:2]
            if ci == 1:
                w = torch.cat([w, w.new_zeros(co, self._img_c - 1, 1, 1)], dim=1)
                ci = self._img_c
            t = torch.zeros(1, pack * ci, pack * co)
            for i in range(pack):
                t[0, krows(i, ci, pack, cat_c), i * co:(i + 1) * co] = w[:, :, 0, 0].t()
            self._src[name + ".w"], W[name + ".b"] = t.contiguous(), f32(b.repeat(pack))

        def block(p, pack=1, cat_c=None):                         # ConvBlockRes, rmvpe.py:232-267
            w1, b1 = self._fold_bn(sd[p + "conv.0.weight"], p + "conv.1.")
            w2, b2 = self._fold_bn(sd[p + "conv.3.weight"], p + "conv.4.")
            conv3(p + "c1", w1, b1, pack, cat_c)
            conv3(p + "c2", w2, b2, pack)
            if p + "shortcut.weight" in sd:
                conv1(p + "sc", sd[p + "shortcut.weight"], sd[p + "shortcut.bias"], pack, cat_c)

        for i in range(self.n_levels):
            for j in range(self.n_blocks):
                block(f"unet.encoder.layers.{i}.conv.{j}.", self._pack(i))
        for i in range(self.n_inter):
            for j in range(self.n_blocks):
                block(f"unet.intermediate.layers.{i}.conv.{j}.")
        for i in range(self.n_levels):
            p = f"unet.decoder.layers.{i}."
            wt, bt = self._fold_bn(sd[p + "conv1.0.weight"], p + "conv1.1.", transposed=True)     # [Ci][Co][3][3]
            ci, co = wt.shape[:2]
            # out[2y + py][2x + px] = sum over input offsets (dy, dx) in {0, 1}^2 of in[y + dy][x + dx] . w[:, :, py + 1 - 2 dy, px + 1 - 2 dx]
            t = torch.zeros(4, ci, 4 * co)
            for dy in range(2):
                for dx in range(2):
                    for py in range(2):
                        for px in range(2):
                            ky, kx = py + 1 - 2 * dy, px + 1 - 2 * dx
                            if 0 <= ky <= 2 and 0 <= kx <= 2:
                                t[dy * 2 + dx, :, (py * 2 + px) * co:(py * 2 + px + 1) * co] = wt[:, :, ky, kx]
            self._src[p + "up.w"], W[p + "up.b"] = t, f32(bt.repeat(4))
            lvl = self.n_levels - 1 - i
            for j in range(self.n_blocks):
                block(p + f"conv2.{j}.", self._pack(lvl), cat_c=co if j == 0 else None)
        conv3("cnn", sd["cnn.weight"], sd["cnn.bias"], self._pack(0))
        g = "fc.0.gru."
        self._src["gru.ih.w"] = torch.cat([sd[g + "weight_ih_l0"].t(), sd[g + "weight_ih_l0_reverse"].t()], dim=1)[None].contiguous()
        W["gru.ih.b"] = f32(torch.cat([sd[g + "bias_ih_l0"], sd[g + "bias_ih_l0_reverse"]]))
        W["gru.hh.w"] = f32(torch.stack([sd[g + "weight_hh_l0"], sd[g + "weight_hh_l0_reverse"]]))      # [2][768][256]
        W["gru.hh.b"] = f32(torch.stack([sd[g + "bias_hh_l0"], sd[g + "bias_hh_l0_reverse"]]))
        wfc = torch.zeros(1, 2 * self.gru_h, N_CLASS_PAD)
        wfc[0, :, :N_CLASS] = sd["fc.1.weight"].t()
        self._src["fc.w"] = wfc
        W["fc.b"] = f32(torch.cat([sd["fc.1.bias"], torch.zeros(N_CLASS_PAD - N_CLASS)]))
        # input BatchNorm (1 channel) as an affine on the log-mel
        s0 = float(sd["unet.encoder.bn.weight"][0].double() / math.sqrt(float(sd["unet.encoder.bn.running_var"][0]) + BN_EPS))
        self._bn_scale = s0
        self._bn_shift = float(sd["unet.encoder.bn.bias"][0]) - float(sd["unet.encoder.bn.running_mean"][0]) * s0
        # mel front-end tables
        try:
            from scipy.signal import get_window
            win = get_window("hann", N_FFT, fftbins=True)
        except Exception:  # noqa: BLE001  (scipy is the reference's own dependency; same table either way to float32)
            win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(N_FFT) / N_FFT)
        W["mel.window"] = torch.from_numpy(np.asarray(win, dtype=np.float32))
        ang = 2.0 * np.pi * np.arange(N_FFT // 2) / N_FFT
        W["mel.twiddle"] = torch.from_numpy(np.stack([np.cos(ang), -np.sin(ang)], axis=1).astype(np.float32)).contiguous()
        fb = mel_filterbank()
        W["mel.basis"] = torch.from_numpy(fb).contiguous()
        rng = np.zeros((N_MELS, 2), dtype=np.int32)
        for m in range(N_MELS):
            nz = np.nonzero(fb[m])[0]
            rng[m] = (nz[0], nz[-1] + 1) if nz.size else (0, 0)
        W["mel.range"] = torch.from_numpy(rng).contiguous()
        self._host = W

    def _wimg(self, name: str, n_tile: int) -> torch.Tensor:
        key = f"{name}@{n_tile}"
        if key not in self._w:
            self._w[key] = pack_tc(self._src[name], torch.float16, n_tile).to(self.device)
        return self._w[key]

    # ---- reference object protocol ---------------------------------------------------------------------------------------
    def to(self, device=None, *a, **k):
        if device is not None and not isinstance(device, torch.dtype):
            dev = torch.device(device)
            if dev.type == "cuda" and dev.index is None:
                dev = torch.device("cuda", 0)
            if dev != self.device:
                self.device, self._w = dev, None
        return self

    # ---- one generic tcgen05 contraction over an image ---------------------------------------------------------------------
    @staticmethod
    def _n_tile(rows: int, cout: int) -> int:
        """Widest N in {256, 128, 64} that still gives a full wave of 128-row tiles; narrow layers use N = C_out."""
        if cout <= 64:
            return cout
        mt = (rows + 127) // 128
        for cand in (256, 128):
            if cout % cand == 0 and mt * (cout // cand) >= 132:
                return cand
        return 64

    def _conv(self, x16: int, rows: int, cin: int, ldx: int, name: str, cout: int, W: int, *, taps: int, relu: bool,
              y16: int = 0, ldy16: int = 0, y32: int = 0, ldy32: int = 0, res32: int = 0, ldr32: int = 0, mask: bool = True):
        """taps = 9: 3 x 3 "same" convolution, 4: the 2 x 2 phase taps of a transposed convolution, 1: 1 x 1; over an image of
        `rows` = H * (W + 1) pixels.  epilogue: + bias -> ReLU (optional) -> + res32 -> y32 fp32 and / or y16 fp16; pad pixels -> 0."""
        Wp = W + 1
        n_tile = self._n_tile(rows, cout)
        d = _lib.TcConvDesc()
        d.x16, d.L_in, d.padf = x16, rows, PADF
        d.w16, d.bias = self._wimg(name + ".w", n_tile).data_ptr(), self._w[name + ".b"].data_ptr()
        d.Cin, d.ntaps, d.dil, d.G = cin, taps, 1, 1
        if taps == 9:
            d.tap_w, d.dil2, d.g_off[0] = 3, Wp, -(Wp + 1)
        elif taps == 4:
            d.tap_w, d.dil2, d.g_off[0] = 2, Wp, 0
        else:
            d.tap_w, d.dil2, d.g_off[0] = 0, 0, 0
        halo = (taps // d.tap_w - 1) * Wp + (d.tap_w - 1) if d.tap_w else 0
        d.a_mode = 2 if halo > 127 else 0          # wide images: one activation box per kernel row (conv_tc.cu, a_mode 2)
        d.N, d.Cout_total = n_tile, cout
        d.Lj, d.out_stride, d.Lp_out = rows, 1, ((rows + 127) // 128) * 128 + 128
        d.div, d.out_slope, d.alpha = 1.0, 1.0, 1.0
        d.pre_slope = 0.0 if relu else 1.0
        d.generic, d.f32_cl, d.ldx16 = 1, 1, ldx
        if mask:
            d.pad_period, d.pad_valid, d.mask_post = Wp, W, 1
        if y16:
            d.y16, d.ldy16 = y16, ldy16
        if y32:
            d.y32, d.ldy32 = y32, ldy32
        if res32:
            d.res32, d.ldr32, d.res_mode = res32, ldr32, 1
        st = _lib.load().rvcb200_op_conv_tc(C.byref(d), 1, self._stream)
        if st != 0:
            raise RuntimeError(f"rvcb200_op_conv_tc failed with status {st} ({name}: Cin={cin}, Cout={cout}, taps={taps}, rows={rows}, W={W})")
        self.last_launches += 1

    def _check(self, st: int, what: str):
        if st != 0:
            raise RuntimeError(f"{what} failed with status {st} ({_lib.STATUS.get(st, '?')})")
        self.last_launches += 1

    def _block(self, p: str, x16: int, ldx: int, cin: int, x32: Optional[torch.Tensor], cout: int, H: int, W: int,
               out16: int = 0, ld_out16: int = 0):
        """ConvBlockRes (rmvpe.py:232-267).  Returns (y32 tensor [rows][cout], y16 pointer, its row stride, owner of that memory).
        Everything runs on one stream, so the caching allocator may hand a dropped tensor's memory to the next `torch.empty`
        without a hazard -- but a tensor whose POINTER is still going to be passed to a later launch must stay referenced."""
        dev, rows = self.device, H * (W + 1)
        if (p + "sc.w") in self._src:
            res = torch.empty(rows, cout, dtype=torch.float32, device=dev)
            self._conv(x16, rows, cin, ldx, p + "sc", cout, W, taps=1, relu=False, y32=res.data_ptr(), ldy32=cout, mask=False)
        else:
            res = x32
        h16 = torch.empty(rows, cout, dtype=torch.float16, device=dev)
        self._conv(x16, rows, cin, ldx, p + "c1", cout, W, taps=9, relu=True, y16=h16.data_ptr(), ldy16=cout)
        y32 = torch.empty(rows, cout, dtype=torch.float32, device=dev)
        y16t = None
        if not out16:
            y16t = torch.empty(rows, cout, dtype=torch.float16, device=dev)
            out16, ld_out16 = y16t.data_ptr(), cout
        self._conv(h16.data_ptr(), rows, cout, cout, p + "c2", cout, W, taps=9, relu=True, y16=out16, ldy16=ld_out16,
                   y32=y32.data_ptr(), ldy32=cout, res32=res.data_ptr(), ldr32=cout)
        return y32, out16, ld_out16, y16t

    # ---- the model -------------------------------------------------------------------------------------------------------
    def _hidden_from_img(self, img: torch.Tensor, Tp: int, taps: Optional[dict] = None):
        """img fp16 [Tp][P0][img_c] -> (logits fp32 [Tp][384] on the device).  Tp a multiple of 32.
        Every image is [H][P pixels][C] in memory; the convolutions see it as [H][FP rows][pack C] (see _geom)."""
        lib, Wt, dev = _lib.load(), self._w, self.device
        cats = []
        own = img                               # owner of the memory behind x16 (see _block)
        x16, x32 = img.data_ptr(), None
        cin_pix = self._img_c                   # channels per pixel of the tensor behind x16
        for i in range(self.n_levels):                                        # Encoder, rmvpe.py:296-305
            c, H, Wd, pk, P, FP, rows = self._geom(i, Tp)
            cf = pk * c
            cat = torch.zeros(rows, 2 * cf, dtype=torch.float16, device=dev)  # per row: [pk x c up-sampled | pk x c skip]
            cats.append(cat)
            for j in range(self.n_blocks):
                last = j == self.n_blocks - 1
                x32, x16, ldx, own = self._block(f"unet.encoder.layers.{i}.conv.{j}.", x16, pk * cin_pix, pk * cin_pix, x32, cf, H, FP - 1,
                                                 out16=cat.data_ptr() + 2 * cf if last else 0, ld_out16=2 * cf if last else 0)
                cin_pix = c
            _, H2, W2, _, P2, _, _ = self._geom(i + 1, Tp)
            pooled = torch.empty(H2 * P2, c, dtype=torch.float16, device=dev)
            self._check(lib.rvcb200_op_rmvpe_pool(C.c_void_p(x32.data_ptr()), c, C.c_void_p(pooled.data_ptr()), H2, W2, c, P, P2,
                                                  self._stream), "rvcb200_op_rmvpe_pool")
            x16, x32, own = pooled.data_ptr(), None, pooled
        c, H, Wd, pk, P, FP, rows = self._geom(self.n_levels, Tp)
        if taps is not None:
            taps["enc16"] = pooled
        for i in range(self.n_inter):                                         # Intermediate, rmvpe.py:343-347
            for j in range(self.n_blocks):
                x32, x16, ldx, own = self._block(f"unet.intermediate.layers.{i}.conv.{j}.", x16, cin_pix, cin_pix, x32, c, H, FP - 1)
                cin_pix = c
        if taps is not None:
            taps["inter32"] = x32
        for i in range(self.n_levels):                                        # Decoder, rmvpe.py:389-392, 371-377
            p = f"unet.decoder.layers.{i}."
            _, Hi, Wi, _, Pi, _, _ = self._geom(self.n_levels - i, Tp)        # the level the input lives on (pixel rows here)
            c, H, Wd, pk, P, FP, rows = self._geom(self.n_levels - 1 - i, Tp)
            cf = pk * c
            rows_in = Hi * Pi
            g16 = torch.empty(rows_in, 4 * c, dtype=torch.float16, device=dev)
            self._conv(x16, rows_in, cin_pix, cin_pix, p + "up", 4 * c, Pi - 1, taps=4, relu=True, y16=g16.data_ptr(), ldy16=4 * c,
                       mask=False)
            cat = cats[self.n_levels - 1 - i]
            self._check(lib.rvcb200_op_rmvpe_shuffle(C.c_void_p(g16.data_ptr()), C.c_void_p(cat.data_ptr()), Hi, Wi, c, 2 * cf, Pi, FP, pk,
                                                     self._stream), "rvcb200_op_rmvpe_shuffle")
            x16, x32, own = cat.data_ptr(), None, cat
            cin_f = 2 * cf
            for j in range(self.n_blocks):
                x32, x16, ldx, own = self._block(p + f"conv2.{j}.", x16, cin_f, cin_f, x32, cf, H, FP - 1)
                cin_f = cf
            cin_pix = c
        if taps is not None:
            taps["unet32"] = x32
        c, H, Wd, pk, P, FP, rows = self._geom(0, Tp)
        cnn32 = torch.empty(rows, pk * 16, dtype=torch.float32, device=dev)  # cnn, rmvpe.py:450, 468: [H][P][16], channels 0..2
        self._conv(x16, rows, pk * c, pk * c, "cnn", pk * 16, FP - 1, taps=9, relu=False, y32=cnn32.data_ptr(), ldy32=pk * 16)
        gx16 = torch.empty(Tp, 3 * N_MELS, dtype=torch.float16, device=dev)
        self._check(lib.rvcb200_op_rmvpe_gru_pack(C.c_void_p(cnn32.data_ptr()), 16, C.c_void_p(gx16.data_ptr()), Tp, N_MELS, P,
                                                  self._stream), "rvcb200_op_rmvpe_gru_pack")
        if taps is not None:
            taps["gru_in16"] = gx16
        # BiGRU: gi = W_ih x + b_ih for both directions on the tensor cores, then the recurrence
        G3 = 3 * self.gru_h
        gi = torch.empty(Tp, 2 * G3, dtype=torch.float32, device=dev)
        self._gemm(gx16.data_ptr(), Tp, 3 * N_MELS, "gru.ih", 2 * G3, gi.data_ptr())
        h16 = torch.empty(Tp, 2 * self.gru_h, dtype=torch.float16, device=dev)
        h32 = torch.empty(Tp, 2 * self.gru_h, dtype=torch.float32, device=dev) if taps is not None else None
        self._check(lib.rvcb200_op_rmvpe_gru(C.c_void_p(gi.data_ptr()), C.c_void_p(Wt["gru.hh.w"].data_ptr()),
                                             C.c_void_p(Wt["gru.hh.b"].data_ptr()), C.c_void_p(h16.data_ptr()),
                                             C.c_void_p(h32.data_ptr()) if h32 is not None else None, Tp, self._stream),
                    "rvcb200_op_rmvpe_gru")
        if taps is not None:
            taps["gru_out32"] = h32
        logits = torch.empty(Tp, N_CLASS_PAD, dtype=torch.float32, device=dev)
        self._gemm(h16.data_ptr(), Tp, 2 * self.gru_h, "fc", N_CLASS_PAD, logits.data_ptr())
        del own
        return logits

    def _gemm(self, x16: int, rows: int, cin: int, name: str, cout: int, y32: int):
        n_tile = 64
        mt = (rows + 127) // 128
        for cand in (256, 128):
            if cout % cand == 0 and mt * (cout // cand) >= 132:
                n_tile = cand
                break
        d = _lib.TcConvDesc()
        d.x16, d.L_in, d.padf = x16, rows, PADF
        d.w16, d.bias = self._wimg(name + ".w", n_tile).data_ptr(), self._w[name + ".b"].data_ptr()
        d.Cin, d.ntaps, d.dil, d.G = cin, 1, 1, 1
        d.N, d.Cout_total = n_tile, cout
        d.Lj, d.out_stride, d.Lp_out = rows, 1, ((rows + 127) // 128) * 128 + 128
        d.div, d.out_slope, d.alpha, d.pre_slope = 1.0, 1.0, 1.0, 1.0
        d.generic, d.f32_cl = 1, 1
        d.y32, d.ldy32 = y32, cout
        st = _lib.load().rvcb200_op_conv_tc(C.byref(d), 1, self._stream)
        if st != 0:
            raise RuntimeError(f"rvcb200_op_conv_tc failed with status {st} ({name}: Cin={cin}, Cout={cout}, rows={rows})")
        self.last_launches += 1

    @staticmethod
    def _padded_frames(n_frames: int) -> int:
        """rmvpe.py:593-595: the frame axis is reflect-padded to a multiple of 32."""
        pad = min(32 * ((n_frames - 1) // 32 + 1) - n_frames, n_frames)
        if pad >= n_frames and pad > 0:
            raise RuntimeError("Padding size should be less than the corresponding input dimension")     # what F.pad(reflect) raises
        return n_frames + pad

    def _begin(self):
        self._materialize()
        self.last_launches = 0
        self._stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _logmel(self, audio: torch.Tensor, want_mel: bool, want_img: bool):
        """audio fp32 [n] on the device -> (mel fp32 [128][n_frames] | None, img fp16 [Tp][P0][img_c] | None, n_frames, Tp)."""
        lib, Wt, dev = _lib.load(), self._w, self.device
        n = int(audio.shape[0])
        if n <= N_FFT // 2:
            raise RuntimeError(f"audio of {n} samples is shorter than the STFT's reflect padding ({N_FFT // 2})")
        n_frames = n // HOP + 1
        Tp = self._padded_frames(n_frames) if want_img else n_frames
        mel = torch.empty(N_MELS, n_frames, dtype=torch.float32, device=dev) if want_mel else None
        P0 = N_MELS + self._pack(0)
        img = torch.zeros(Tp, P0, self._img_c, dtype=torch.float16, device=dev) if want_img else None
        self._check(lib.rvcb200_op_rmvpe_logmel(C.c_void_p(audio.data_ptr()), n, C.c_void_p(Wt["mel.window"].data_ptr()),
                                                C.c_void_p(Wt["mel.twiddle"].data_ptr()), C.c_void_p(Wt["mel.basis"].data_ptr()),
                                                C.c_void_p(Wt["mel.range"].data_ptr()), self._bn_scale, self._bn_shift,
                                                C.c_void_p(mel.data_ptr()) if want_mel else None,
                                                C.c_void_p(img.data_ptr()) if want_img else None, n_frames, Tp, P0, self._img_c,
                                                self._stream),
                    "rvcb200_op_rmvpe_logmel")
        return mel, img, n_frames, Tp

    # ---- public API (rmvpe.py:559-684) -------------------------------------------------------------------------------------
    @torch.no_grad()
    def mel_extractor(self, audio: torch.Tensor, keyshift=0, speed=1, center=True) -> torch.Tensor:
        """`MelSpectrogram.forward` (rmvpe.py:489-556) for the only setting the class uses: audio [1, n] -> log-mel [1, 128, n // 160 + 1]."""
        if keyshift != 0 or speed != 1 or not center:
            raise NotImplementedError("keyshift / speed / center=False are never used by RMVPE (rmvpe.py:619)")
        if audio.dim() != 2 or audio.shape[0] != 1:
            raise ValueError("one utterance per call ([1, n]), like the reference")
        self._materialize()
        with torch.cuda.device(self.device):
            self._begin()
            mel, _, _, _ = self._logmel(audio[0].to(self.device, torch.float32).contiguous(), True, False)
        return mel[None]

    @torch.no_grad()
    def mel2hidden(self, mel: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        """rmvpe.py:591-608: log-mel [1, 128, n_frames] -> salience [1, n_frames, 360] (fp32, on the device)."""
        if mel.dim() != 3 or mel.shape[0] != 1 or mel.shape[1] != N_MELS:
            raise ValueError("mel must be [1, 128, n_frames]")
        self._materialize()
        with torch.cuda.device(self.device):
            self._begin()
            lib, dev = _lib.load(), self.device
            m = mel[0].to(dev, torch.float32).contiguous()
            n_frames = int(m.shape[1])
            Tp = self._padded_frames(n_frames)
            P0 = N_MELS + self._pack(0)
            img = torch.zeros(Tp, P0, self._img_c, dtype=torch.float16, device=dev)
            self._check(lib.rvcb200_op_rmvpe_mel_to_img(C.c_void_p(m.data_ptr()), C.c_void_p(img.data_ptr()), n_frames, Tp,
                                                        self._bn_scale, self._bn_shift, P0, self._img_c, self._stream),
                        "rvcb200_op_rmvpe_mel_to_img")
            logits = self._hidden_from_img(img, Tp, taps)
            hidden = torch.empty(n_frames, N_CLASS, dtype=torch.float32, device=dev)
            f0 = torch.empty(n_frames, dtype=torch.float64, device=dev)
            self._check(lib.rvcb200_op_rmvpe_decode(C.c_void_p(logits.data_ptr()), N_CLASS_PAD, 0, C.c_void_p(hidden.data_ptr()),
                                                    C.c_void_p(f0.data_ptr()), None, n_frames, 0.03, self._stream),
                        "rvcb200_op_rmvpe_decode")
        return hidden[None]

    def _decode_dev(self, salience, thred: float, want: str) -> np.ndarray:
        self._materialize()
        with torch.cuda.device(self.device):
            self._begin()
            s = torch.as_tensor(np.ascontiguousarray(salience, dtype=np.float32)).to(self.device)
            if s.dim() != 2 or s.shape[1] != N_CLASS:
                raise ValueError("salience must be [frames, 360]")
            T = int(s.shape[0])
            out = torch.empty(T, dtype=torch.float64, device=self.device)
            f0p, cp = (C.c_void_p(out.data_ptr()), None) if want == "f0" else (None, C.c_void_p(out.data_ptr()))
            self._check(_lib.load().rvcb200_op_rmvpe_decode(C.c_void_p(s.data_ptr()), N_CLASS, 1, None, f0p, cp, T, float(thred),
                                                            self._stream), "rvcb200_op_rmvpe_decode")
            return out.cpu().numpy()

    def decode(self, hidden, thred: float = 0.03) -> np.ndarray:
        """rmvpe.py:610-615: salience [frames, 360] (numpy, as the reference passes it) -> f0 [frames] float64."""
        return self._decode_dev(hidden, thred, "f0")

    def to_local_average_cents(self, salience, thred: float = 0.05) -> np.ndarray:
        """rmvpe.py:658-684."""
        return self._decode_dev(salience, thred, "cents")

    @torch.no_grad()
    def infer_from_audio(self, audio, thred: float = 0.03, taps: Optional[dict] = None) -> np.ndarray:
        """rmvpe.py:617-624: 16 kHz audio (numpy or tensor, [n]) -> f0 [n // 160 + 1] float64 (Hz, 0 where the salience <= thred).
        One host -> device copy of the audio, one device -> host copy of the f0 vector."""
        self._materialize()
        with torch.cuda.device(self.device):
            self._begin()
            lib, dev = _lib.load(), self.device
            a = torch.as_tensor(np.ascontiguousarray(audio, dtype=np.float32) if isinstance(audio, np.ndarray) else audio)
            a = a.reshape(-1).to(dev, torch.float32).contiguous()
            mel, img, n_frames, Tp = self._logmel(a, taps is not None, True)
            logits = self._hidden_from_img(img, Tp, taps)
            f0 = torch.empty(n_frames, dtype=torch.float64, device=dev)
            hidden = torch.empty(n_frames, N_CLASS, dtype=torch.float32, device=dev) if taps is not None else None
            self._check(lib.rvcb200_op_rmvpe_decode(C.c_void_p(logits.data_ptr()), N_CLASS_PAD, 0,
                                                    C.c_void_p(hidden.data_ptr()) if hidden is not None else None,
                                                    C.c_void_p(f0.data_ptr()), None, n_frames, float(thred), self._stream),
                        "rvcb200_op_rmvpe_decode")
            if taps is not None:
                taps["mel"], taps["hidden"], taps["logits"] = mel, hidden, logits
            return f0.cpu().numpy()

    def infer_from_audio_with_pitch(self, audio, thred: float = 0.03, f0_min=50, f0_max=1100) -> np.ndarray:
        """rmvpe.py:646-656."""
        return np.clip(self.infer_from_audio(audio, thred), a_min=f0_min, a_max=f0_max)
