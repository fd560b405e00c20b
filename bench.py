#!/usr/bin/env python
"""Headline benchmark: seconds of output audio synthesised per second (RTF^-1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp32|fp16|bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference        # the reference algorithm's CPU path (oracle port), rank 0 only

A "step" is one `net_g.infer()` of the BASELINE.json workload `configs[1]`: 48k_v2
(SynthesizerTrnMs768NSFsid), one 60 s segment of synthetic 768-d features + f0 (T = 6000 frames,
2 880 000 output samples) with seeded random weights, per GPU.  Segments are independent
(SURVEY.md §8e), so N GPUs run N segments with no collective on the data path: weak scaling,
value = N * 60 s / max-over-ranks device time.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from comfy_rvc_b200 import synthetic  # noqa: E402
from comfy_rvc_b200.config import NAMED_CONFIGS  # noqa: E402


def algorithmic_macs(cfg, T: int) -> dict:
    """MACs per batch item by component — the formula of SURVEY.md §8(d) (banded rel-pos, 1 MAC = 2 FLOP)."""
    H, F, Ci = cfg.hidden_channels, cfg.filter_channels, cfg.inter_channels
    U0 = cfg.upsample_initial_channel
    R = sum(len(ds) * 2 * k for k, ds in zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes)) \
        if cfg.resblock == "1" else sum(len(ds) * k for k, ds in zip(cfg.resblock_kernel_sizes, cfg.resblock_dilation_sizes))
    pre = T * Ci * U0 * 7
    ups = noise = res = 0
    L, C = T, U0
    for i, (u, k) in enumerate(zip(cfg.upsample_rates, cfg.upsample_kernel_sizes)):
        Cn, Ln = C // 2, L * u
        ups += L * C * Cn * k
        noise += Ln * Cn * cfg.noise_conv_geometry(i)[0]
        res += Ln * Cn * Cn * R
        L, C = Ln, Cn
    post = L * C * 7
    half = Ci // 2
    flow = cfg.n_flows * T * (half * H + cfg.flow_wn_layers * H * 2 * H * cfg.flow_kernel + 2 * H * 2 * H + H * H + H * half)
    W = 2 * cfg.window_size + 1
    enc_lin = T * cfg.feat_dim * H + cfg.n_layers * (4 * T * H * H + 2 * T * H * F * cfg.kernel_size) + T * H * 2 * Ci
    attn = cfg.n_layers * (2 * T * T * H + 2 * 2 * T * W * (H // cfg.n_heads))
    return {"conv_pre": pre, "ups": ups, "noise": noise, "resblocks": res, "conv_post": post, "flow": flow,
            "enc_linear": enc_lin, "attention": attn}


def pair_is_fused(C: int, k: int) -> bool:
    """ResBlock1 pairs that run as ONE kernel with h in shared memory (csrc/rbpair_tc.cu: rbpair_tc_supported)."""
    return os.environ.get("RVCB200_FUSE_PAIRS", "1") != "0" and C in (32, 64) and k in (3, 7)


def resblock_bytes(cfg, T: int):
    """Algorithmic HBM bytes (read, write) of the decoder resblock convolutions of one item on the tensor path
    (one fp16 stream copy, h as one 16-bit tensor -- or not at all where the pair is fused --, planar fp16
    branch sum)."""
    rd = wr = 0
    L, C = T, cfg.upsample_initial_channel
    nk = len(cfg.resblock_kernel_sizes)
    for i, u in enumerate(cfg.upsample_rates):
        L, C = L * u, C // 2
        E = L * C
        last_stage = i == len(cfg.upsample_rates) - 1
        for j, ds in enumerate(cfg.resblock_dilation_sizes):
            for d_i in range(len(ds)):
                if cfg.resblock == "1" and pair_is_fused(C, cfg.resblock_kernel_sizes[j]):
                    rd += 2 * E                                   # fused pair: stream in (h and the residual stay on the SM)
                elif cfg.resblock == "1":
                    rd += 2 * E; wr += 2 * E                      # conv1: stream in, h out
                    rd += 4 * E                                   # conv2: h + residual from the stream
                else:
                    rd += 2 * E
                if d_i < len(ds) - 1:
                    wr += 2 * E                                   # stream out
                else:
                    rd += 2 * E if j > 0 else 0                   # branch sum in
                    wr += 2 * E                                   # branch sum out
                    if j == nk - 1 and not last_stage:
                        wr += 2 * E                               # operand of the next transposed conv
    return rd, wr


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured (MEASURED_PEAKS.json, sustained)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(cfg, T: int, reps: int, warmup: int, threads: int):
    """RTF^-1 of the reference algorithm on the host cores (oracle port, fp32 PyTorch CPU)."""
    from oracle import rvc_oracle
    torch.set_num_threads(threads)
    sd = synthetic.make_state_dict(cfg)
    w = rvc_oracle.fold_weight_norm(sd)
    phone, lens, pitch, pitchf, sid = synthetic.make_inputs(cfg, 1, T)
    times = []
    for i in range(warmup + reps):
        noise = synthetic.draw_noise(cfg, 1, T, seed=7 + i)
        t0 = time.perf_counter()
        rvc_oracle.infer(w, cfg, phone, lens, pitch, pitchf, sid, *noise)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    audio_s = T * cfg.upp / cfg.sr
    return audio_s / float(np.median(times)), audio_s, times


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: everything libraries print there (NCCL's version banner under
    NCCL_DEBUG=VERSION, warnings of child tools) is sent to stderr for the lifetime of the process."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["fp32", "fp16", "bf16"],
                    help="bf16 (default; BASELINE north_star's tensor path, SNR >= 45 dB gate), fp16 (the reference's own GPU "
                         "dtype, same speed, ~61 dB) or fp32 (CUDA-core path, +-1 LSB gate)")
    ap.add_argument("--config", default="48k_v2")
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--cpu-frames", type=int, default=300, help="frames of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    cfg = NAMED_CONFIGS[args.config]
    T = int(round(args.seconds * 100))
    audio_s = T * cfg.upp / cfg.sr
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"{args.config} SynthesizerTrnMs{cfg.feat_dim}NSFsid.infer, B=1, T={T} frames ({audio_s:.0f} s segment) per GPU"
    metric = "seconds of output audio synthesised per second (RTF^-1)"

    # ------------------------------------------------------------------ reference arm (CPU) -------
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        steps, warm = max(1, args.steps), max(0, args.warmup)
        # bounded sample: cpu_frames (default 300 = 3 s of audio, ~0.3 s of CPU work on a 16-core host) of the same workload
        # per step, EXACTLY `steps` timed steps after `warm` untimed ones; value = audio produced / total time of the K steps
        _, sample_s, times = cpu_oracle_rate(cfg, args.cpu_frames, steps, warm, threads)
        rate = steps * sample_s / float(np.sum(times))
        line = {
            "impl": "reference", "metric": metric, "value": rate, "unit": "audio-s/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": float(np.mean(times)) * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "timed_sample": f"{args.cpu_frames} frames ({sample_s:.1f} s of audio) per step"},
            "cpu_baseline": {"value": rate, "unit": "audio-s/s", "cores": threads, "kind": "port",
                             "sample": f"oracle/rvc_oracle.py (fp32 PyTorch CPU restatement of the reference infer), "
                                       f"{args.cpu_frames} frames = {sample_s:.1f} s audio per step"},
            "e2e": {"value": rate, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        emit(line)
        return

    # ------------------------------------------------------------------ product arm (CUDA) --------
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    import comfy_rvc_b200 as rvc
    from comfy_rvc_b200 import _lib
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    sd = synthetic.make_state_dict(cfg)
    cls = rvc.SynthesizerTrnMs256NSFsid if cfg.feat_dim == 256 else rvc.SynthesizerTrnMs768NSFsid
    net = cls(*cfg.to_positional(), is_half=args.precision != "fp32")
    del net.enc_q
    net.load_state_dict({k: v.half() for k, v in sd.items()}, strict=False)
    net.eval().to(dev).set_precision(args.precision)

    phone, lens, pitch, pitchf, sid = synthetic.make_inputs(cfg, 1, T, seed=1 + rank)
    host = [t.pin_memory() for t in (phone, lens, pitch, pitchf, sid)]
    devin = [t.to(dev) for t in host]
    torch.manual_seed(1234 + rank)
    lib = _lib.load()

    def step_resident():
        return net.infer(*devin)[0]          # draws its noise with torch on the device, like the reference

    for _ in range(max(3, args.warmup)):
        step_resident()
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        for _ in range(8):                   # keep the GPU under load until nvidia-smi has delivered its first samples
            step_resident()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = net.last_launches * args.steps
    # per-class kernel times: a second pass of the same K steps with a CUDA-event pair around every launch (the event
    # records sit between kernels and serialise their programmatic dependent launches, so they stay out of `value`)
    lib.rvcb200_profile_enable(net._ctx, 1)
    for _ in range(args.steps):
        step_resident()
    torch.cuda.synchronize()
    cls_ms = (C.c_double * 8)()
    cls_n = (C.c_int64 * 8)()
    lib.rvcb200_profile_collect(net._ctx, cls_ms, cls_n)
    lib.rvcb200_profile_enable(net._ctx, 0)

    # end-to-end through the public API with host buffers: every step copies its inputs from pinned host memory and its PCM
    # back.  The driver is pipelined the way a serving loop would be: inputs of step i+1 travel on a copy stream while step i
    # computes, the PCM of step i-1 leaves on another; the synthesizer itself runs on the current stream.
    out_host = torch.empty(1, 1, T * cfg.upp, dtype=torch.float32).pin_memory()
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream(dev)

    # two preallocated sets of device input buffers (no allocator traffic in the loop): set k is refilled on the copy
    # stream as soon as the step that last read it has finished
    dev_in = [[torch.empty(t.shape, dtype=t.dtype, device=dev) for t in host] for _ in range(2)]
    read_done = [None, None]

    def fetch(i):
        k = i % 2
        with torch.cuda.stream(s_in):
            if read_done[k] is not None:
                s_in.wait_event(read_done[k])
            for d, h in zip(dev_in[k], host):
                d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s_in)
        return dev_in[k], ev

    def run_e2e(n):
        nxt = fetch(0)
        for i in range(n):
            ins, ev = nxt
            cur.wait_event(ev)
            nxt = fetch(i + 1)                         # H2D of the next step's inputs overlaps this step's kernels
            o = net.infer(*ins)[0]
            done = torch.cuda.Event()
            done.record(cur)
            read_done[i % 2] = done
            s_out.wait_event(done)
            with torch.cuda.stream(s_out):
                out_host.copy_(o, non_blocking=True)   # D2H of this step's PCM overlaps the next step's kernels
            o.record_stream(s_out)
        cur.wait_stream(s_out)
        cur.wait_stream(s_in)
        read_done[0] = read_done[1] = None

    run_e2e(max(3, args.warmup) + 3)         # warm-up: lets the caching allocator reach its steady set of blocks
    barrier()
    f0_, f1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0_.record()
    run_e2e(args.steps)
    f1_.record()
    barrier()
    ms_e2e = f0_.elapsed_time(f1_)
    clocks = sampler.stop() if sampler else None

    if dist is not None:
        print(f"rank {rank}: {ms / args.steps:.3f} ms/step resident, {ms_e2e / args.steps:.3f} ms/step end to end", file=sys.stderr)
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = world * args.steps * audio_s / (ms / 1e3)
    e2e_value = world * args.steps * audio_s / (ms_e2e / 1e3)
    macs = algorithmic_macs(cfg, T)
    conv_flops = 2.0 * macs["resblocks"]                      # class 0: the dominant kernel (decoder resblock convs)
    conv_ms_per_launch = cls_ms[0] / max(cls_n[0], 1)
    conv_launches_per_step = cls_n[0] / args.steps
    peaks = load_peaks()
    achieved = conv_flops / conv_launches_per_step / (conv_ms_per_launch * 1e-3) / 1e12
    rb_rd, rb_wr = resblock_bytes(cfg, T)
    rb_bytes = (rb_rd + rb_wr) / max(conv_launches_per_step, 1)
    rb_traffic = (1.001 * rb_rd + 0.80 * rb_wr) / max(conv_launches_per_step, 1) if args.precision != "fp32" else None
    h2d = sum(t.numel() * t.element_size() for t in host)
    line = {
        "metric": metric, "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "fp16": "f16 operands, f32 accumulate", "bf16": "bf16 operands (resblocks) / f16 operands (ladder, encoder, flow), f32 accumulate"}[args.precision],
        "data": "synthetic (seeded random-init weights stored as fp16, N(0,1) features, contour f0)",
        "config": {"workload": workload, "precision": args.precision, "parallelism": f"segments x{world}, no collective",
                   "l2": "no flush needed: each step streams >1 GB of stage activations (>> 126 MB L2)",
                   "roofline_pass": "per-launch CUDA events in a second pass of the same K steps (event records between kernels "
                                    "would serialise the programmatic dependent launches of the timed pass)"},
        "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": T * cfg.upp * 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "decoder resblock convolutions (rbconv_tc_kernel + fused-pair rbpair_tc_kernel on tcgen05; conv_f32_kernel in fp32 mode)",
                     "achieved": achieved, "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": achieved / peaks["tflops"],
                     "traffic": rb_traffic, "traffic_note": "dram__bytes_read+write per launch, average over the class's launches of a step: "
                     "algorithmic bytes x the read/write ratios ncu measured on sampled launches (profiles/r1_ncu_rbconv_v10.md, "
                     "r1_ncu_rbpair_v2.md: reads 1.001 x algorithmic, writes 0.77-0.80 x -- the rest is still dirty in L2 at kernel end)",
                     "algorithmic_bytes_per_launch": rb_bytes, "achieved_hbm_gbs": rb_bytes / (conv_ms_per_launch * 1e-3) / 1e9,
                     "hbm_peak_gbs": peaks["hbm_gbs"], "hbm_frac": rb_bytes / (conv_ms_per_launch * 1e-3) / 1e9 / peaks["hbm_gbs"],
                     "peak_source": peaks["src"],
                     "launches_per_step": conv_launches_per_step, "avg_launch_ms": conv_ms_per_launch,
                     "algorithmic_gflop_per_step": conv_flops / 1e9,
                     "share_of_step": cls_ms[0] / max(sum(cls_ms), 1e-9)},
        "time_by_class_ms_per_step": {"dec_resblocks": cls_ms[0] / args.steps, "attention": cls_ms[1] / args.steps,
                                      "sine_source": cls_ms[2] / args.steps, "glue": cls_ms[3] / args.steps,
                                      "dec_pre_ups": cls_ms[4] / args.steps, "flow": cls_ms[5] / args.steps,
                                      "enc_linear": cls_ms[6] / args.steps},
        "algorithmic_gflop_per_step": {k: 2.0 * v / 1e9 for k, v in macs.items()},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, sample_s, _ = cpu_oracle_rate(cfg, args.cpu_frames, 3, 1, threads)
        line["cpu_baseline"] = {"value": rate, "unit": "audio-s/s", "cores": threads, "kind": "port",
                                "sample": f"oracle/rvc_oracle.py fp32 PyTorch CPU, {args.cpu_frames} frames = {sample_s:.1f} s audio, median of 3"}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
