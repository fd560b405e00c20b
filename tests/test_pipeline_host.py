"""CPU: host logic of the segmented driver (comfy_rvc_b200.pipeline) against the pipeline oracle and the reference
golden fixtures, including the sharded path over world_size-2 gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from comfy_rvc_b200 import pipeline as pl
from comfy_rvc_b200 import synthetic
from oracle import pipeline_oracle
from tests._host_vc import HostVC, OracleNet
from tests._util import load_pipeline_golden


@pytest.mark.parametrize("secs,tiers,seed", [(9.0, (1, 1, 2, 3), 3), (2.0, (1, 1, 2, 3), 4), (31.0, (1, 2, 7, 8), 5)])
def test_plan_matches_oracle(secs, tiers, seed):
    audio = synthetic.make_song(secs, seed=seed)
    vc = pl.VC(40000, pl.PipelineConfig(*tiers, is_half=False))
    c = pipeline_oracle.Constants(*tiers, tgt_sr=40000)
    hp, audio_pad, opt_ts, segs = vc.plan(audio.copy())
    from scipy import signal
    ref_hp = signal.filtfilt(pipeline_oracle._BH, pipeline_oracle._AH, audio)
    assert np.array_equal(hp, ref_hp)
    assert opt_ts == pipeline_oracle.split_points(ref_hp, c)
    want = pipeline_oracle.segments(audio.shape[0], opt_ts, c)
    got = [(s.start, None if s is segs[-1] else s.end) for s in segs]
    assert got == want
    assert segs[-1].end == audio_pad.shape[0]
    assert [s.index for s in segs] == list(range(len(segs)))


@pytest.mark.parametrize("n_threads", [0, 1, 3, 16])
def test_native_quiet_point_is_the_reference_sum(n_threads):
    """csrc/host_plan.cu: per-element sequential float64 window sums, first minimum of |sum| -- bit for bit the
    reference's `audio_sum += audio_pad[i : i - window]` loop (vc_infer_pipeline.py:127-135), for any thread count,
    with plateaus of equal minima (exact zeros), ragged ranges and ranges shorter than a thread's share."""
    from comfy_rvc_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(11)
    window = 160
    a = np.round(rng.standard_normal(400000) * 8) / 8.0                      # coarse values: many exactly equal sums
    a[150000:170000] = 0.0                                                   # a plateau of exact zeros
    pad = np.pad(a, (window // 2, window // 2), mode="reflect")
    full = np.zeros_like(a)
    for i in range(window):
        full += pad[i: i - window]
    for lo, hi in ((0, 400000), (100000, 260001), (399000, 400000), (5, 6), (123, 4500)):
        seg = np.abs(full[lo:hi])
        want = int(np.where(seg == seg.min())[0][0])
        assert lib.rvcb200_host_quiet_point(pad.ctypes.data, lo, hi, window, n_threads) == want, (lo, hi)
    assert lib.rvcb200_host_quiet_point(pad.ctypes.data, 10, 10, window, n_threads) == -1
    assert lib.rvcb200_host_quiet_point(None, 0, 10, window, n_threads) == -1


def test_constants_and_tiers():
    vc = pl.VC(48000, pl.PipelineConfig.for_device(is_half=True))
    assert (vc.x_pad, vc.x_query, vc.x_center, vc.x_max) == (3, 10, 60, 64)               # config.py:124-129
    assert (vc.sr, vc.window, vc.t_pad, vc.t_pad_tgt, vc.t_pad2, vc.t_query, vc.t_center, vc.t_max) == \
        (16000, 160, 48000, 144000, 96000, 160000, 960000, 1024000)
    assert pl.PipelineConfig.for_device(is_half=False).x_center == 38                    # config.py:130-135
    assert pl.PipelineConfig.for_device(is_half=True, gpu_mem_gb=4).x_max == 32          # config.py:137-141


@pytest.mark.parametrize("key", [0, 3, -12])
def test_get_f0_matches_oracle(key):
    vc = pl.VC(40000, pl.PipelineConfig(1, 1, 2, 3, is_half=False))
    vc.f0_method_dict["synthetic"] = synthetic.pipeline_f0
    x = np.zeros(16000 * 4)
    coarse, f0 = vc.get_f0(x, key, "synthetic", f0_min=50, f0_max=1100)
    want_c, want_f = pipeline_oracle.f0_post(synthetic.pipeline_f0(x=x), key)
    assert coarse.dtype == np.int16 and np.array_equal(coarse, want_c) and np.array_equal(f0, want_f)
    assert coarse.min() >= 1 and coarse.max() <= 255
    with pytest.raises(Exception, match="not found"):
        vc.get_f0(x, 0, "harvest")
    # a one-element list is unwrapped like pitch_extraction.py:262; a callable works as a method too
    assert np.array_equal(vc.get_f0(x, key, ["synthetic"])[0], want_c)
    assert np.array_equal(vc.get_f0(x, key, synthetic.pipeline_f0)[0], want_c)


def test_lpt_assignment():
    lengths = [66, 52, 86, 46, 70, 61, 58, 49, 80, 47]
    for world in (1, 2, 4, 8):
        a = pl.assign_segments(lengths, world)
        assert sorted(i for r in a for i in r) == list(range(len(lengths)))
        loads = [sum(lengths[i] for i in r) for r in a]
        assert max(loads) <= sum(lengths) / world + max(lengths)            # LPT guarantee
        assert abs(pl.makespan_bound(lengths, world) - sum(lengths) / max(loads)) < 1e-12
    assert pl.assign_segments([5, 5, 5], 2) == [[0, 2], [1]]                 # ties by index, deterministic
    assert pl.assign_segments([], 2) == [[], []]


def test_product_vc_has_no_cpu_path():
    vc = pl.VC(40000, pl.PipelineConfig(1, 1, 2, 3, is_half=False, device="cpu"))
    vc.f0_method_dict["synthetic"] = synthetic.pipeline_f0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vc.pipeline(synthetic.FakeHubert(256), object(), 0, synthetic.make_song(1.0), [0, 0, 0], 0, "synthetic", "median", "",
                    0.0, 1, 3, 40000, 0, 1.0, "v1", 0.5, 160, False, False)
    for kw in (dict(rms=0.5, rs=0), dict(rms=1.0, rs=22050)):
        with pytest.raises(NotImplementedError):
            vc.pipeline(None, None, 0, np.zeros(16000), [0, 0, 0], 0, "synthetic", "median", "", 0.0, 1, 3, 40000, kw["rs"],
                        kw["rms"], "v1", 0.5, 160, False, False)


def _run_pipeline(case, group_ready):
    cfg = case["cfg"]
    vc = HostVC(cfg.sr, pl.PipelineConfig(*case["tiers"], is_half=False, device="cpu"), noise="reference")
    vc.f0_method_dict["synthetic"] = synthetic.pipeline_f0
    net = OracleNet(cfg, case["sd"])
    torch.manual_seed(case["rseed"])
    out = vc.pipeline(case["hubert"], net, 0, case["audio"].copy(), [0, 0, 0], case["f0_up_key"], "synthetic", "median",
                      case["file_index"], case["index_rate"], 1, 3, cfg.sr, 0, 1.0, case["version"], case["protect"], 160,
                      False, False, None, 50, 1100)
    return out, vc.last_plan


def _worker(rank, world, port, name, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out, plan = _run_pipeline(load_pipeline_golden(name), True)
        q.put((rank, None if out is None else out.copy(), plan["assignment"]))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("name", ["p1_40k_v1_4seg", "p2_32k_v2_protect_index"])
def test_sharded_pipeline_world2_gloo_matches_single_process_and_reference(name):
    """Segments sharded over 2 ranks (gloo): rank 0's song is bit-identical to the unsharded run and within ±1 LSB of
    the reference golden; the other rank returns None; each segment ran on exactly one rank."""
    torch.set_num_threads(1)
    case = load_pipeline_golden(name)
    single, plan1 = _run_pipeline(case, False)
    gold = case["gold"]["out_i16"]
    assert single.dtype == np.int16 and single.shape == gold.shape
    assert np.abs(single.astype(np.int32) - gold.astype(np.int32)).max() <= 1
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        r, out, assignment = q.get(timeout=600)
        res[r] = (out, assignment)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1][0] is None
    a = res[0][1]
    assert sorted(i for r in a for i in r) == list(range(len(plan1["segments"]))) and all(len(r) > 0 for r in a)
    assert np.array_equal(res[0][0], single)


@pytest.mark.parametrize("secs,seed,t_pad", [(41.7, 3, 48000), (7.0, 4, 16000), (2.5, 5, 48000), (0.001, 6, 16000)])
def test_filtfilt_pad_is_bit_identical_to_scipy(secs, seed, t_pad):
    """csrc/host_plan.cu restates scipy's filtfilt (odd padding, transposed direct form II, zi scaling) + np.pad reflect:
    bit for bit the same doubles as `np.pad(signal.filtfilt(bh, ah, audio), (t_pad, t_pad), mode="reflect")`
    (vc_infer_pipeline.py:122, :141); clips too short for the C path fall back to scipy / numpy."""
    from scipy import signal
    if secs < 0.01:
        audio = np.random.default_rng(seed).standard_normal(19) * 0.1          # n = padlen + 1: scipy's shortest input
        t_pad = 10
    else:
        audio = synthetic.make_song(secs, seed=seed)
    want = np.pad(signal.filtfilt(pl._BH, pl._AH, audio), (t_pad, t_pad), mode="reflect")
    got = pl.filtfilt_pad(audio, t_pad)
    assert got.dtype == np.float64 and got.shape == want.shape
    assert np.array_equal(got, want)
    buf = np.full(want.shape[0] + 7, np.nan)
    assert np.array_equal(pl.filtfilt_pad(audio, t_pad, out=buf), want) and np.isnan(buf[want.shape[0]:]).all()


def test_plan_matches_reference_order_of_operations():
    """VC.plan (C filtfilt + fused padding + quiet points on a view of the padded buffer) against the reference's own
    sequence of scipy / numpy calls (vc_infer_pipeline.py:122-141)."""
    from scipy import signal
    vc = pl.VC(40000, pl.PipelineConfig(1, 6, 38, 41, is_half=False, device="cuda:0"))
    audio = synthetic.make_song(150.0, seed=9)
    a, ap, opt_ts, segs = vc.plan(audio)
    a_ref = signal.filtfilt(pl._BH, pl._AH, audio)
    assert np.array_equal(a, a_ref)
    assert np.array_equal(ap, np.pad(a_ref, (vc.t_pad, vc.t_pad), mode="reflect"))
    assert opt_ts == pl.split_points(a_ref, vc.window, vc.t_query, vc.t_center, vc.t_max) and len(opt_ts) >= 2
    assert [(s.start, s.end) for s in segs] == [(s.start, s.end) for s in pl.plan_segments(ap.shape[0], opt_ts, vc.window, vc.t_pad2)]
