"""GPU: launch-bound sizes replay a captured CUDA graph (synthesizer.SynthesizerB200._infer_graphed).  A replay must be
bit-identical to the eager launch sequence, must see the inputs of THIS call, and must hand out tensors the caller owns."""
import pytest
import torch

from comfy_rvc_b200 import synthetic
from comfy_rvc_b200.config import NAMED_CONFIGS, nono
from tests.test_parity_gpu import build_net

pytestmark = pytest.mark.gpu


def _call(net, cfg, inputs, noise):
    ins = [t.cuda() for t in inputs]
    if cfg.f0:
        return net.infer(*ins, noise=noise)
    return net.infer(ins[0], ins[1], ins[4], noise=noise[:1])


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("cfg_name,f0", [("48k_v2", True), ("40k", True), ("32k_v2", False)])
def test_graph_replay_equals_eager(cfg_name, f0, precision):
    cfg = NAMED_CONFIGS[cfg_name] if f0 else nono(NAMED_CONFIGS[cfg_name])
    sd = synthetic.make_state_dict(cfg)
    eager = build_net(cfg, sd, precision)
    eager.graph_max_frames = 0
    graphed = build_net(cfg, sd, precision)
    assert graphed.graph_max_frames >= 200
    keep = []
    for call, (B, T, seed) in enumerate([(1, 100, 1), (2, 64, 2), (1, 100, 3), (2, 64, 4), (1, 100, 5)]):
        inputs = synthetic.make_inputs(cfg, B, T, seed=seed, lengths=[T] + [T - 9] * (B - 1))
        noise = synthetic.draw_noise(cfg, B, T, seed=10 + seed)
        want = _call(eager, cfg, inputs, noise)
        got = _call(graphed, cfg, inputs, noise)
        torch.cuda.synchronize()
        assert eager.last_graph_replay is False
        assert graphed.last_graph_replay is (call >= 2)          # first call of a key runs eagerly and captures
        assert graphed.last_launches == eager.last_launches > 0
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
        for a, b in zip(got[2], want[2]):
            assert torch.equal(a, b)
        keep.append((got[0], want[0].clone()))
    for got_o, want_o in keep:                                    # earlier results were not overwritten by later replays
        assert torch.equal(got_o, want_o)


def test_graph_draws_fresh_noise_and_survives_precision_switch():
    cfg = NAMED_CONFIGS["48k_v2"]
    sd = synthetic.make_state_dict(cfg)
    net = build_net(cfg, sd, "bf16")
    ins = [t.cuda() for t in synthetic.make_inputs(cfg, 1, 80)]
    torch.manual_seed(3)
    a = net.infer(*ins)[0]
    b = net.infer(*ins)[0]                                        # replay; noise drawn by torch outside the graph
    assert net.last_graph_replay and not torch.equal(a, b)
    torch.manual_seed(3)
    a2 = net.infer(*ins)[0]
    assert torch.equal(a, a2)                                     # same RNG stream as the eager path
    net.set_precision("fp16")                                     # new weight images: the captured graphs are dropped
    c = net.infer(*ins)[0]
    assert net.last_graph_replay is False
    net.set_precision("bf16")
    torch.manual_seed(3)
    assert torch.equal(net.infer(*ins)[0], a)
    net.graph_max_frames = 0
    torch.manual_seed(3)
    assert torch.equal(net.infer(*ins)[0], a) and net.last_graph_replay is False
    assert c.shape == a.shape


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_programmatic_dependent_launch_is_bit_identical(precision):
    """Launch-bound sizes run with programmatic dependent launch (engine.cu PdlScope: kernel n+1's prologue under kernel
    n's tail; every kernel waits on `griddepcontrol.wait` before it touches a predecessor's data).  A call that taps an
    intermediate runs without it (and without graph replay): same waveform bit for bit."""
    cfg = NAMED_CONFIGS["48k_v2"]
    sd = synthetic.make_state_dict(cfg)
    net = build_net(cfg, sd, precision)
    for T in (40, 150):
        inputs = [t.cuda() for t in synthetic.make_inputs(cfg, 2, T, seed=T, lengths=[T, T - 7])]
        noise = synthetic.draw_noise(cfg, 2, T, seed=T + 1)
        plain = net.infer(*inputs, noise=noise, taps={"x_enc": None})[0]
        for _ in range(3):                                        # eager + capture, then replays
            assert torch.equal(net.infer(*inputs, noise=noise)[0], plain)
