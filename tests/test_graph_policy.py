"""CPU: which infer() calls replay a CUDA graph (synthesizer.graph_policy; the replay itself is tests/test_graphs_gpu.py)."""
from collections import OrderedDict

from comfy_rvc_b200.synthesizer import graph_policy


def test_launch_bound_sizes_always_and_larger_ones_from_their_second_occurrence():
    seen, graphs = OrderedDict(), {}
    small, big = (1, 100, 2), (1, 6000, 2)
    assert graph_policy(small, 100, 2500, True, seen, graphs)            # launch-bound: from the first call
    assert not graph_policy(big, 6000, 2500, True, seen, graphs)         # first occurrence: stream launches
    assert graph_policy(big, 6000, 2500, True, seen, graphs)             # seen before: capture + replay
    assert graph_policy((4, 625, 2), 2500, 2500, True, seen, graphs)     # B * T counts, not T
    assert not graph_policy((1, 6000, 1), 6000, 2500, True, seen, graphs)   # precision is part of the key


def test_switches():
    seen, graphs = OrderedDict(), {}
    big = (1, 6000, 2)
    for _ in range(3):
        assert not graph_policy(big, 6000, 2500, False, seen, graphs)    # RVCB200_GRAPH_REPEAT=0
    graphs[big] = object()
    assert graph_policy(big, 6000, 2500, False, seen, graphs)            # a key that has a graph keeps using it
    seen2 = OrderedDict()
    for _ in range(3):
        assert not graph_policy((1, 50, 2), 50, 0, True, seen2, {})      # graph_max_frames = 0 switches everything off
    assert len(seen2) == 0


def test_seen_set_is_bounded_lru():
    seen = OrderedDict()
    for t in range(300):
        graph_policy((1, 3000 + t, 2), 3000 + t, 2500, True, seen, {})
    assert len(seen) == 256 and (1, 3000, 2) not in seen and (1, 3299, 2) in seen
    assert graph_policy((1, 3299, 2), 3299, 2500, True, seen, {})
    assert not graph_policy((1, 3000, 2), 3000, 2500, True, seen, {})    # evicted: counts as new again
