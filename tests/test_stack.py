"""The cost-volume stack runner: input grouping / stage order on CPU, and on the GPU that the CUDA-graph replay
(single fused graph and per-input-group graphs fed by the HostFeeder) reproduces the eager kernel-by-kernel run."""
import pytest
import torch

from rpeflow_b200.stack import (CONFIGS, INPUT_GROUPS, CostVolumeStack, GraphedStack, HostFeeder, group_tensors,
                                make_host_inputs, tensors_nbytes, to_device)


def test_input_groups_cover_every_tensor_once():
    host = make_host_inputs(CONFIGS["tiny"], 2)
    groups = group_tensors(host)
    assert list(groups) == INPUT_GROUPS
    ids = [id(t) for g in groups.values() for t in g]
    assert len(ids) == len(set(ids))
    assert sum(t.numel() * t.element_size() for g in groups.values() for t in g) == tensors_nbytes(host)
    # dependency order of the copies: point data first, the largest activations (level 1) last
    assert INPUT_GROUPS[0] == "points" and INPUT_GROUPS[-1] == "lvl1"
    assert sum(t.numel() for t in groups["lvl1"]) > sum(t.numel() for t in groups["lvl2"])


def test_segments_follow_the_copy_order():
    stack = CostVolumeStack.__new__(CostVolumeStack)         # no device needed to list the segments
    assert [g for g, _ in CostVolumeStack.segments(stack)] == INPUT_GROUPS


def _same(a, b):
    if isinstance(a, torch.Tensor):
        return torch.equal(a, b)
    if isinstance(a, dict):
        return a.keys() == b.keys() and all(_same(a[k], b[k]) for k in a)
    return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))


@pytest.mark.gpu
def test_graph_replay_and_feeder_match_eager_run():
    dev = torch.device("cuda", 0)
    cfg = CONFIGS["tiny"]
    host = make_host_inputs(cfg, 2, pin=True)
    stack = CostVolumeStack(cfg, dev)
    x = to_device(host, dev)
    stack.concurrent = False
    want, _ = stack.run(x)                                   # kernel by kernel, one stream
    want = {k: v for k, v in want.items() if k != "event_voxel"}      # atomics: order-dependent last ulp
    stack.concurrent = True
    ints_w, flts_w = stack.checksum(stack.run(x)[0])

    fused = GraphedStack(stack, x, fused=True)
    for _ in range(2):
        out = fused.replay()
    torch.cuda.synchronize()
    assert _same({k: v for k, v in out.items() if k != "event_voxel"}, want)
    assert torch.equal(fused.checksum[0], ints_w)
    torch.testing.assert_close(fused.checksum[1], flts_w, rtol=1e-6, atol=1e-3)

    feeder = HostFeeder(host, dev, depth=2)                  # per-group graphs, inputs streamed from pinned memory
    runners = [GraphedStack(stack, slot, fused=False) for slot in feeder.slots]
    main = torch.cuda.current_stream()
    for step in range(3):
        slot = step % 2
        evs = feeder.issue(slot)
        out = runners[slot].replay(wait=lambda g, evs=evs: main.wait_event(evs[g]))
        main.synchronize()
        assert _same({k: v for k, v in out.items() if k != "event_voxel"}, want)
        assert torch.equal(runners[slot].checksum[0], ints_w)
