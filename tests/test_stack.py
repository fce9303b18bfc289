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


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["dsec", "hd"])
def test_other_baseline_configs_run_and_index_ops_stay_exact(name):
    """BASELINE.json configs[3] (DSEC-shaped: 640x480, tri-linear voxels) and configs[4] (1920x1080, 32768 points: FPS
    over a 4-CTA cluster, KNN over 32768 inputs): one frame pair through the whole stack; FPS and a sample of the KNN
    results are compared with the CPU oracle, float outputs must be finite and the voxel grid must hold every event."""
    import numpy as np
    from oracle import spec
    dev = torch.device("cuda", 0)
    cfg = CONFIGS[name]
    host = make_host_inputs(cfg, 1)
    stack = CostVolumeStack(cfg, dev)
    out, _ = stack.run(to_device(host, dev))
    torch.cuda.synchronize()
    pcs = host["pcs"]
    both = torch.cat([pcs[:, :3], pcs[:, 3:]], 0).transpose(1, 2).contiguous().numpy()
    np.testing.assert_array_equal(out["fps_idx"].cpu().numpy(), spec.fps(both, 4096))
    idx1 = out["fps_idx"][0].cpu()
    for lvl in (1, 4):
        n = [4096, 2048, 1024, 512, 256][lvl - 1]
        xyz = pcs[0, :3][:, idx1[:n]].t().contiguous().numpy()[None]
        want = spec.knn(xyz, xyz[:, :64], 16)
        np.testing.assert_array_equal(out["knn_self"][lvl][:, :64].cpu().numpy(), want)
    for lvl in range(1, 6):
        for t in [out["corr2d"][lvl], out["corr3d"][lvl]] + out["proj"][lvl] + out["sample"][lvl]:
            assert bool(torch.isfinite(t).all())
    n_ev = cfg.n_events
    total = float(out["event_voxel"].double().sum())
    assert abs(total - n_ev) <= 1e-4 * n_ev if name == "hd" else total > 0     # integer-pixel voxels conserve the event count
