"""N>1 host logic on CPU: batch-sample sharding is deterministic and rank-disjoint, and the verification
exchange (all-gather of checksums + max-over-ranks time) works over gloo with world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rpeflow_b200.stack import CONFIGS, census_work, make_host_inputs, tensors_nbytes


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, B = CONFIGS["tiny"], 2
    shard = make_host_inputs(cfg, B, first_sample=rank * B)                 # what bench.py does per rank
    digest = torch.stack([shard["pcs"].double().sum(), shard["events"].double().sum(),
                          shard["feat2d"][1][0].double().sum()])
    got = [torch.empty_like(digest) for _ in range(world)]
    dist.all_gather(got, digest)                                            # the verification collective
    common = make_host_inputs(cfg, 1, first_sample=0)["pcs"].double().sum().reshape(1)
    same = [torch.empty_like(common) for _ in range(world)]
    dist.all_gather(same, common)
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                                # max-over-ranks timing
    if rank == 0:
        ret["digests"] = [g.tolist() for g in got]
        ret["common_equal"] = bool(torch.equal(same[0], same[1]))
        ret["tmax"] = float(t.item())
    dist.destroy_process_group()


def test_two_rank_sharding_over_gloo():
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        d = ret["digests"]
        assert d[0] != d[1]                          # ranks own different frame pairs
        assert ret["common_equal"]                   # the common verification sample is identical on every rank
        assert ret["tmax"] == 11.0
    whole = make_host_inputs(CONFIGS["tiny"], 4)     # shards tile the global batch exactly
    r1 = make_host_inputs(CONFIGS["tiny"], 2, first_sample=2)
    assert torch.equal(whole["pcs"][2:], r1["pcs"]) and torch.equal(whole["events"][2:], r1["events"])
    assert torch.equal(whole["feat3d"][3][1][2:], r1["feat3d"][3][1])


def test_strong_scaling_split_tiles_the_global_batch():
    """bench.py --scaling strong: rank r of W takes the contiguous slice [r*G//W, (r+1)*G//W) of the global batch."""
    for G in (74, 7, 8):
        for W in (1, 2, 4, 8):
            cuts = [(r * G // W, (r + 1) * G // W) for r in range(W)]
            assert cuts[0][0] == 0 and cuts[-1][1] == G
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(W - 1))
            assert max(b - a for a, b in cuts) - min(b - a for a, b in cuts) <= 1
    whole = make_host_inputs(CONFIGS["tiny"], 3)
    lo, hi = 1 * 3 // 2, 2 * 3 // 2                   # rank 1 of 2 at a global batch of 3
    part = make_host_inputs(CONFIGS["tiny"], hi - lo, first_sample=lo)
    assert torch.equal(whole["pcs"][lo:hi], part["pcs"])


def test_census_matches_survey():
    w = census_work(CONFIGS["things"])
    assert w["knn_pairs"] == 535_756_288            # 535.8 M pairs per sample at cfg1
    assert w["fps_updates"] == 67_108_864           # 67.1 M
    assert [w["corr2d_bytes"][l] for l in range(1, 6)] == [20_044_800, 7_223_040, 2_358_720, 727_920, 251_100]
    assert tensors_nbytes(make_host_inputs(CONFIGS["tiny"], 1)) > 0
