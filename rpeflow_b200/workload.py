"""Workload definitions for the cost-volume stack: shapes, synthetic inputs and the algorithmic-work census.

Pure Python + torch CPU tensors: this module never touches libb200flow.so (bench.py's reference arm loads it by
path so that the CPU baseline process maps no product library).  Shapes follow SURVEY §3.1 / §8d (file:line in
danqu130/RPEFlow are given where a number comes from the reference).
"""
from dataclasses import dataclass

import torch

LEVEL_CHANNELS = [32, 64, 96, 128, 192]          # levels 1..5 (pwc2d_core.py:28-40 / pwc3d_core.py:44-57)
PYRAMID_POINTS = [4096, 2048, 1024, 512, 256]    # RPEFlow.py:74 (hard-coded)
# bench.py's default frame pairs per GPU per step: 148 pairs = 296 point clouds = two FPS chains per SM of a B200 (the
# two-clouds-per-SM FPS kernel); round 1 and most of round 2 measured 74 (one cloud per SM)
BENCH_BATCH = 148
DEFAULT_BATCH = {"things": BENCH_BATCH, "dsec": BENCH_BATCH, "hd": 18, "hd_scaled": 18, "tiny": 4}   # hd: 36 clouds x 4-CTA clusters


@dataclass
class StackConfig:
    name: str = "things"          # "things" (cfg1, integer-pixel voxels) | "dsec" (cfg4, tri-linear voxels)
    height: int = 540
    width: int = 960
    n_points: int = 8192
    n_events: int = 1_000_000
    event_bins: int = 10
    k: int = 16
    max_displacement: int = 4
    focal: float = 1050.0
    max_depth: float = 35.0
    precision: int = 2            # Correlation3D arithmetic (include/b200flow.h): 2 = 3xTF32 on tcgen05, meets the fp32 bar
    pyramid: tuple = (4096, 2048, 1024, 512, 256)   # point pyramid (RPEFlow.py:74 hard-codes this list)

    @property
    def padded(self):             # resize_to_64x (models/utils.py:227-241)
        return (self.height + 63) // 64 * 64, (self.width + 63) // 64 * 64

    def level_hw(self, level):    # level 1..5 -> feature map size
        h, w = self.padded
        return h >> (level + 1), w >> (level + 1)

    @property
    def sensor(self):             # IDS parallel sensor (conf/test/things.yaml:18-20: divisor 32)
        h, w = self.padded
        return h // 32, w // 32


CONFIGS = {
    "things": StackConfig(),
    "dsec": StackConfig(name="dsec", height=480, width=640, n_events=1_500_000, focal=1050.0 * 640 / 960),
    "hd": StackConfig(name="hd", height=1080, width=1920, n_points=32768, n_events=4_000_000, focal=2100.0),
    # cfg5 with the point pyramid scaled with the cloud ("report both"; 8.28 G KNN pairs per sample)
    "hd_scaled": StackConfig(name="hd_scaled", height=1080, width=1920, n_points=32768, n_events=4_000_000, focal=2100.0,
                             pyramid=(16384, 8192, 4096, 2048, 1024)),
    "tiny": StackConfig(name="tiny", height=128, width=192, n_points=8192, n_events=20_000),
}


def make_host_inputs(cfg, batch, first_sample=0, seed_base=1000, pin=False):
    """Synthetic inputs for `batch` frame pairs (sample i is seeded seed_base + first_sample + i, so a shard can be
    regenerated anywhere).  Everything lives in host memory (pinned on request): point clouds already in the
    model's parallel-projection coordinates (models/utils.py:320-346), raw events, and the activations the hot
    ops consume."""
    hs, ws = cfg.sensor
    hp, wp = cfg.padded
    out = {"pcs": torch.empty(batch, 6, cfg.n_points), "feat2d": {}, "efeat2d": {}, "feat3d": {}, "flowfeat": {}}
    if cfg.name == "dsec":
        out["ev_x"] = torch.empty(batch, cfg.n_events)
        out["ev_y"] = torch.empty(batch, cfg.n_events)
        out["ev_p"] = torch.empty(batch, cfg.n_events)
        out["ev_t"] = torch.empty(batch, cfg.n_events, dtype=torch.int64)
    else:
        out["events"] = torch.empty(batch, cfg.n_events, 4)
    for lvl, c in zip(range(1, 6), LEVEL_CHANNELS):
        h, w = cfg.level_hw(lvl)
        n = cfg.pyramid[lvl - 1]
        out["feat2d"][lvl] = (torch.empty(batch, c, h, w), torch.empty(batch, c, h, w))   # image 1 / image 2 features
        out["efeat2d"][lvl] = torch.empty(batch, c, h, w)             # event features
        out["feat3d"][lvl] = (torch.empty(batch, c, n), torch.empty(batch, c, n))
        out["flowfeat"][lvl] = (torch.empty(batch, 96, h, w), torch.empty(batch, 64, n))   # decoder features (pwc2d_core.py:119)
    scale_w, scale_h = (ws - 1) / (wp - 1), (hs - 1) / (hp - 1)
    for i in range(batch):
        g = torch.Generator().manual_seed(seed_base + first_sample + i)
        for half in range(2):     # FT3D-shaped cloud -> perspective projection -> IDS parallel coordinates
            u = torch.rand(cfg.n_points, generator=g) * (cfg.width - 1)
            v = torch.rand(cfg.n_points, generator=g) * (cfg.height - 1)
            z = torch.rand(cfg.n_points, generator=g) * (cfg.max_depth - 2.0) + 2.0
            out["pcs"][i, 3 * half + 0] = u * (wp - 1) / (cfg.width - 1) * scale_w - (ws - 1) / 2
            out["pcs"][i, 3 * half + 1] = v * (hp - 1) / (cfg.height - 1) * scale_h - (hs - 1) / 2
            out["pcs"][i, 3 * half + 2] = (cfg.focal * torch.log(z) + 1.0) * min(scale_w, scale_h)
        n = cfg.n_events
        if cfg.name == "dsec":
            out["ev_x"][i] = torch.rand(n, generator=g) * (cfg.width - 1)
            out["ev_y"][i] = torch.rand(n, generator=g) * (cfg.height - 1)
            out["ev_t"][i] = torch.sort(torch.randint(0, 100_000, (n,), generator=g)).values
            out["ev_p"][i] = torch.randint(0, 2, (n,), generator=g).float()
        else:
            ev = out["events"][i]
            ev[:, 0] = torch.randint(0, cfg.width, (n,), generator=g).float()
            ev[:, 1] = torch.randint(0, cfg.height, (n,), generator=g).float()
            ev[:, 2] = torch.sort(torch.rand(n, generator=g)).values
            ev[:, 3] = torch.randint(0, 2, (n,), generator=g).float() * 2 - 1
        for lvl in range(1, 6):
            for half in range(2):
                out["feat2d"][lvl][half][i].normal_(generator=g)
                out["feat3d"][lvl][half][i].normal_(generator=g)
            out["efeat2d"][lvl][i].normal_(generator=g)
            out["flowfeat"][lvl][0][i].normal_(generator=g)
            out["flowfeat"][lvl][1][i].normal_(generator=g)
    if pin:
        out = _map_tensors(out, lambda t: t.pin_memory())
    return out


def _map_tensors(obj, fn):
    if isinstance(obj, torch.Tensor):
        return fn(obj)
    if isinstance(obj, dict):
        return {k: _map_tensors(v, fn) for k, v in obj.items()}
    if isinstance(obj, (tuple, list)):
        return type(obj)(_map_tensors(v, fn) for v in obj)
    return obj


def tensors_nbytes(obj):
    total = 0

    def add(t):
        nonlocal total
        total += t.numel() * t.element_size()
        return t
    _map_tensors(obj, add)
    return total


def to_device(host_inputs, device, non_blocking=True):
    return _map_tensors(host_inputs, lambda t: t.to(device, non_blocking=non_blocking))



# ---- algorithmic work of one frame pair, used by bench.py for the roofline arithmetic ----------------
def grid_sample_bytes(c, n, h, w):
    """grid_sample_wrapper, one call: 8N + 4CN + min(16CN, 4CHW)."""
    return 8 * n + 4 * c * n + min(16 * c * n, 4 * c * h * w)


def project_bytes(c2, c3, n, h, w):
    """project_feat_with_nn_corr, one call: HW(8 + 4C2 + 4(C3+3)) + 4N(C2+C3+2)."""
    return h * w * (8 + 4 * c2 + 4 * (c3 + 3)) + 4 * n * (c2 + c3 + 2)


def census_work(cfg):
    """Algorithmic bytes / pairs / flops of ONE frame pair, per op and per pyramid level."""
    hs = {}
    n_lvls = [cfg.n_points] + list(cfg.pyramid)
    knn = {"pyramid_k16": 0, "2d_k1": 0, "self_k16": 0, "cross_k16": 0, "interp_k3": 0}
    for lvl in range(5):
        knn["pyramid_k16"] += 2 * n_lvls[lvl] * n_lvls[lvl + 1]
    corr2d_bytes, sample_bytes, proj_bytes, corr3d_flops, corr3d_mma = {}, {}, {}, 0, 0
    gather_xyz = 0
    for lvl, c in zip(range(1, 6), LEVEL_CHANNELS):
        h, w = cfg.level_hw(lvl)
        n = cfg.pyramid[lvl - 1]
        knn["2d_k1"] += 2 * n * h * w
        knn["self_k16"] += n * n
        knn["cross_k16"] += n * n
        if lvl < 5:
            knn["interp_k3"] += cfg.pyramid[lvl] * n + n * n
        corr2d_bytes[lvl] = 4 * h * w * (2 * c + 81)
        sample_bytes[lvl] = sum(grid_sample_bytes(cc, n, h, w) for cc in (c, c, 83, c, 96))
        proj_bytes[lvl] = sum(project_bytes(c2, c3, n, h, w) for c2, c3 in ((c, c), (c, c), (81, c + 2), (96, 64)))
        corr3d_flops += 2 * n * cfg.k * ((2 * c + 3) * c + c * c) + 4 * n * cfg.k * (24 + 64 + 8 * c) + 4 * n * cfg.k * c
        corr3d_mma += 2 * n * cfg.k * c * c                 # the only contraction issued as tcgen05.mma (x3 for 3xTF32)
        gather_xyz += 2 * (8 * n + 2 * 4 * 3 * n)           # batch_indexing_channel_first on xyz: idx + read + write
    for i in range(5):
        knn["interp_k3"] += n_lvls[i + 1] * n_lvls[i]
    hs["knn_pairs_by_group"] = knn
    hs["knn_pairs"] = sum(knn.values())
    hs["fps_updates"] = 2 * max(cfg.pyramid) * cfg.n_points
    hs["corr2d_bytes"] = corr2d_bytes
    hs["grid_sample_bytes"] = sample_bytes
    hs["project_bytes"] = proj_bytes
    hs["gather_bytes"] = sum(sample_bytes.values()) + sum(proj_bytes.values())
    hs["gather_xyz_bytes"] = gather_xyz
    hs["corr3d_flops"] = corr3d_flops
    hs["corr3d_mma_flops"] = corr3d_mma
    hs["event_voxel_bytes"] = 16 * cfg.n_events + 4 * 2 * cfg.event_bins * cfg.height * cfg.width
    return hs
