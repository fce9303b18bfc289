"""Launches, once each, the HBM-bound ops of one `things` step at the bench batch so that one `ncu --set full` capture holds
their DRAM traffic:  the four project_feat_with_nn_corr calls of level 1, correlation2d level 1, the event voxeliser.
Run under ncu (profiles/r2_traffic.sh); profiles/r2_traffic_parse.py turns the raw csv into profiles/traffic.json."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rpeflow_b200 as b200
from rpeflow_b200 import events as E
from rpeflow_b200.workload import CONFIGS

dev = torch.device("cuda", 0)
cfg = CONFIGS["things"]
B = int(os.environ.get("BATCH", "74"))
H, W = cfg.level_hw(1)
N = cfg.pyramid[0]
g = torch.Generator(device="cpu").manual_seed(5)
xy = (torch.rand(B, 2, N, generator=g) * torch.tensor([W - 1.0, H - 1.0]).view(1, 2, 1)).to(dev)
ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev), torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
grid = torch.stack([xs, ys], 0).reshape(1, 2, H * W).expand(B, 2, H * W).contiguous()
nn = b200.k_nearest_neighbor(xy, grid, 1)[..., 0].contiguous()
torch.cuda.synchronize()
for C2, C3 in ((32, 32), (32, 32), (81, 34), (96, 64)):
    f2 = torch.randn(B, C2, H, W, device=dev); f3 = torch.randn(B, C3, N, device=dev)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("project_L1")
    b200.project_feat_with_nn_corr(xy, f2, f3, nn)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
    del f2, f3
a, c = torch.randn(B, 32, H, W, device=dev), torch.randn(B, 32, H, W, device=dev)
torch.cuda.synchronize()
b200.correlation2d(a, c, 4)
torch.cuda.synchronize()
n = cfg.n_events
ev = torch.empty(4, n, 4, device=dev)
ev[..., 0] = torch.randint(0, cfg.width, (4, n), device=dev).float()
ev[..., 1] = torch.randint(0, cfg.height, (4, n), device=dev).float()
ev[..., 2] = torch.sort(torch.rand(4, n, device=dev), dim=1).values
ev[..., 3] = torch.randint(0, 2, (4, n), device=dev).float() * 2 - 1
grids = torch.empty(4, 20, cfg.height, cfg.width, device=dev)
torch.cuda.synchronize()
for i in range(4):
    E.events_to_voxel_device(ev[i], 10, cfg.height, cfg.width, True, check_range=False, out=grids[i])
torch.cuda.synchronize()
print("done")
