"""project_feat_with_nn_corr at level-1 / level-2 shapes, batch 74: device time per call (L2 flushed), GB/s of algorithmic bytes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import rpeflow_b200 as b200
from rpeflow_b200.workload import project_bytes

dev = torch.device("cuda", 0)
B = int(os.environ.get("BATCH", "74"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
cases = [(32, 32, 144, 240, 4096), (81, 34, 144, 240, 4096), (96, 64, 144, 240, 4096), (64, 64, 72, 120, 2048), (96, 64, 72, 120, 2048), (192, 192, 9, 15, 256)]
if os.environ.get("CASES"):
    cases = [cases[int(i)] for i in os.environ["CASES"].split(",")]
for C2, C3, H, W, N in cases:
    f2 = torch.randn(B, C2, H, W, device=dev); f3 = torch.randn(B, C3, N, device=dev)
    xy = torch.rand(B, 2, N, device=dev) * torch.tensor([W - 1.0, H - 1.0], device=dev).view(1, 2, 1)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev), torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
    grid = torch.stack([xs, ys], 0).reshape(1, 2, H * W).expand(B, 2, H * W).contiguous()
    nn = b200.k_nearest_neighbor(xy, grid, 1)[..., 0].contiguous()
    fn = lambda: b200.project_feat_with_nn_corr(xy, f2, f3, nn)
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(7):
        flush.zero_(); a.record(); fn(); b.record(); b.synchronize(); ms.append(a.elapsed_time(b))
    ms.sort()
    nb = project_bytes(C2, C3, N, H, W) * B
    print(f"C2={C2:3d} C3={C3:3d} {H}x{W} N={N}: {ms[3]*1e3:8.1f} us  {nb/ms[3]/1e6:7.1f} GB/s ({nb/ms[3]/1e6/6539.5:.2f} of peak)", flush=True)
