"""The KNN call shapes of one forward at batch 74 (FT3D-shaped clouds from the bench generator): device time per call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import rpeflow_b200 as b200
from rpeflow_b200 import ops, projection
from rpeflow_b200.workload import CONFIGS, make_host_inputs

dev = torch.device("cuda", 0)
B = int(os.environ.get("BATCH", "74"))
cfg = CONFIGS["things"]
host = make_host_inputs(cfg, B)
pc1 = host["pcs"][:, :3].contiguous().to(dev)
idx = ops.furthest_point_sampling(pc1.transpose(1, 2).contiguous(), 4096)
lv = [pc1] + [projection.batch_indexing_channel_first(pc1, idx[:, :n]) for n in cfg.pyramid]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, reps=7):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(reps):
        flush.zero_(); a.record(); fn(); b.record(); b.synchronize(); ms.append(a.elapsed_time(b))
    ms.sort(); return ms[len(ms) // 2] * 1e3
tot = 0
for name, inp, qry, k in [("pyr 8192->4096 k16", lv[0], lv[1], 16), ("pyr 4096->2048 k16", lv[1], lv[2], 16), ("pyr 2048->1024 k16", lv[2], lv[3], 16),
                          ("self 4096 k16", lv[1], lv[1], 16), ("self 2048 k16", lv[2], lv[2], 16), ("self 1024 k16", lv[3], lv[3], 16),
                          ("interp 2048->4096 k3", lv[2], lv[1], 3), ("interp 4096->8192 k3", lv[1], lv[0], 3), ("self 4096 k32", lv[1], lv[1], 32)]:
    us = t(lambda: ops.k_nearest_neighbor(inp, qry, k))
    print(f"{name:24s} {us:8.1f} us", flush=True)
