"""Event voxelisers at the bench sizes: a9 (things: 1 M events, 540x960) and a10 (dsec: 1.5 M events, 480x640, tri-linear)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from rpeflow_b200 import events as E
from rpeflow_b200.workload import CONFIGS, make_host_inputs
dev = torch.device("cuda", 0)
B = int(os.environ.get("BATCH", "8"))
def t(fn, reps=5):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(reps):
        a.record(); fn(); b.record(); b.synchronize(); ms.append(a.elapsed_time(b))
    ms.sort(); return ms[len(ms) // 2]
for name in ("things", "dsec"):
    cfg = CONFIGS[name]
    host = make_host_inputs(cfg, B)
    grids = torch.empty((B, 20, cfg.height, cfg.width), device=dev)
    if name == "dsec":
        x, y, tt, p = (host[k].to(dev) for k in ("ev_x", "ev_y", "ev_t", "ev_p"))
        fn = lambda: [E.events_to_voxel_trilinear_device(x[i], y[i], tt[i], p[i], 10, cfg.height, cfg.width, True, out=grids[i]) for i in range(B)]
    else:
        ev = host["events"].to(dev)
        fn = lambda: [E.events_to_voxel_device(ev[i], 10, cfg.height, cfg.width, True, check_range=False, out=grids[i]) for i in range(B)]
    ms = t(fn)
    print(f"{name}: {ms / B * 1e3:.1f} us per sample ({cfg.n_events} events)", flush=True)
