import os, sys, torch
sys.path.insert(0, "/root/repo")
from rpeflow_b200 import ops
from rpeflow_b200.stack import CONFIGS, make_host_inputs
x = make_host_inputs(CONFIGS["things"], 74)
pcs = x["pcs"]
both = torch.cat([pcs[:, :3], pcs[:, 3:]], 0).transpose(1, 2).contiguous().cuda()
ref = None
for t in ("", "2", "1", "full"):
    os.environ.pop("B200_FPS_T", None); os.environ.pop("B200_FPS_FULL_SCAN", None)
    if t == "full": os.environ["B200_FPS_FULL_SCAN"] = "1"
    elif t: os.environ["B200_FPS_T"] = t
    for _ in range(2): out = ops.furthest_point_sampling(both, 4096)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = ops.furthest_point_sampling(both, 4096); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    ref = out if ref is None else ref
    print(t or "512", min(ts), torch.equal(out, ref))
