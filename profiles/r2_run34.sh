#!/bin/bash
true
python - <<'PY'
import os, sys, torch
sys.path.insert(0, '.')
import rpeflow_b200 as b200
g = torch.Generator().manual_seed(1)
for B in (148, 296, 444):
    x = (torch.rand(B, 8192, 3, generator=g) * torch.tensor([30.0, 17.0, 90.0])).cuda()
    for knob in ("0", "3", "4"):
        os.environ["B200_FPS_T"] = knob
        for _ in range(2): b200.ops._furthest_point_sampling_cuda(x, 4096)
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); b200.ops._furthest_point_sampling_cuda(x, 4096); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
        print(f"clouds {B} variant {knob}: {min(ts):.3f} ms", flush=True)
PY
