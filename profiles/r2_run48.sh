#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fps" 2>&1 | tail -2
python - <<'PY'
import os, sys, torch
sys.path.insert(0, '.')
import rpeflow_b200 as b200
g = torch.Generator().manual_seed(1)
x = (torch.rand(296, 8192, 3, generator=g) * torch.tensor([30.0, 17.0, 90.0])).cuda()
for _ in range(2): b200.ops._furthest_point_sampling_cuda(x, 4096)
ts = []
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); b200.ops._furthest_point_sampling_cuda(x, 4096); b.record(); b.synchronize(); ts.append(a.elapsed_time(b))
print(f"296 clouds, shared variant at 53 registers: {min(ts):.3f} ms")
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-model --no-e2e --no-cpu-baseline --no-ref-cuda > gpurun_out/r2_f56.json 2> gpurun_out/r2_f56.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_f56.json') if l.startswith('{')][-1])
print(round(d['value'],1), 'fp/s', round(d['ms_per_step'],3), 'ms/step', 'fps', d['ops']['per_op_ms_per_step']['fps'])
PY
