"""FPS time vs cluster size (B200_FPS_CLUSTER) for 64 clouds x 8192 points -> 4096 samples (bench batch 32)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CODE = r'''
import sys, torch; sys.path.insert(0, %r)
from rpeflow_b200 import ops
for clouds in (64, 16):
    x = torch.rand(clouds, 8192, 3, device="cuda")
    for _ in range(2): ops._furthest_point_sampling_cuda(x, 4096)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): ops._furthest_point_sampling_cuda(x, 4096)
    b.record(); torch.cuda.synchronize()
    print("clouds", clouds, "ms", a.elapsed_time(b) / 5, "ns/iter", a.elapsed_time(b) / 5 * 1e6 / 4095)
''' % ROOT
for cs in ("1", "2", "4"):
    env = dict(os.environ, B200_FPS_CLUSTER=cs)
    print("cluster", cs, flush=True)
    subprocess.run([sys.executable, "-c", CODE], env=env, check=True)
