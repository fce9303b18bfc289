import sys, torch
sys.path.insert(0, "/root/repo")
from rpeflow_b200 import ops
a = torch.randn(32, 32, 144, 240, device="cuda"); b = torch.randn_like(a)
for _ in range(3): ops.correlation2d(a, b, 4)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.correlation2d(a, b, 4); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1))
print(sorted(ts)[5])
