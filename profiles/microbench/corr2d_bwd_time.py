"""corr2d backward (a2): b200_corr2d_bwd vs the reference's own CUDA kernels (oracle/_ref, when built) at the pyramid levels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rpeflow_b200 import ops
from oracle import refcuda
B = int(os.environ.get("BATCH", "8"))


def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)


for (C, H, W) in ((32, 144, 240), (64, 72, 120), (96, 36, 60), (128, 18, 30)):
    a = torch.randn(B, H, W, C, device="cuda"); b = torch.randn_like(a)
    go = torch.randn(B, 81, H, W, device="cuda")
    ours = t(lambda: ops._correlation_backward_cuda(go, a, b, 4))
    fwd = t(lambda: ops._correlation_forward_cuda(a, b, 4))
    line = f"C={C} {H}x{W} B={B}: bwd {ours:.3f} ms (fwd {fwd:.3f} ms)"
    if refcuda.available():
        ref = t(lambda: refcuda.corr2d_bwd(go, a, b, 4, sync=False))
        rf = t(lambda: refcuda.corr2d_fwd(a, b, 4, sync=False))
        g1, g2 = ops._correlation_backward_cuda(go, a, b, 4)
        r1, r2 = refcuda.corr2d_bwd(go, a, b, 4)
        err = max((g1 - r1).abs().max().item(), (g2 - r2).abs().max().item())
        line += f"; reference kernels bwd {ref:.3f} ms fwd {rf:.3f} ms; max|diff| {err:.2e}"
    print(line)
