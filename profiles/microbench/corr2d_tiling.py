"""corr2d tilings side by side: B200_CORR2D_TILING=32 (corr2d_nchw.cu) vs 48 (corr2d_diag.cu) at the five pyramid levels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rpeflow_b200 import ops
B = int(os.environ.get("BATCH", "74"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (C, H, W) in ((32, 144, 240), (64, 72, 120), (96, 36, 60), (128, 18, 30), (192, 9, 15)):
    if W % 4:
        continue
    a = torch.randn(B, C, H, W, device="cuda"); b = torch.randn_like(a)
    res = {}
    for til in ("32", "48"):
        os.environ["B200_CORR2D_TILING"] = til
        for _ in range(3): out = ops.correlation2d(a, b, 4)
        ts = []
        for _ in range(9):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = ops.correlation2d(a, b, 4); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[til] = (sorted(ts)[4], out)
    err = (res["32"][1] - res["48"][1]).abs().max().item()
    gb = 4 * B * H * W * (2 * C + 81) / 1e9
    print(f"C={C} {H}x{W} B={B}: tiling32 {res['32'][0]:.4f} ms ({gb/res['32'][0]:.0f} GB/s)  tiling48 {res['48'][0]:.4f} ms ({gb/res['48'][0]:.0f} GB/s)  max|diff| {err:.2e}")
