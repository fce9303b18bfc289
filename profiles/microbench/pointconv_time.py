"""PointConv forward (b200_pointconv_fwd) at the feature-pyramid shapes of config 1, batch 32: time per call, precision 1 and 2."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rpeflow_b200 as b200
from rpeflow_b200 import ops, pointconv as pc
B = int(os.environ.get("BATCH", "32"))
dev = "cuda"
for (C, cout, N, S) in ((32, 64, 8192, 4096), (64, 96, 4096, 2048), (96, 128, 2048, 1024), (128, 192, 1024, 512), (64, 64, 4096, 4096)):
    xyz = torch.rand(B, 3, N, device=dev) * 10
    feat = torch.randn(B, C, N, device=dev)
    samp = xyz[:, :, :S].contiguous()
    knn = ops.k_nearest_neighbor(xyz, samp, 16)
    torch.manual_seed(0)
    w = {n: v.to(dev) for n, v in pc.pack_pointconv_weights(pc.PointConvDownSampling(C, cout)).items()}
    res = []
    for prec in (2, 1):
        for _ in range(2): b200.pointconv_forward(xyz, feat, samp, knn, w, prec)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); b200.pointconv_forward(xyz, feat, samp, knn, w, prec); e1.record(); e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        res.append(min(ts))
    gf = 2.0 * B * S * (16 * (C + 3) * cout + 16 * 16 * (C + 3)) / 1e9
    print(f"C={C}->{cout} N={N} S={S} B={B}: 3xTF32 {res[0]:.3f} ms ({gf/res[0]:.0f} dense GFLOP/ms)  TF32 {res[1]:.3f} ms")
