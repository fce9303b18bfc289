"""clock64 trace of pass 2 (thin CTA) at C = 32: where does a tile's time go?  Needs the trace build:
bash profiles/build_variant.sh trace "-DV2_TRACE"; B200FLOW_LIB=gpurun_variants/libb200flow_trace.so python profiles/microbench/corr3d_trace.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import rpeflow_b200 as b200
from rpeflow_b200 import pwc3d, _lib
dev = torch.device("cuda", 0)
B = 74
for C, N in [(32, 4096), (64, 2048), (128, 512)]:
    torch.manual_seed(0)
    mod = pwc3d.Correlation3D(C, C, k=16)
    w = {n: v.to(dev) for n, v in pwc3d.pack_weights(mod).items()}
    xyz1 = torch.rand(B, 3, N, device=dev); xyz2 = xyz1 + 0.02 * torch.randn(B, 3, N, device=dev)
    f1 = torch.randn(B, C, N, device=dev); f2 = torch.randn(B, C, N, device=dev)
    knn11 = b200.k_nearest_neighbor(xyz1, xyz1, 16); knn12 = b200.k_nearest_neighbor(xyz2, xyz1, 16)
    for _ in range(3):
        pwc3d.correlation3d_forward(xyz1, f1, xyz2, f2, w, knn12, knn11, 2)
    torch.cuda.synchronize()
    buf = np.zeros(8 * 512, np.int64)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    rc = lib.b200_debug_read_v2_trace(buf.ctypes.data_as(ctypes.c_void_p), buf.size)
    t = buf.reshape(512, 8)[8:72]           # skip the first tiles (cold)
    names = ["meta loads", "hidden+sync+HID", "K blocks (gather, convert, arrive)", "(gap)", "wait accumulator", "epilogue"]
    d = np.diff(t[:, :7], axis=1).astype(float)
    tile = np.diff(t[:, 0]).mean()
    print(f"C={C}: {tile:.0f} cycles per tile per CTA; phases (mean cycles): " + ", ".join(f"{n} {v:.0f}" for n, v in zip(names, d.mean(0))), flush=True)
