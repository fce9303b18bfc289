"""One PointConv call shape for ncu captures: SHAPE=C,cout,N,S  BATCH=74  PREC=2."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import rpeflow_b200 as b200
from rpeflow_b200 import ops, pointconv as pc
C, cout, N, S = (int(v) for v in os.environ.get("SHAPE", "195,128,1024,1024").split(","))
B, prec = int(os.environ.get("BATCH", "74")), int(os.environ.get("PREC", "2"))
xyz = torch.rand(B, 3, N, device="cuda") * 10
feat = torch.randn(B, C, N, device="cuda")
samp = xyz[:, :, :S].contiguous()
knn = ops.k_nearest_neighbor(xyz, samp, 16)
w = {n: v.cuda() for n, v in pc.pack_pointconv_weights(pc.PointConvDownSampling(C, cout)).items()}
for _ in range(2):
    b200.pointconv_forward(xyz, feat, samp, knn, w, prec)
torch.cuda.synchronize()
