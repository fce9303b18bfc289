"""Which re-bound op moves the model's flows?  Arm A = the reference's torch path on the GPU; then install() one op group at a
time (everything else stays the reference's torch code) and report max|flow - flow_A|."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from refmodel_util import reference_root, reference_extensions_bound
from rpeflow_b200 import refhost
import rpeflow_b200.install as inst

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda", 0)
model = refhost.build_rpeflow(reference_root(), device=dev, install=False, seed=0)
host = refhost.synthetic_model_inputs(2, 540, 960, 8192, seed=0)
inputs = {k: v.to(dev) for k, v in host.items()}
a = refhost.forward(model, inputs)
res = {"scale": {k: a[k].abs().max().item() for k in a}}
def diff(o):
    return {k: (o[k] - a[k]).abs().max().item() for k in a}
res["A_again"] = diff(refhost.forward(model, inputs))
with reference_extensions_bound():
    res["C_ref_cuda_ext"] = diff(refhost.forward(model, inputs))
groups = {"ext_shims_only": (None, True), "k_nearest_neighbor": ({"k_nearest_neighbor"}, False), "furthest_point_sampling": ({"furthest_point_sampling"}, False),
          "correlation2d": ({"correlation2d"}, False), "gathers": ({"batch_indexing_channel_first", "batch_indexing_channel_last"}, False),
          "grid_sample_wrapper": ({"grid_sample_wrapper"}, False), "project_feat_with_nn_corr": ({"project_feat_with_nn_corr"}, False),
          "knn_interpolation+backwarp_3d": ({"knn_interpolation", "backwarp_3d"}, False), "backwarp_2d": ({"backwarp_2d"}, False),
          "convex_upsample": ({"convex_upsample"}, False), "Correlation3D": ({"Correlation3D"}, False), "PointConv": ({"PointConv"}, False),
          "CorrFeatureFuser3D": ({"CorrFeatureFuser3D"}, False), "all": (None, True)}
for name, (only, ext) in groups.items():
    if name == "ext_shims_only":
        inst.install(patch_python=False)
    else:
        inst.install(patch_python=True, only=only, extensions=ext)
    try:
        res[name] = diff(refhost.forward(model, inputs))
    finally:
        inst.uninstall()
    print(name, res[name], flush=True)
# precision ablation for the tensor-core ops: fp32 path
import models.pwc3d_core as p3, models.pointconv as pc
for cls in (p3.Correlation3D, pc.PointConvDownSampling, pc.PointConvNoSampling):
    cls.b200_precision = 0 if cls is p3.Correlation3D else 2
inst.install(patch_python=True, only={"Correlation3D"}, extensions=False)
res["Correlation3D_fp32_path"] = diff(refhost.forward(model, inputs)); inst.uninstall()
print(json.dumps(res, indent=1))
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r2_model_ablation.json"), "w"), indent=1)
