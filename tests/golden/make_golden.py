"""Generate tests/golden/*.npz by running the UNMODIFIED reference (torch CPU path) on seeded inputs.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Each fixture stores the inputs and the reference's outputs, so the oracle (and, on the GPU, the CUDA path)
can be checked without the reference being present.  Recipe per fixture = the reference function named in
its key; file:line citations are in SURVEY.md §8(a).
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("RPEFLOW_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

# event_utils.py / dsec.py import data-loading packages that are absent here and unused by the voxelisers.
for m in ('hdf5plugin', 'h5py', 'imageio', 'skimage', 'omegaconf', 'matplotlib', 'matplotlib.colors'):
    sys.modules.setdefault(m, types.ModuleType(m))
sys.modules['h5py'].File = object
sys.modules['omegaconf'].OmegaConf = object
sys.modules['omegaconf'].DictConfig = dict
sys.modules['matplotlib.colors'].hsv_to_rgb = None

from models.csrc import correlation2d, furthest_point_sampling, k_nearest_neighbor   # noqa: E402
from models import utils as mutils                                                     # noqa: E402
from models.pwc3d_core import Correlation3D                                            # noqa: E402
import event_utils                                                                     # noqa: E402


def save(name, **arrays):
    arrays = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print(name, {k: v.shape for k, v in arrays.items()})


def g(seed):
    return torch.Generator().manual_seed(seed)


def corr2d():
    for tag, (B, C, H, W, md) in {"a": (2, 8, 12, 16, 4), "b": (1, 5, 7, 9, 2), "c": (1, 33, 6, 40, 4)}.items():
        gen = g(10)
        f1 = torch.randn(B, C, H, W, generator=gen, requires_grad=True)
        f2 = torch.randn(B, C, H, W, generator=gen, requires_grad=True)
        go = torch.rand(B, (2 * md + 1) ** 2, H, W, generator=gen)
        out = correlation2d(f1, f2, md)
        out.backward(go)
        save("corr2d_" + tag, feat1=f1, feat2=f2, md=md, out=out, grad_out=go, grad1=f1.grad, grad2=f2.grad)


def fps():
    gen = g(20)
    xyz = torch.rand(3, 257, 3, generator=gen)
    xyz[1, 100:140] = xyz[1, 0:40]                      # exact duplicates (datasets resample with replacement)
    xyz[2] = xyz[2] * torch.tensor([30.0, 18.0, 20.0]) + torch.tensor([-15.0, -9.0, 100.0])   # IDS-like magnitudes
    save("fps", xyz=xyz, n_samples=96, idx=furthest_point_sampling(xyz, 96))


def knn():
    gen = g(30)
    inp = torch.rand(2, 300, 3, generator=gen)
    qry = torch.rand(2, 130, 3, generator=gen)
    save("knn3d", input=inp, query=qry, k=16, idx=k_nearest_neighbor(inp, qry, 16))
    save("knn3d_k3_cf", input=inp.transpose(1, 2), query=qry.transpose(1, 2), k=3,
         idx=k_nearest_neighbor(inp.transpose(1, 2).contiguous(), qry.transpose(1, 2).contiguous(), 3))
    H, W = 9, 15
    pts = torch.rand(2, 2, 64, generator=gen) * torch.tensor([W + 2.0, H + 2.0]).view(1, 2, 1) - 1.0
    grid = mutils.mesh_grid(2, H, W, 'cpu').reshape(2, 2, -1)
    save("knn2d", input=pts, query=grid, k=1, idx=k_nearest_neighbor(pts, grid, 1))


def gathers():
    gen = g(40)
    data = torch.randn(2, 7, 50, generator=gen)
    idx = torch.randint(0, 50, (2, 20, 4), generator=gen)
    save("gather_cf", data=data, idx=idx, out=mutils.batch_indexing_channel_first(data, idx))
    data_l = torch.randn(2, 50, 7, generator=gen)
    save("gather_cl", data=data_l, idx=idx, out=mutils.batch_indexing_channel_last(data_l, idx))


def projection():
    gen = g(50)
    B, C2, C3, H, W, N = 2, 6, 5, 9, 13, 70
    feat2d = torch.randn(B, C2, H, W, generator=gen)
    feat3d = torch.randn(B, C3, N, generator=gen)
    xy = torch.rand(B, 2, N, generator=gen) * torch.tensor([W + 3.0, H + 3.0]).view(1, 2, 1) - 2.0   # some outside
    xy[0, :, 0] = torch.tensor([3.0, 4.0])             # exactly on a pixel centre
    xy[0, :, 1] = torch.tensor([W - 1.0, H - 1.0])     # last pixel
    save("grid_sample", feat=feat2d, xy=xy, out=mutils.grid_sample_wrapper(feat2d, xy))
    grid = mutils.mesh_grid(B, H, W, 'cpu').reshape(B, 2, -1)
    nn = k_nearest_neighbor(xy, grid, 1)[..., 0]
    save("project_nn_corr", xy=xy, feat2d=feat2d, feat3d=feat3d, nn=nn,
         out=mutils.project_feat_with_nn_corr(xy, feat2d, feat3d, nn))


def corr3d():
    names = {"W1": "cost_mlp.convs.0", "W2": "cost_mlp.convs.1",
             "n1_Wa": "weight_net1.convs.0", "n1_Wb": "weight_net1.convs.1", "n1_Wc": "weight_net1.convs.2",
             "n2_Wa": "weight_net2.convs.0", "n2_Wb": "weight_net2.convs.1", "n2_Wc": "weight_net2.convs.2"}
    for tag, (B, C, N, k) in {"a": (2, 16, 96, 16), "b": (1, 24, 40, 8)}.items():
        torch.manual_seed(60)
        mod = Correlation3D(C, C, k=k).eval()
        sd = mod.state_dict()
        gen = g(61)
        xyz1 = torch.rand(B, 3, N, generator=gen)
        xyz2 = xyz1 + 0.05 * torch.randn(B, 3, N, generator=gen)
        f1 = torch.randn(B, C, N, generator=gen)
        f2 = torch.randn(B, C, N, generator=gen)
        knn11 = k_nearest_neighbor(xyz1, xyz1, k)
        knn12 = k_nearest_neighbor(xyz2, xyz1, k)
        with torch.no_grad():
            out = mod(xyz1, f1, xyz2, f2, knn11)
        wts = {}
        for short, long in names.items():
            wts[short] = sd[long + ".conv_fn.weight"][:, :, 0, 0]
            wts[short.replace("W", "b")] = sd[long + ".conv_fn.bias"]
        save("corr3d_" + tag, xyz1=xyz1, feat1=f1, xyz2=xyz2, feat2=f2, knn11=knn11, knn12=knn12, k=k, out=out,
             **{"w_" + n: v for n, v in wts.items()})


def events():
    rng = np.random.default_rng(70)
    H, W, n = 20, 30, 4000
    ev = np.zeros((n, 4), np.float32)
    ev[:, 0] = rng.integers(0, W, n)
    ev[:, 1] = rng.integers(0, H, n)
    ev[:, 2] = np.sort(rng.random(n).astype(np.float32)) * 0.05 + 3.0
    ev[:, 3] = rng.choice([-1.0, 1.0], n)
    save("event_voxel_pol", events=ev, bins=10, H=H, W=W,
         vox=event_utils.eventsToVoxel(ev, num_bins=10, height=H, width=W, event_polarity=True))
    save("event_voxel_nopol", events=ev, bins=5, H=H, W=W,
         vox=event_utils.eventsToVoxel(ev, num_bins=5, height=H, width=W, event_polarity=False))

    import dsec                                          # NB: sets torch.set_num_threads(1)
    obj = dsec.DSECTrain.__new__(dsec.DSECTrain)
    d = {"x": (rng.random(n) * (W + 1.5) - 0.75).astype(np.float32),     # some outside the sensor
         "y": (rng.random(n) * (H + 1.5) - 0.75).astype(np.float32),
         "t": np.sort(rng.integers(1_000_000, 1_100_000, n)).astype(np.int64),
         "p": rng.integers(0, 2, n).astype(np.uint8)}
    save("event_trilinear_pol", bins=10, H=H, W=W, **d,
         vox=obj.eventsToVoxelInter(d, 10, H, W, event_polarity=True))
    save("event_trilinear_nopol", bins=5, H=H, W=W, **d,
         vox=obj.eventsToVoxelInter(d, 5, H, W, event_polarity=False))


def pointconv():
    from models.pointconv import PointConvDownSampling, PointConvNoSampling
    for tag, cls, (B, C, cout, N, S, k) in (("down", PointConvDownSampling, (2, 13, 24, 120, 60, 16)),
                                            ("nosample", PointConvNoSampling, (1, 32, 32, 70, 70, 16))):
        torch.manual_seed(80)
        mod = cls(C, cout, norm=None, k=k).eval()
        gen = g(81)
        xyz = torch.rand(B, 3, N, generator=gen) * 3
        feat = torch.randn(B, C, N, generator=gen)
        sd = mod.state_dict()
        if cls is PointConvDownSampling:
            sampled = xyz[:, :, :S].contiguous()
            idx = k_nearest_neighbor(xyz, sampled, k)
            with torch.no_grad():
                out = mod(xyz, feat, sampled)
        else:
            sampled = xyz
            idx = k_nearest_neighbor(xyz, xyz, k)
            with torch.no_grad():
                out = mod(xyz, feat, idx)
        save("pointconv_" + tag, xyz=xyz, feat=feat, sampled=sampled, knn=idx, out=out,
             w_Wa=sd["weight_net.convs.0.conv_fn.weight"][:, :, 0, 0], w_ba=sd["weight_net.convs.0.conv_fn.bias"],
             w_Wb=sd["weight_net.convs.1.conv_fn.weight"][:, :, 0, 0], w_bb=sd["weight_net.convs.1.conv_fn.bias"],
             w_L=sd["linear.weight"], w_bias=sd["linear.bias"])


def interpolation():
    gen = g(70)
    xyz_in = torch.rand(2, 3, 300, generator=gen) * 4
    feat = torch.randn(2, 7, 300, generator=gen)
    xyz_q = torch.rand(2, 3, 450, generator=gen) * 4
    xyz_q[0, :, :5] = xyz_in[0, :, :5]                      # coincident points: the 1e-8 clamp decides the weights
    idx = k_nearest_neighbor(xyz_in, xyz_q, 3)
    out = mutils.knn_interpolation(xyz_in, feat, xyz_q, k=3)
    flow = 0.1 * torch.randn(2, 3, 300, generator=gen)
    xyz2 = xyz_in + flow + 0.01 * torch.randn(2, 3, 300, generator=gen)
    warp = mutils.backwarp_3d(xyz_in, xyz2, flow, k=3)
    save("knn_interpolation", input_xyz=xyz_in, input_feat=feat, query_xyz=xyz_q, idx=idx, out=out,
         xyz2=xyz2, flow12=flow, backwarp=warp)


def warp2d():
    """f3: backwarp_2d (border) and the decode expression leaky_relu(correlation2d(f1, backwarp_2d(f2, flow)), 0.1)."""
    from torch.nn.functional import leaky_relu
    gen = torch.Generator().manual_seed(77)
    f1 = torch.randn(2, 8, 12, 20, generator=gen)
    f2 = torch.randn(2, 8, 12, 20, generator=gen)
    flow = 3.0 * torch.randn(2, 2, 12, 20, generator=gen)        # many samples leave the image: the border clamp decides
    flow[0, :, 0, 0] = torch.tensor([-50.0, -50.0])
    flow[0, :, 11, 19] = torch.tensor([50.0, 50.0])
    flow[1, :, 5, 5] = 0.0                                        # exact pixel hit
    warped = mutils.backwarp_2d(f2, flow, padding_mode='border')
    cost = leaky_relu(correlation2d(f1, warped, 4), 0.1)
    save("warp2d", f1=f1, f2=f2, flow=flow, warped=warped, cost=cost)


def upsample():
    """f4: convex_upsample at the model's scale (4, RPEFlow_core.py:424) and at the default 8."""
    gen = torch.Generator().manual_seed(88)
    for s, (h, w) in ((4, (9, 14)), (8, (5, 6))):
        flow = 5.0 * torch.randn(2, 2, h, w, generator=gen)
        mask = 3.0 * torch.randn(2, 9 * s * s, h, w, generator=gen)
        save("convex_upsample_s%d" % s, flow=flow, mask=mask, s=s, out=mutils.convex_upsample(flow, mask, scale_factor=s))


if __name__ == "__main__":
    only = sys.argv[1:]
    for fn in (corr2d, fps, knn, gathers, projection, corr3d, events, interpolation, pointconv, warp2d, upsample):
        if not only or fn.__name__ in only:
            fn()
