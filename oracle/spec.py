"""numpy front-end of oracle/spec.c (see oracle/__init__.py — test infrastructure only).

Every function takes/returns numpy arrays with the layouts of include/b200flow.h, so a parity test reads
``assert_equal(cuda_result, spec.fn(same inputs))``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile oracle/spec.c (and, when /root/reference is present, oracle/_ref) with oracle/Makefile."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "spec.c")):
            subprocess.run(["make", "-s", "-C", _HERE, path], check=True, stdout=subprocess.DEVNULL)
        _LIB = ctypes.CDLL(path)
        _LIB.orc_event_voxel_int.restype = ctypes.c_int64
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def fps(xyz, n_samples):
    """xyz [B,N,3] -> [B,n_samples] int64 (reference: models/csrc/wrapper.py:75-103)."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    out = np.empty((B, n_samples), np.int64)
    lib().orc_fps(_p(xyz), _p(out), B, N, n_samples)
    return out


def knn(input_xyz, query_xyz, k, fused=False):
    """input [B,M,D], query [B,Q,D] -> [B,Q,k] int64, ordered by (distance, index)."""
    input_xyz, query_xyz = _f32(input_xyz), _f32(query_xyz)
    B, M, D = input_xyz.shape
    Q = query_xyz.shape[1]
    assert D in (2, 3) and query_xyz.shape[2] == D and 1 <= k <= 32
    out = np.empty((B, Q, k), np.int64)
    lib().orc_knn(_p(input_xyz), _p(query_xyz), _p(out), B, M, Q, D, k, int(fused))
    return out


def corr2d_fwd(in1_nhwc, in2_nhwc, md):
    in1_nhwc, in2_nhwc = _f32(in1_nhwc), _f32(in2_nhwc)
    B, H, W, C = in1_nhwc.shape
    out = np.empty((B, (2 * md + 1) ** 2, H, W), np.float32)
    lib().orc_corr2d_fwd(_p(in1_nhwc), _p(in2_nhwc), _p(out), B, C, H, W, md)
    return out


def corr2d_bwd(gout, in1_nhwc, in2_nhwc, md):
    gout, in1_nhwc, in2_nhwc = _f32(gout), _f32(in1_nhwc), _f32(in2_nhwc)
    B, H, W, C = in1_nhwc.shape
    g1 = np.empty((B, C, H, W), np.float32)
    g2 = np.empty((B, C, H, W), np.float32)
    lib().orc_corr2d_bwd(_p(gout), _p(in1_nhwc), _p(in2_nhwc), _p(g1), _p(g2), B, C, H, W, md)
    return g1, g2


def gather_cf(data, idx):
    """data [B,C,N] (4-byte dtype), idx [B,...] -> [B,C,...]."""
    data = np.ascontiguousarray(data)
    assert data.dtype.itemsize == 4
    idx = _i64(idx)
    B, C, N = data.shape
    I = int(np.prod(idx.shape[1:], dtype=np.int64))
    out = np.empty((B, C, I), data.dtype)
    lib().orc_gather_cf(_p(data), _p(idx), _p(out), B, C, N, ctypes.c_int64(I))
    return out.reshape((B, C) + idx.shape[1:])


def gather_cl(data, idx):
    """data [B,N,C] (4-byte dtype), idx [B,...] -> [B,...,C]."""
    data = np.ascontiguousarray(data)
    assert data.dtype.itemsize == 4
    idx = _i64(idx)
    B, N, C = data.shape
    I = int(np.prod(idx.shape[1:], dtype=np.int64))
    out = np.empty((B, I, C), data.dtype)
    lib().orc_gather_cl(_p(data), _p(idx), _p(out), B, C, N, ctypes.c_int64(I))
    return out.reshape((B,) + idx.shape[1:] + (C,))


def grid_sample_pts(feat, xy):
    feat, xy = _f32(feat), _f32(xy)
    B, C, H, W = feat.shape
    N = xy.shape[2]
    out = np.empty((B, C, N), np.float32)
    lib().orc_grid_sample_pts(_p(feat), _p(xy), _p(out), B, C, H, W, N)
    return out


def backwarp2d_border(x, flow):
    """x [B,C,H,W], flow [B,2,H,W] -> [B,C,H,W] (models/utils.py:186-198, padding_mode='border')."""
    x, flow = _f32(x), _f32(flow)
    B, C, H, W = x.shape
    out = np.empty_like(x)
    lib().orc_backwarp2d_border(_p(x), _p(flow), _p(out), B, C, H, W)
    return out


def convex_upsample(flow, mask, s):
    """flow [B,2,H,W], mask [B,9*s*s,H,W] -> [B,2,s*H,s*W] (models/utils.py:201-214)."""
    flow, mask = _f32(flow), _f32(mask)
    B, _, H, W = flow.shape
    out = np.empty((B, 2, H * s, W * s), np.float32)
    lib().orc_convex_upsample(_p(flow), _p(mask), _p(out), B, H, W, int(s))
    return out


def project_nn_corr(xy, feat2d, feat3d, nn):
    xy, feat2d, feat3d, nn = _f32(xy), _f32(feat2d), _f32(feat3d), _i64(nn)
    B, C2, H, W = feat2d.shape
    C3, N = feat3d.shape[1], feat3d.shape[2]
    out = np.empty((B, C3 + 3, H, W), np.float32)
    lib().orc_project_nn_corr(_p(xy), _p(feat2d), _p(feat3d), _p(nn), _p(out), B, C2, C3, H, W, N)
    return out


class _Corr3dWeights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in (
        "W1", "b1", "W2", "b2",
        "n1_Wa", "n1_ba", "n1_Wb", "n1_bb", "n1_Wc", "n1_bc",
        "n2_Wa", "n2_ba", "n2_Wb", "n2_bb", "n2_Wc", "n2_bc")]


CORR3D_WEIGHT_NAMES = [n for n, _ in _Corr3dWeights._fields_]


def corr3d_fwd(xyz1, feat1, xyz2, feat2, knn12, knn11, weights):
    """weights: dict name -> array, names = CORR3D_WEIGHT_NAMES (shapes in include/b200flow.h)."""
    xyz1, feat1, xyz2, feat2 = _f32(xyz1), _f32(feat1), _f32(xyz2), _f32(feat2)
    knn12, knn11 = _i64(knn12), _i64(knn11)
    B, Cin, N1 = feat1.shape
    N2 = feat2.shape[2]
    k = knn12.shape[2]
    keep = {n: _f32(weights[n]) for n in CORR3D_WEIGHT_NAMES}
    Cout = keep["W2"].shape[0]
    w = _Corr3dWeights(**{n: keep[n].ctypes.data for n in CORR3D_WEIGHT_NAMES})
    out = np.empty((B, Cout, N1), np.float32)
    lib().orc_corr3d_fwd(_p(xyz1), _p(feat1), _p(xyz2), _p(feat2), _p(knn12), _p(knn11), ctypes.byref(w),
                         _p(out), B, Cin, Cout, N1, N2, k)
    return out


def event_voxel_int(events, bins, H, W, polarity):
    """events [n,4] fp32 (x,y,t,p) -> ([bins*(2 if polarity else 1),H,W], n_out_of_range)."""
    events = _f32(events)
    vox = np.empty((bins * (2 if polarity else 1), H, W), np.float32)
    bad = lib().orc_event_voxel_int(_p(events), ctypes.c_int64(events.shape[0]), _p(vox), bins, H, W, int(polarity))
    return vox, int(bad)


def event_voxel_trilinear(x, y, t, p, bins, H, W, polarity):
    x, y, p, t = _f32(x), _f32(y), _f32(p), _i64(t)
    vox = np.empty((bins * (2 if polarity else 1), H, W), np.float32)
    lib().orc_event_voxel_trilinear(_p(x), _p(y), _p(t), _p(p), ctypes.c_int64(x.shape[0]), _p(vox),
                                    bins, H, W, int(polarity))
    return vox


def knn_interpolate(input_xyz, input_feat, query_xyz, idx):
    """input_xyz [B,3,M], input_feat [B,C,M], query_xyz [B,3,Q], idx [B,Q,k] -> [B,C,Q] (models/utils.py:140-156)."""
    input_xyz, input_feat, query_xyz, idx = _f32(input_xyz), _f32(input_feat), _f32(query_xyz), _i64(idx)
    B, C, M = input_feat.shape
    Q, k = idx.shape[1], idx.shape[2]
    out = np.empty((B, C, Q), np.float32)
    lib().orc_knn_interpolate(_p(input_xyz), _p(input_feat), _p(query_xyz), _p(idx), _p(out), B, C, M, Q, k)
    return out


def pointconv_fwd(xyz, feat, sampled_xyz, knn, w):
    """xyz [B,3,N], feat [B,C,N], sampled_xyz [B,3,S], knn [B,S,k]; w: dict Wa [8,3], ba, Wb [16,8], bb, L [out,16(C+3)],
    bias -> [B,out,S] (models/pointconv.py:33-61 / :90-122)."""
    xyz, feat, sampled_xyz, knn = _f32(xyz), _f32(feat), _f32(sampled_xyz), _i64(knn)
    ws = {n: _f32(w[n]) for n in ("Wa", "ba", "Wb", "bb", "L", "bias")}
    B, C, N = feat.shape
    S, k = knn.shape[1], knn.shape[2]
    cout = ws["L"].shape[0]
    assert ws["L"].shape[1] == 16 * (C + 3)
    out = np.empty((B, cout, S), np.float32)
    lib().orc_pointconv_fwd(_p(xyz), _p(feat), _p(sampled_xyz), _p(knn), _p(ws["Wa"]), _p(ws["ba"]), _p(ws["Wb"]),
                            _p(ws["bb"]), _p(ws["L"]), _p(ws["bias"]), _p(out), B, C, N, S, k, cout)
    return out
