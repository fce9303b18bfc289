"""Drop-in for ``models/pointconv.py`` (danqu130/RPEFlow): ``PointConvDownSampling`` (:7-61) and
``PointConvNoSampling`` (:64-122) — SURVEY §8f rank 1.

Both classes keep the reference's constructor and ``state_dict`` keys (``weight_net.convs.{0,1}.conv_fn.{weight,bias}``,
``linear.{weight,bias}``), so released checkpoints load unchanged.  ``forward`` is the cell-grid KNN (b200_knn_grid)
followed by the fused tcgen05 kernel of ``b200_pointconv_fwd``: gathers, weight net, the [16 x k].[k x (C+3)] product,
the Linear layer and the LeakyReLU in one pass, no [B,S,k,*] tensor in memory.

Forward only (the reference trains through these layers; training keeps ``models.pointconv``).  Built for the
configuration the model uses: ``norm=None``, ``activation='leaky_relu'``, ``k=16``.
"""
import ctypes

import torch
import torch.nn as nn

from ._lib import PointConvWeights, check, lib
from .ops import k_nearest_neighbor

__all__ = ["PointConvDownSampling", "PointConvNoSampling", "pointconv_forward", "pack_pointconv_weights"]


class _Conv(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv_fn = nn.Conv2d(cin, cout, 1)


class _MLP(nn.Module):
    def __init__(self, cin, widths):
        super().__init__()
        chans = [cin] + list(widths)
        self.convs = nn.ModuleList(_Conv(a, b) for a, b in zip(chans[:-1], chans[1:]))


def pack_pointconv_weights(module):
    """dict name -> contiguous fp32 tensors in the order of b200_pointconv_weights."""
    def wb(conv):
        w = conv.conv_fn.weight
        return w.detach().reshape(w.shape[0], w.shape[1]).contiguous().float(), conv.conv_fn.bias.detach().contiguous().float()
    out = {}
    out["Wa"], out["ba"] = wb(module.weight_net.convs[0])
    out["Wb"], out["bb"] = wb(module.weight_net.convs[1])
    out["L"] = module.linear.weight.detach().contiguous().float()
    out["bias"] = module.linear.bias.detach().contiguous().float()
    return out


def pointconv_forward(xyz, features, sampled_xyz, knn_indices, weights, precision=2):
    """xyz [B,3,N], features [B,C,N], sampled_xyz [B,3,S], knn_indices [B,S,16] -> [B,out,S]; all CUDA."""
    if not (xyz.is_cuda and features.is_cuda and sampled_xyz.is_cuda):
        raise RuntimeError("rpeflow_b200.pointconv_forward: CUDA tensors required — no CPU/torch fallback")
    xyz, features, sampled_xyz = (t.contiguous().float() for t in (xyz, features, sampled_xyz))
    knn = knn_indices.to(torch.int64).contiguous()
    B, C, N = features.shape
    S, k = sampled_xyz.shape[2], knn.shape[2]
    cout = weights["L"].shape[0]
    assert tuple(weights["L"].shape) == (cout, 16 * (C + 3)) and tuple(knn.shape) == (B, S, k)
    keep = {n: weights[n].to(xyz.device) for n in PointConvWeights.NAMES}
    w = PointConvWeights(**{n: keep[n].data_ptr() for n in PointConvWeights.NAMES})
    out = torch.empty((B, cout, S), dtype=torch.float32, device=xyz.device)
    scratch = torch.empty((lib.b200_pointconv_scratch_floats(B, C, cout, N),), dtype=torch.float32, device=xyz.device)
    with torch.cuda.device(xyz.device):
        check(lib.b200_pointconv_fwd(xyz.data_ptr(), features.data_ptr(), sampled_xyz.data_ptr(), knn.data_ptr(),
                                     ctypes.byref(w), out.data_ptr(), scratch.data_ptr(), B, C, cout, N, S, k,
                                     int(precision), torch.cuda.current_stream(xyz.device).cuda_stream),
              "b200_pointconv_fwd")
    return out


class _PointConv(nn.Module):
    def __init__(self, in_channels, out_channels, norm=None, activation='leaky_relu', k=16, precision=2):
        super().__init__()
        if norm is not None or activation != 'leaky_relu':
            raise NotImplementedError("rpeflow_b200 PointConv is built for norm=None, activation='leaky_relu' (what RPEFlow uses)")
        self.k = k
        self.precision = precision
        self.weight_net = _MLP(3, [8, 16])
        self.linear = nn.Linear(16 * (in_channels + 3), out_channels)

    def _check_grad(self, features):
        if torch.is_grad_enabled() and (features.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("rpeflow_b200 PointConv is forward-only; run it under torch.no_grad() "
                               "(training keeps models.pointconv)")


class PointConvDownSampling(_PointConv):
    def forward(self, xyz, features, sampled_xyz):
        """xyz [B,3,N], features [B,C,N], sampled_xyz [B,3,S] -> [B,out,S] (pointconv.py:33-61)."""
        self._check_grad(features)
        knn = k_nearest_neighbor(xyz, sampled_xyz, self.k)                       # pointconv.py:46
        return pointconv_forward(xyz, features, sampled_xyz, knn, pack_pointconv_weights(self), self.precision)


class PointConvNoSampling(_PointConv):
    def forward(self, xyz, features, knn_indices=None):
        """xyz [B,3,N], features [B,C,N], knn_indices [B,N,>=k] or None -> [B,out,N] (pointconv.py:90-122)."""
        self._check_grad(features)
        if knn_indices is not None:
            assert knn_indices.shape[:2] == torch.Size([xyz.shape[0], xyz.shape[2]]) and knn_indices.shape[2] >= self.k
            knn_indices = knn_indices[:, :, :self.k]
        else:
            knn_indices = k_nearest_neighbor(xyz, xyz, self.k)                   # pointconv.py:107
        return pointconv_forward(xyz, features, xyz, knn_indices, pack_pointconv_weights(self), self.precision)
