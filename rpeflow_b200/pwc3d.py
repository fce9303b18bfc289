"""Drop-in for ``models/pwc3d_core.py::Correlation3D`` (danqu130/RPEFlow, :60-117): the 3-D point cost volume.

``Correlation3D`` below keeps the reference's constructor, parameter tree and ``state_dict`` keys
(``cost_mlp.convs.{0,1}.conv_fn.{weight,bias}``, ``weight_net{1,2}.convs.{0,1,2}.conv_fn.{weight,bias}``) so
released checkpoints load unchanged, and the same ``forward(xyz1, feat1, xyz2, feat2, knn_indices_1in1=None)``.
The whole forward is the fused sm_100a path of ``b200_corr3d_fwd`` (+ ``b200_knn`` for the neighbour search).

Forward only: the fused kernels do not build an autograd graph.  ``forward`` therefore refuses to run while
gradients are enabled for any input/parameter that requires them (training keeps the torch module).
"""
import ctypes

import torch
import torch.nn as nn

from ._lib import Corr3dWeights, check, lib
from .ops import k_nearest_neighbor

__all__ = ["Correlation3D", "correlation3d_forward", "pack_weights", "build_pc_pyramid"]


class _Conv(nn.Module):
    """1x1 Conv2d + activation with the reference's attribute names (models/utils.py:37-64, norm=None)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv_fn = nn.Conv2d(cin, cout, 1)


class _MLP(nn.Module):
    """models/utils.py:84-98 (MLP2d): only the parameter container is needed here."""

    def __init__(self, cin, widths):
        super().__init__()
        chans = [cin] + list(widths)
        self.convs = nn.ModuleList(_Conv(a, b) for a, b in zip(chans[:-1], chans[1:]))


def pack_weights(module):
    """dict name -> contiguous fp32 [out,in] / [out] tensors in the order of b200_corr3d_weights."""
    def wb(conv):
        w = conv.conv_fn.weight
        return w.detach().reshape(w.shape[0], w.shape[1]).contiguous().float(), conv.conv_fn.bias.detach().contiguous().float()
    out = {}
    out["W1"], out["b1"] = wb(module.cost_mlp.convs[0])
    out["W2"], out["b2"] = wb(module.cost_mlp.convs[1])
    for tag, net in (("n1", module.weight_net1), ("n2", module.weight_net2)):
        for letter, conv in zip("abc", net.convs):
            out[f"{tag}_W{letter}"], out[f"{tag}_b{letter}"] = wb(conv)
    return out


def correlation3d_forward(xyz1, feat1, xyz2, feat2, weights, knn12, knn11, precision=2):
    """Functional form: all tensors CUDA fp32; weights = dict from pack_weights(); returns [B,Cout,N1].

    precision: arithmetic of the Cout x Cout cost_mlp layer — 0 = fp32 FFMA, 1 = TF32 tensor cores (tcgen05; what cuDNN
    does for the reference's 1x1 convs under torch's default allow_tf32), 2 = 3xTF32 tensor cores (default; within the
    fp32 path's 1e-4 tolerance).  Shapes the tensor-core kernel does not cover (k != 16, Cout % 32 != 0) run fp32."""
    xyz1, feat1, xyz2, feat2 = (t.contiguous().float() for t in (xyz1, feat1, xyz2, feat2))
    if not xyz1.is_cuda:
        raise RuntimeError("rpeflow_b200.correlation3d_forward: CUDA tensors required — no CPU/torch fallback")
    knn12 = knn12.to(torch.int64).contiguous()
    knn11 = knn11.to(torch.int64).contiguous()
    B, Cin, N1 = feat1.shape
    N2 = feat2.shape[2]
    k = knn12.shape[2]
    Cout = weights["W2"].shape[0]
    assert tuple(weights["W1"].shape) == (Cout, 2 * Cin + 3)
    assert tuple(knn11.shape) == (B, N1, k) and tuple(knn12.shape) == (B, N1, k)
    keep = {n: weights[n].to(xyz1.device) for n in Corr3dWeights.NAMES}
    w = Corr3dWeights(**{n: keep[n].data_ptr() for n in Corr3dWeights.NAMES})
    out = torch.empty((B, Cout, N1), dtype=torch.float32, device=xyz1.device)
    n_scratch = lib.b200_corr3d_scratch_floats(B, Cin, Cout, N1, N2, k)
    scratch = torch.empty((n_scratch,), dtype=torch.float32, device=xyz1.device)
    with torch.cuda.device(xyz1.device):
        check(lib.b200_corr3d_fwd(xyz1.data_ptr(), feat1.data_ptr(), xyz2.data_ptr(), feat2.data_ptr(),
                                  knn12.data_ptr(), knn11.data_ptr(), ctypes.byref(w), out.data_ptr(),
                                  scratch.data_ptr(), B, Cin, Cout, N1, N2, k, int(precision),
                                  torch.cuda.current_stream(xyz1.device).cuda_stream), "b200_corr3d_fwd")
    return out


class Correlation3D(nn.Module):
    def __init__(self, in_channels, out_channels, k=16, precision=2):
        super().__init__()
        self.k = k
        self.precision = precision
        self.cost_mlp = _MLP(3 + 2 * in_channels, [out_channels, out_channels])
        self.weight_net1 = _MLP(3, [8, 8, out_channels])
        self.weight_net2 = _MLP(3, [8, 8, out_channels])

    def forward(self, xyz1, feat1, xyz2, feat2, knn_indices_1in1=None):
        """xyz* [B,3,N], feat* [B,C,N], knn_indices_1in1 [B,N,k] or None -> [B,Cout,N]."""
        if torch.is_grad_enabled() and (feat1.requires_grad or feat2.requires_grad or
                                        any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("rpeflow_b200.Correlation3D is forward-only; run it under torch.no_grad() "
                               "(training keeps models.pwc3d_core.Correlation3D)")
        knn12 = k_nearest_neighbor(input_xyz=xyz2, query_xyz=xyz1, k=self.k)        # pwc3d_core.py:81
        if knn_indices_1in1 is None:
            knn_indices_1in1 = k_nearest_neighbor(input_xyz=xyz1, query_xyz=xyz1, k=self.k)   # :104
        else:
            assert knn_indices_1in1.shape == torch.Size([feat1.shape[0], feat1.shape[2], self.k])
        return correlation3d_forward(xyz1, feat1, xyz2, feat2, pack_weights(self), knn12, knn_indices_1in1,
                                     self.precision)


def build_pc_pyramid(pc1, pc2, n_samples_list):
    """models/pwc3d_core.py:8-28: one FPS call on cat[pc1,pc2] for max(n_samples_list); levels are prefixes."""
    from .ops import furthest_point_sampling
    from .projection import batch_indexing_channel_first
    batch_size, _, n_points = pc1.shape
    both = torch.cat([pc1, pc2], dim=0)
    picked = furthest_point_sampling(both.transpose(1, 2), max(n_samples_list))
    idx1, idx2 = picked[:batch_size], picked[batch_size:]
    lv0 = torch.arange(n_points, device=pc1.device)[None, :].expand(batch_size, n_points)
    xyzs1, xyzs2, ids1, ids2 = [pc1], [pc2], [lv0], [lv0]
    for n in n_samples_list:
        ids1.append(idx1[:, :n])
        ids2.append(idx2[:, :n])
        xyzs1.append(batch_indexing_channel_first(pc1, idx1[:, :n]))
        xyzs2.append(batch_indexing_channel_first(pc2, idx2[:, :n]))
    return xyzs1, xyzs2, ids1, ids2
