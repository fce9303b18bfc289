"""Wire the sm_100a kernels into an UNMODIFIED checkout of danqu130/RPEFlow.

    import sys; sys.path.insert(0, "/path/to/RPEFlow")
    import rpeflow_b200.install as b200; b200.install()        # BEFORE the first `import models`
    from models.RPEFlow import RPEFlow                          # wrapper.py now finds its three extensions

What ``install()`` does
1. registers ``models.csrc._correlation_cuda``, ``models.csrc._furthest_point_sampling_cuda`` and
   ``models.csrc._k_nearest_neighbor_cuda`` in ``sys.modules`` (the shims in rpeflow_b200/shims), so the
   byte-for-byte unchanged ``models/csrc/wrapper.py:4-8`` imports them instead of falling back to torch;
2. (``patch_python_ops=True``) re-binds the pure-python hot ops that have no extension boundary in the
   reference — ``batch_indexing_channel_{first,last}``, ``grid_sample_wrapper``, ``project_feat_with_nn_corr``
   (models/utils.py) in every module that imported them by name, ``Correlation3D.forward``
   (models/pwc3d_core.py:69) for inference, ``event_utils.eventsToVoxel`` and
   ``dsec.DSECTrain.eventsToVoxelInter``.  Gathers keep the torch path for tensors that require grad.
"""
import importlib
import sys

import torch

from . import events, ops, pointconv, projection, pwc3d
from .shims import _correlation_cuda, _furthest_point_sampling_cuda, _k_nearest_neighbor_cuda

_SHIMS = {
    "models.csrc._correlation_cuda": _correlation_cuda,
    "models.csrc._furthest_point_sampling_cuda": _furthest_point_sampling_cuda,
    "models.csrc._k_nearest_neighbor_cuda": _k_nearest_neighbor_cuda,
}


def register_extension_shims():
    if "models.csrc.wrapper" in sys.modules:
        raise RuntimeError("rpeflow_b200.install() must run before `models.csrc` is first imported "
                           "(wrapper.py binds the extension symbols at import time)")
    for name, mod in _SHIMS.items():
        sys.modules[name] = mod


def _grad_aware(fast, slow):
    """Use the CUDA kernel unless autograd needs to differentiate through the op or the data is not on a GPU."""
    def op(data, *rest):
        if not data.is_cuda or (torch.is_grad_enabled() and data.requires_grad):
            return slow(data, *rest)
        return fast(data, *rest)
    op.__name__ = fast.__name__
    op.__doc__ = fast.__doc__
    return op


def _corr3d_forward(self, xyz1, feat1, xyz2, feat2, knn_indices_1in1=None):
    needs_graph = torch.is_grad_enabled() and (feat1.requires_grad or feat2.requires_grad or
                                               any(p.requires_grad for p in self.parameters()))
    if needs_graph or not feat1.is_cuda:
        return self._b200_reference_forward(xyz1, feat1, xyz2, feat2, knn_indices_1in1)
    knn12 = ops.k_nearest_neighbor(input_xyz=xyz2, query_xyz=xyz1, k=self.k)
    if knn_indices_1in1 is None:
        knn_indices_1in1 = ops.k_nearest_neighbor(input_xyz=xyz1, query_xyz=xyz1, k=self.k)
    return pwc3d.correlation3d_forward(xyz1, feat1, xyz2, feat2, pwc3d.pack_weights(self), knn12, knn_indices_1in1,
                                       getattr(self, "b200_precision", 2))


def _pointconv_fast_ok(self, features):
    """The fused kernel covers the configuration RPEFlow uses; anything else (or autograd) keeps the reference forward."""
    import torch.nn as nn
    needs_graph = torch.is_grad_enabled() and (features.requires_grad or any(p.requires_grad for p in self.parameters()))
    return (features.is_cuda and not needs_graph and self.k == 16 and isinstance(self.norm_fn, nn.Identity)
            and isinstance(self.activation_fn, nn.LeakyReLU) and abs(self.activation_fn.negative_slope - 0.1) < 1e-12
            and self.linear.out_features <= 256)


def _pointconv_down_forward(self, xyz, features, sampled_xyz):
    if not _pointconv_fast_ok(self, features):
        return self._b200_reference_forward(xyz, features, sampled_xyz)
    knn = ops.k_nearest_neighbor(xyz, sampled_xyz, self.k)
    return pointconv.pointconv_forward(xyz, features, sampled_xyz, knn, pointconv.pack_pointconv_weights(self),
                                       getattr(self, "b200_precision", 2))


def _pointconv_nosample_forward(self, xyz, features, knn_indices=None):
    if not _pointconv_fast_ok(self, features):
        return self._b200_reference_forward(xyz, features, knn_indices)
    knn = knn_indices[:, :, :self.k] if knn_indices is not None else ops.k_nearest_neighbor(xyz, xyz, self.k)
    return pointconv.pointconv_forward(xyz, features, xyz, knn, pointconv.pack_pointconv_weights(self),
                                       getattr(self, "b200_precision", 2))


def patch_python_ops(patch_events=True):
    mutils = importlib.import_module("models.utils")
    wrapper = importlib.import_module("models.csrc.wrapper")

    def correlation2d(input1, input2, max_displacement, cpp_impl=True):
        """wrapper.py:55-72 with the NCHW fast path (no permutes) when autograd is not involved."""
        if input1.is_cuda and cpp_impl and not (torch.is_grad_enabled() and (input1.requires_grad or input2.requires_grad)):
            return ops.correlation2d(input1, input2, max_displacement)
        return wrapper_correlation2d(input1, input2, max_displacement, cpp_impl)
    wrapper_correlation2d = wrapper.correlation2d
    replaced = {
        "correlation2d": correlation2d,
        "batch_indexing_channel_first": _grad_aware(projection.batch_indexing_channel_first,
                                                    mutils.batch_indexing_channel_first),
        "batch_indexing_channel_last": _grad_aware(projection.batch_indexing_channel_last,
                                                   mutils.batch_indexing_channel_last),
        "grid_sample_wrapper": _grad_aware(projection.grid_sample_wrapper, mutils.grid_sample_wrapper),
        "project_feat_with_nn_corr": projection.project_feat_with_nn_corr if torch.cuda.is_available()
        else mutils.project_feat_with_nn_corr,
    }
    ref_interp, ref_backwarp = mutils.knn_interpolation, mutils.backwarp_3d

    def knn_interpolation(input_xyz, input_features, query_xyz, k=3):
        if not input_xyz.is_cuda or k > 8 or (torch.is_grad_enabled() and (input_features.requires_grad or input_xyz.requires_grad
                                                                           or query_xyz.requires_grad)):
            return ref_interp(input_xyz, input_features, query_xyz, k)
        return projection.knn_interpolation(input_xyz, input_features, query_xyz, k)

    def backwarp_3d(xyz1, xyz2, flow12, k=3):
        if not xyz1.is_cuda or k > 8 or (torch.is_grad_enabled() and (flow12.requires_grad or xyz1.requires_grad
                                                                      or xyz2.requires_grad)):
            return ref_backwarp(xyz1, xyz2, flow12, k)
        return projection.backwarp_3d(xyz1, xyz2, flow12, k)
    ref_backwarp_2d = mutils.backwarp_2d

    def backwarp_2d(x, flow12, padding_mode):
        if not x.is_cuda or padding_mode != "border" or (torch.is_grad_enabled() and (x.requires_grad or flow12.requires_grad)):
            return ref_backwarp_2d(x, flow12, padding_mode)
        return projection.backwarp_2d(x, flow12, padding_mode)
    replaced["knn_interpolation"] = knn_interpolation
    replaced["backwarp_3d"] = backwarp_3d
    replaced["backwarp_2d"] = backwarp_2d
    ref_convex = mutils.convex_upsample

    def convex_upsample(flow, mask, scale_factor=8):
        if not flow.is_cuda or scale_factor not in (2, 4, 8) or (torch.is_grad_enabled() and (flow.requires_grad or mask.requires_grad)):
            return ref_convex(flow, mask, scale_factor)
        return projection.convex_upsample(flow, mask, scale_factor)
    replaced["convex_upsample"] = convex_upsample
    for modname in ("models.utils", "models.RPEFlow_core", "models.pwc2d_core", "models.pwc3d_core", "models.pointconv",
                    "models.losses3d", "models.RPEFlow", "models.csrc", "models.csrc.wrapper"):
        try:
            mod = importlib.import_module(modname)
        except Exception:
            continue
        for name, fn in replaced.items():
            if hasattr(mod, name):
                setattr(mod, name, fn)
    pcm = importlib.import_module("models.pointconv")
    for cls, fwd in ((pcm.PointConvDownSampling, _pointconv_down_forward), (pcm.PointConvNoSampling, _pointconv_nosample_forward)):
        if not hasattr(cls, "_b200_reference_forward"):
            cls._b200_reference_forward = cls.forward
            cls.forward = fwd
    core = importlib.import_module("models.pwc3d_core")
    if not hasattr(core.Correlation3D, "_b200_reference_forward"):
        core.Correlation3D._b200_reference_forward = core.Correlation3D.forward
        core.Correlation3D.forward = _corr3d_forward
    if patch_events:
        for modname, attr, fn in (("event_utils", "eventsToVoxel", events.eventsToVoxel),):
            try:
                setattr(importlib.import_module(modname), attr, fn)
            except Exception:
                pass
        try:
            dsec = importlib.import_module("dsec")
            dsec.DSECTrain.eventsToVoxelInter = lambda self, ev, num_bins, height, width, event_polarity=False: \
                events.eventsToVoxelInter(ev, num_bins, height, width, event_polarity)
        except Exception:
            pass


def install(patch_python=True, patch_events=False):
    register_extension_shims()
    if patch_python:
        patch_python_ops(patch_events=patch_events)
