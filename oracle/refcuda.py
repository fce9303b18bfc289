"""ctypes front-end of oracle/_ref/libref_kernels.so — the reference's own CUDA kernels (unmodified sources
from /root/reference/models/csrc, compiled for sm_100a by oracle/Makefile).  TEST INFRASTRUCTURE ONLY.

Used on the GPU box as a second checker ("does the new kernel agree with the kernel it replaces?") and as
the "kernel to beat" timing in profiles/.  Host-side allocation mirrors the reference's C++ wrappers
(correlation.cpp:15-17, furthest_point_sampling.cpp:11-12, k_nearest_neighbor.cpp:16).  The reference launches
on the legacy default stream; callers must synchronise torch's stream before calling.
"""
import ctypes
import os

import torch

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libref_kernels.so")
_LIB = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(_PATH)
    return _LIB


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _ok(rc, what):
    if rc != 0:
        raise RuntimeError(f"reference kernel {what} failed with cudaError {rc}")


def corr2d_fwd(in1_nhwc, in2_nhwc, md, sync=True):
    B, H, W, C = in1_nhwc.shape
    out = torch.zeros((B, (2 * md + 1) ** 2, H, W), dtype=torch.float32, device=in1_nhwc.device)
    torch.cuda.synchronize()
    _ok(lib().ref_corr2d_fwd(_p(in1_nhwc), _p(in2_nhwc), _p(out), B, C, H, W, md, int(sync)), "corr2d_fwd")
    return out


def corr2d_bwd(gout, in1_nhwc, in2_nhwc, md, sync=True):
    B, H, W, C = in1_nhwc.shape
    g1 = torch.empty((B, C, H, W), dtype=torch.float32, device=in1_nhwc.device)
    g2 = torch.empty_like(g1)
    torch.cuda.synchronize()
    _ok(lib().ref_corr2d_bwd(_p(gout), _p(in1_nhwc), _p(in2_nhwc), _p(g1), _p(g2), B, C, H, W, md, int(sync)), "corr2d_bwd")
    return g1, g2


def fps(xyz, n_samples, sync=True):
    B, N, _ = xyz.shape
    idx = torch.empty((B, n_samples), dtype=torch.int64, device=xyz.device)
    tmp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
    torch.cuda.synchronize()
    _ok(lib().ref_fps(_p(xyz), _p(tmp), _p(idx), B, N, n_samples, int(sync)), "fps")
    return idx


def knn(input_xyz, query_xyz, k, sync=True):
    B, M, D = input_xyz.shape
    Q = query_xyz.shape[1]
    idx = torch.zeros((B, Q, k), dtype=torch.int64, device=query_xyz.device)
    torch.cuda.synchronize()
    _ok(lib().ref_knn(_p(input_xyz), _p(query_xyz), _p(idx), B, M, Q, D, k, int(sync)), "knn")
    return idx
