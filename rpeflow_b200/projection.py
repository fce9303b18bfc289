"""Drop-in replacements for the reference's gather / projection helpers (models/utils.py of danqu130/RPEFlow).

Same signatures and return shapes as ``batch_indexing_channel_first`` (:119-137), ``batch_indexing_channel_last``
(:101-116), ``grid_sample_wrapper`` (:288-294) and ``project_feat_with_nn_corr`` (:297-317); each is one or two
sm_100a kernels behind include/b200flow.h.  Forward only: the reference calls the last two under
``torch.no_grad()`` everywhere (RPEFlow_core.py:52,105,156; models/utils.py:297), and the gathers are used on
index/geometry tensors.  When autograd needs a gradient through a gather, keep the torch version.
"""
import torch

from ._lib import check, lib

__all__ = ["batch_indexing_channel_first", "batch_indexing_channel_last", "grid_sample_wrapper", "backwarp_2d", "convex_upsample",
           "project_feat_with_nn_corr", "knn_interpolation", "backwarp_3d"]


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _prep_data(data, name):
    if not data.is_cuda:
        raise RuntimeError(f"rpeflow_b200.{name}: CUDA tensors required — no CPU/torch fallback")
    if data.element_size() != 4:
        raise RuntimeError(f"rpeflow_b200.{name}: only 4-byte element types are supported (got {data.dtype})")
    return data.contiguous()


def _prep_idx(idx, device):
    return idx.to(device=device, dtype=torch.int64).contiguous()


# Deliberate deviation from the reference: the gather kernels wrap negative indices once (python / advanced-indexing
# semantics; torch.gather itself raises on them) and CLAMP out-of-range ones instead of raising, because raising needs a
# device-to-host sync per call.  With B200_CHECK_INDICES=1 every gather counts the indices it had to clamp and raises
# IndexError like torch.gather does (one sync per call: a debugging aid for corrupt KNN / FPS outputs, not for production).
import os as _os
CHECK_INDICES = _os.environ.get("B200_CHECK_INDICES", "0") not in ("", "0")


def _bad_counter(device):
    return torch.zeros((1,), dtype=torch.int32, device=device) if CHECK_INDICES else None


def _raise_if_bad(counter, what, n):
    if counter is not None:
        bad = int(counter.item())
        if bad:
            raise IndexError(f"rpeflow_b200.{what}: {bad} indices are out of range for a dimension of size {n}")


def batch_indexing_channel_first(batched_data, batched_indices):
    """[B,C,N], [B,I1..Im] -> [B,C,I1..Im] (bit-exact move of 4-byte elements)."""
    assert batched_data.shape[0] == batched_indices.shape[0]
    data = _prep_data(batched_data, "batch_indexing_channel_first")
    if data.dim() != 3:
        raise RuntimeError("batch_indexing_channel_first expects data of shape [B,C,N]")
    idx = _prep_idx(batched_indices, data.device)
    B, C, N = data.shape
    I = idx.numel() // max(B, 1)
    out = torch.empty((B, C, I), dtype=data.dtype, device=data.device)
    bad = _bad_counter(data.device)
    with torch.cuda.device(data.device):
        check(lib.b200_gather_cf(data.data_ptr(), idx.data_ptr(), out.data_ptr(), B, C, N, I,
                                 bad.data_ptr() if bad is not None else None, _stream(data)), "b200_gather_cf")
    _raise_if_bad(bad, "batch_indexing_channel_first", N)
    return out.view([B, C] + list(idx.shape[1:]))


def batch_indexing_channel_last(batched_data, batched_indices):
    """[B,N,C] (or [B,N]), [B,I1..Im] -> [B,I1..Im,C] (or [B,I1..Im])."""
    assert batched_data.shape[0] == batched_indices.shape[0]
    data = _prep_data(batched_data, "batch_indexing_channel_last")
    flat = data.dim() == 2
    if flat:
        data = data.unsqueeze(-1)
    if data.dim() != 3:
        raise RuntimeError("batch_indexing_channel_last expects data of shape [B,N,C] or [B,N]")
    idx = _prep_idx(batched_indices, data.device)
    B, N, C = data.shape
    I = idx.numel() // max(B, 1)
    out = torch.empty((B, I, C), dtype=data.dtype, device=data.device)
    bad = _bad_counter(data.device)
    with torch.cuda.device(data.device):
        check(lib.b200_gather_cl(data.data_ptr(), idx.data_ptr(), out.data_ptr(), B, C, N, I,
                                 bad.data_ptr() if bad is not None else None, _stream(data)), "b200_gather_cl")
    _raise_if_bad(bad, "batch_indexing_channel_last", N)
    shape = [B] + list(idx.shape[1:])
    return out.view(shape) if flat else out.view(shape + [C])


class _SampleMemo:
    """Samples that project_feat_with_nn_corr(xy, feat_2d, ...) has already taken: bilinear(feat_2d, xy) as [B,C,N].

    The model samples the same map at the same points twice in four of its five fuser pairs per level — the 3D->2D fuser
    calls project_feat_with_nn_corr(xy, feat_2d, ...) and the 2D->3D fuser then calls grid_sample_wrapper(feat_2d, xy)
    (RPEFlow_core.py:31+53, :80+107, :134+157).  The first call's sampler kernel writes that second result as a by-product
    (b200_project_nn_corr_sampled); it is parked here and handed out ONCE to the next grid_sample_wrapper call with the
    same (map, points).  An entry holds references to both key tensors — their storage cannot be recycled while it is
    parked — and records their version counters, so an in-place write to either one invalidates it.  A handful of
    entries (a level has four pairs in flight) bounds what is kept alive."""

    def __init__(self, capacity=6):
        self.capacity = capacity
        self.entries = []
        self.hits = self.misses = 0

    @staticmethod
    def _key(t):
        return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t.dtype, t._version)

    def put(self, feat_2d, xy, sampled):
        self.entries.append((feat_2d, self._key(feat_2d), xy, self._key(xy), sampled))
        del self.entries[:-self.capacity]

    def take(self, feat_2d, xy):
        kf, kx = self._key(feat_2d), self._key(xy)
        for i in range(len(self.entries) - 1, -1, -1):
            f, f_key, x, x_key, sampled = self.entries[i]
            if f_key == kf and x_key == kx and self._key(f) == kf and self._key(x) == kx:
                del self.entries[i]
                self.hits += 1
                return sampled
        self.misses += 1
        return None

    def clear(self):
        self.entries.clear()


SAMPLE_MEMO = _SampleMemo()


def grid_sample_wrapper(feat_2d, xy):
    """feat_2d [B,C,H,W], xy [B,2,N] (pixels) -> [B,C,N]; bilinear, align_corners=True, zero padding.
    If project_feat_with_nn_corr has just sampled this very map at these very points, its result is returned
    (see _SampleMemo) and no kernel runs."""
    if not (feat_2d.is_cuda and xy.is_cuda):
        raise RuntimeError("rpeflow_b200.grid_sample_wrapper: CUDA tensors required — no CPU/torch fallback")
    parked = SAMPLE_MEMO.take(feat_2d, xy)
    if parked is not None:
        return parked
    feat = feat_2d.contiguous().float()
    pts = xy.contiguous().float()
    B, C, H, W = feat.shape
    N = pts.shape[2]
    out = torch.empty((B, C, N), dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        check(lib.b200_grid_sample_pts(feat.data_ptr(), pts.data_ptr(), out.data_ptr(), B, C, H, W, N,
                                       _stream(feat)), "b200_grid_sample_pts")
    return out


def backwarp_2d(x, flow12, padding_mode="border"):
    """x [B,C,H,W], flow12 [B,2,H,W] -> x sampled at (pixel + flow), bilinear, align_corners=True (models/utils.py:186-198).
    The model only ever passes padding_mode='border' (RPEFlow_core.py:351); other modes are not built."""
    if not (x.is_cuda and flow12.is_cuda):
        raise RuntimeError("rpeflow_b200.backwarp_2d: CUDA tensors required — no CPU/torch fallback")
    if padding_mode != "border":
        raise RuntimeError("rpeflow_b200.backwarp_2d: only padding_mode='border' is built (the mode RPEFlow uses)")
    assert x.shape[-2:] == flow12.shape[-2:]                       # models/utils.py:193
    if torch.is_grad_enabled() and (x.requires_grad or flow12.requires_grad):
        raise RuntimeError("rpeflow_b200.backwarp_2d: forward only; run training through the reference's torch ops")
    feat = x.contiguous().float()
    flow = flow12.contiguous().float()
    B, C, H, W = feat.shape
    out = torch.empty_like(feat)
    with torch.cuda.device(feat.device):
        check(lib.b200_backwarp2d(feat.data_ptr(), flow.data_ptr(), out.data_ptr(), B, C, H, W, _stream(feat)),
              "b200_backwarp2d")
    return out


def convex_upsample(flow, mask, scale_factor=8):
    """flow [B,2,H,W], mask [B,9*s*s,H,W] -> [B,2,s*H,s*W]: convex combination of the 3x3 neighbourhood with softmax
    weights (models/utils.py:201-214).  Forward only (the model trains through the reference's torch ops)."""
    if not (flow.is_cuda and mask.is_cuda):
        raise RuntimeError("rpeflow_b200.convex_upsample: CUDA tensors required — no CPU/torch fallback")
    if torch.is_grad_enabled() and (flow.requires_grad or mask.requires_grad):
        raise RuntimeError("rpeflow_b200.convex_upsample: forward only; run training through the reference's torch ops")
    s = int(scale_factor)
    B, C, H, W = flow.shape
    if C != 2 or s not in (2, 4, 8) or tuple(mask.shape) != (B, 9 * s * s, H, W):
        raise RuntimeError("rpeflow_b200.convex_upsample: needs flow [B,2,H,W], mask [B,9*s*s,H,W], s in {2,4,8}")
    f, m = flow.contiguous().float(), mask.contiguous().float()
    out = torch.empty((B, 2, H * s, W * s), dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        check(lib.b200_convex_upsample(f.data_ptr(), m.data_ptr(), out.data_ptr(), B, H, W, s, _stream(f)),
              "b200_convex_upsample")
    return out


@torch.no_grad()
def project_feat_with_nn_corr(xy, feat_2d, feat_3d, nn_indices=None, keep_samples=True):
    """xy [B,2,N], feat_2d [B,C2,H,W], feat_3d [B,C3,N], nn_indices [B,H*W] -> [B,C3+3,H,W].
    keep_samples: also keep bilinear(feat_2d, xy) as [B,C2,N] for the grid_sample_wrapper(feat_2d, xy) call that
    follows in the model (one extra 4*C2*N-byte write instead of a second pass over the map)."""
    if not (xy.is_cuda and feat_2d.is_cuda and feat_3d.is_cuda):
        raise RuntimeError("rpeflow_b200.project_feat_with_nn_corr: CUDA tensors required — no CPU/torch fallback")
    f2 = feat_2d.contiguous().float()
    f3 = feat_3d.contiguous().float()
    pts = xy.contiguous().float()
    B, C2, H, W = f2.shape
    C3, N = f3.shape[1], f3.shape[2]
    if nn_indices is None:                              # models/utils.py:304-305: nearest projected point per pixel
        from .ops import k_nearest_neighbor
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=f2.device),
                                torch.arange(W, dtype=torch.float32, device=f2.device), indexing="ij")
        grid = torch.stack([xs, ys], 0).reshape(1, 2, H * W).expand(B, 2, H * W)
        nn_indices = k_nearest_neighbor(pts, grid, k=1)[..., 0]
    else:
        assert tuple(nn_indices.shape) == (B, H * W)
    nn = _prep_idx(nn_indices, f2.device)
    out = torch.empty((B, C3 + 3, H, W), dtype=torch.float32, device=f2.device)
    scratch = torch.empty((max(1, lib.b200_project_nn_corr_scratch_floats(B, C2, C3, N)),), dtype=torch.float32, device=f2.device)
    sampled = torch.empty((B, C2, N), dtype=torch.float32, device=f2.device) if keep_samples else None
    with torch.cuda.device(f2.device):
        check(lib.b200_project_nn_corr_sampled(pts.data_ptr(), f2.data_ptr(), f3.data_ptr(), nn.data_ptr(), out.data_ptr(),
                                               scratch.data_ptr(), sampled.data_ptr() if keep_samples else None,
                                               B, C2, C3, H, W, N, _stream(f2)), "b200_project_nn_corr")
    if keep_samples:
        SAMPLE_MEMO.put(feat_2d, xy, sampled)
    return out


def knn_interpolation(input_xyz, input_features, query_xyz, k=3, knn_indices=None):
    """input_xyz [B,3,M], input_features [B,C,M], query_xyz [B,3,Q] -> [B,C,Q] (models/utils.py:140-156): the k-nearest
    search (cell-grid KNN) followed by ONE kernel for gathers + inverse-distance weights + weighted sum.  Forward only."""
    if not (input_xyz.is_cuda and input_features.is_cuda and query_xyz.is_cuda):
        raise RuntimeError("rpeflow_b200.knn_interpolation: CUDA tensors required — no CPU/torch fallback")
    from .ops import k_nearest_neighbor
    xin, feat, xq = (t.contiguous().float() for t in (input_xyz, input_features, query_xyz))
    if knn_indices is None:
        knn_indices = k_nearest_neighbor(xin, xq, k)                          # [B,Q,k]
    idx = _prep_idx(knn_indices, xin.device)
    B, C, M = feat.shape
    Q = xq.shape[2]
    assert tuple(idx.shape) == (B, Q, k) and xin.shape == (B, 3, M)
    out = torch.empty((B, C, Q), dtype=torch.float32, device=xin.device)
    with torch.cuda.device(xin.device):
        check(lib.b200_knn_interpolate(xin.data_ptr(), feat.data_ptr(), xq.data_ptr(), idx.data_ptr(), out.data_ptr(),
                                       B, C, M, Q, int(k), _stream(xin)), "b200_knn_interpolate")
    return out


def backwarp_3d(xyz1, xyz2, flow12, k=3):
    """models/utils.py:159-169: xyz2 + interpolation of -flow12 from the warped cloud xyz1 + flow12."""
    return xyz2 + knn_interpolation(xyz1 + flow12, -flow12, xyz2, k)
