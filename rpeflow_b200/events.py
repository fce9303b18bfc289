"""Drop-in replacements for the reference's event voxelisers.

``eventsToVoxel`` mirrors event_utils.py:109-110 of danqu130/RPEFlow (numpy in, numpy out; integer pixels,
temporal-bilinear weights, optional polarity split) and ``eventsToVoxelInter`` mirrors
``DSECTrain.eventsToVoxelInter`` (dsec.py:570-604; float pixels, tri-linear splat).  Both copy the events to the
GPU, run one atomic-scatter kernel (b200_event_voxel_int / b200_event_voxel_trilinear) and copy the grid back.
``*_device`` variants keep everything on the GPU for pipelines whose events are already resident.
"""
import numpy as np
import torch

from ._lib import check, lib

__all__ = ["eventsToVoxel", "eventsToVoxelInter", "events_to_voxel_device", "events_to_voxel_trilinear_device"]


def _dev(device):
    if not torch.cuda.is_available():
        raise RuntimeError("rpeflow_b200.events: a CUDA device is required — no CPU/torch fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def _grid(out, num_bins, height, width, event_polarity, device):
    shape = (num_bins * (2 if event_polarity else 1), height, width)
    if out is None:
        return torch.empty(shape, dtype=torch.float32, device=device)
    if tuple(out.shape) != shape or out.dtype != torch.float32 or not out.is_contiguous() or out.device != device:
        raise RuntimeError(f"out must be a contiguous CUDA fp32 tensor of shape {shape}")
    return out


def events_to_voxel_device(events, num_bins, height, width, event_polarity, check_range=True, out=None):
    """events: CUDA fp32 [n,4] (x,y,t,p), time-sorted -> CUDA fp32 [bins*(2|1),H,W] (written into `out` if given,
    e.g. one sample's slice of a batched grid)."""
    ev = events.contiguous().float()
    n = ev.shape[0]
    vox = _grid(out, num_bins, height, width, event_polarity, ev.device)
    status = torch.empty((1,), dtype=torch.int32, device=ev.device)
    with torch.cuda.device(ev.device):
        check(lib.b200_event_voxel_int(ev.data_ptr(), n, vox.data_ptr(), int(num_bins), int(height), int(width),
                                       int(bool(event_polarity)), status.data_ptr(),
                                       torch.cuda.current_stream(ev.device).cuda_stream), "b200_event_voxel_int")
    if check_range:
        bad = int(status.item())
        if bad:                                   # the reference's index_put_ raises IndexError here
            raise IndexError(f"{bad} events fall outside the {height}x{width} sensor")
    return vox


def eventsToVoxel(events, num_bins=5, height=None, width=None, event_polarity=False, temporal_bilinear=True,
                  device=None):
    """numpy [n,4] float (x,y,t,p) -> numpy [bins*(2|1),H,W] float32 (event_utils.py:109-128)."""
    if not temporal_bilinear:
        raise NotImplementedError("temporal_bilinear=False is not on the hot path (every caller uses the default)")
    events = np.asarray(events)
    if events.dtype != np.float32:
        # The reference normalises (t - t0) / (dt + 1e-6) in the INPUT dtype (event_utils.py:30-34); float64 stamps with a
        # large absolute value (microseconds since boot) would quantise if they were cast first.  Shift in the original
        # dtype, then cast: the kernel's own (t - t_first) becomes a no-op on the already shifted stamps.
        shifted = np.empty(events.shape, dtype=np.float32)
        shifted[:, 0], shifted[:, 1], shifted[:, 3] = events[:, 0], events[:, 1], events[:, 3]
        shifted[:, 2] = events[:, 2] - events[0, 2]
        events = shifted
    events = np.ascontiguousarray(events, dtype=np.float32)
    if height is None or width is None:           # event_utils.py:116-118
        width = int(events[:, 0].astype(np.int32).max()) + 1
        height = int(events[:, 1].astype(np.int32).max()) + 1
    ev = torch.from_numpy(events).to(_dev(device), non_blocking=True)
    return events_to_voxel_device(ev, num_bins, height, width, event_polarity).cpu().numpy()


def events_to_voxel_trilinear_device(x, y, t, p, num_bins, height, width, event_polarity, out=None):
    """x,y,p: CUDA fp32 [n]; t: CUDA int64 [n] (sorted) -> CUDA fp32 [bins*(2|1),H,W]."""
    x, y, p = (a.contiguous().float() for a in (x, y, p))
    t = t.contiguous().to(torch.int64)
    n = x.shape[0]
    vox = _grid(out, num_bins, height, width, event_polarity, x.device)
    scratch = torch.empty((8,), dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.b200_event_voxel_trilinear(x.data_ptr(), y.data_ptr(), t.data_ptr(), p.data_ptr(), n, vox.data_ptr(),
                                             int(num_bins), int(height), int(width), int(bool(event_polarity)),
                                             scratch.data_ptr(), torch.cuda.current_stream(x.device).cuda_stream),
              "b200_event_voxel_trilinear")
    return vox


def eventsToVoxelInter(events, num_bins, height, width, event_polarity=False, device=None):
    """dict of numpy x,y (float), t (int microseconds), p -> numpy [bins*(2|1),H,W] (dsec.py:570-604)."""
    dev = _dev(device)
    x = torch.from_numpy(np.ascontiguousarray(events['x'], dtype=np.float32)).to(dev, non_blocking=True)
    y = torch.from_numpy(np.ascontiguousarray(events['y'], dtype=np.float32)).to(dev, non_blocking=True)
    p = torch.from_numpy(np.ascontiguousarray(events['p'], dtype=np.float32)).to(dev, non_blocking=True)
    t = torch.from_numpy(np.ascontiguousarray(events['t']).astype(np.int64)).to(dev, non_blocking=True)
    return events_to_voxel_trilinear_device(x, y, t, p, num_bins, height, width, event_polarity).cpu().numpy()
