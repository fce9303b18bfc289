"""torch-CPU restatement of the reference's own fallback path (see oracle/__init__.py — test infrastructure).

Each function re-states, with the same torch operators in the same order, what the reference executes when
its CUDA extensions are absent, so that on CPU the numbers agree with the reference to the last bit wherever
ATen is deterministic.  tests/test_oracle_golden.py pins every function against fixtures produced by the
unmodified reference.  ``bench.py --impl reference`` times exactly these functions (kind = "port").

Reference citations: file:line in danqu130/RPEFlow.
"""
import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------ a1  models/csrc/wrapper.py:55-65
def correlation2d(feat1, feat2, md):
    """feat1, feat2 [B,C,H,W] -> [B,(2md+1)^2,H,W]; channel = row-shift * (2md+1) + column-shift."""
    h, w = feat1.shape[-2:]
    padded = F.pad(feat2, (md, md, md, md))
    span = 2 * md + 1
    planes = [(feat1 * padded[:, :, r:r + h, c:c + w]).mean(dim=1, keepdim=True)
              for r in range(span) for c in range(span)]
    return torch.cat(planes, dim=1)


# ------------------------------------------------------------------ a3  models/csrc/wrapper.py:83-96
def furthest_point_sampling(xyz, n_samples):
    """xyz [B,N,3] -> [B,n_samples] int64; start at index 0, distances start at 1e10, first max wins."""
    nb, npts, _ = xyz.shape
    assert xyz.shape[2] == 3 and npts > n_samples                      # wrapper.py:98
    picked = torch.zeros(nb, n_samples, dtype=torch.int64)
    nearest = torch.full((nb, npts), 1e10, dtype=xyz.dtype)
    rows = torch.arange(nb)
    cur = torch.zeros(nb, dtype=torch.int64)
    for s in range(n_samples):
        picked[:, s] = cur
        centre = xyz[rows, cur].unsqueeze(1)
        d = ((xyz - centre) ** 2).sum(-1)
        closer = d < nearest
        nearest[closer] = d[closer]
        cur = nearest.max(dim=-1).indices
    return picked


# ------------------------------------------------------------------ a4  models/csrc/wrapper.py:40-52,106-127
def squared_distance(a, b):
    """a [B,Na,D], b [B,Nb,D] -> [B,Na,Nb] via -2ab^T + |a|^2 + |b|^2 (the reference's expanded form)."""
    d = -2 * torch.matmul(a, b.transpose(1, 2))
    d += (a ** 2).sum(-1).unsqueeze(2)
    d += (b ** 2).sum(-1).unsqueeze(1)
    return d


def k_nearest_neighbor(input_xyz, query_xyz, k):
    """Accepts [B,D,N] (D<=3) or [B,N,D]; returns [B,Q,k] int64 (topk of the expanded distance)."""
    if input_xyz.shape[1] <= 3:
        input_xyz = input_xyz.transpose(1, 2).contiguous()
        query_xyz = query_xyz.transpose(1, 2).contiguous()
    return squared_distance(query_xyz, input_xyz).topk(k, dim=2, largest=False).indices.long()


# ------------------------------------------------------------------ a6  models/utils.py:101-137
def batch_indexing_channel_first(data, idx):
    """data [B,C,N], idx [B,...] -> [B,C,...]."""
    nb, nc = data.shape[:2]
    tail = list(idx.shape[1:])
    flat = idx.reshape(nb, 1, -1).expand(nb, nc, -1).to(torch.int64)
    return torch.gather(data, 2, flat).view([nb, nc] + tail)


def batch_indexing_channel_last(data, idx):
    """data [B,N,C], idx [B,...] -> [B,...,C]."""
    nb = data.shape[0]
    b = torch.arange(nb).view([nb] + [1] * (idx.dim() - 1)).expand_as(idx)
    return data[b, idx.to(torch.int64)]


# ------------------------------------------------------------------ a7  models/utils.py:288-294
def grid_sample_wrapper(feat_2d, xy):
    """feat_2d [B,C,H,W], xy [B,2,N] in pixels -> [B,C,N] (bilinear, align_corners, zero padding)."""
    h, w = feat_2d.shape[2:]
    gx = 2.0 * xy[:, 0] / (w - 1) - 1.0
    gy = 2.0 * xy[:, 1] / (h - 1) - 1.0
    grid = torch.stack([gx, gy], dim=-1).unsqueeze(2)                  # [B,N,1,2]
    return F.grid_sample(feat_2d, grid, mode='bilinear', align_corners=True)[..., 0]


# ------------------------------------------------------------------ a8  models/utils.py:172-183,297-317
def pixel_grid(nb, h, w):
    """[B,2,H*W] float pixel coordinates, x in channel 0."""
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32),
                            indexing='ij')
    return torch.stack([xs, ys], 0).reshape(1, 2, h * w).expand(nb, 2, h * w)


def project_feat_with_nn_corr(xy, feat_2d, feat_3d, nn_indices):
    nb, _, h, w = feat_2d.shape
    grid = pixel_grid(nb, h, w)
    at_pts = grid_sample_wrapper(feat_2d, xy)
    nn_feat2d = batch_indexing_channel_first(at_pts, nn_indices)
    nn_feat3d = batch_indexing_channel_first(feat_3d, nn_indices)
    nn_offset = batch_indexing_channel_first(xy, nn_indices) - grid
    nn_corr = (nn_feat2d * feat_2d.reshape(nb, -1, h * w)).mean(dim=1, keepdim=True)
    return torch.cat([nn_offset, nn_corr, nn_feat3d], dim=1).reshape(nb, -1, h, w)


# ------------------------------------------------------------------ a5  models/pwc3d_core.py:69-117
def _mlp(x, layers, act):
    for wgt, bias in layers:
        x = act(F.conv2d(x, wgt[:, :, None, None], bias))
    return x


def correlation3d(xyz1, feat1, xyz2, feat2, wts, k=16, knn11=None, knn12=None):
    """wts: dict with the names of include/b200flow.h::b200_corr3d_weights (torch tensors)."""
    lrelu = lambda t: F.leaky_relu(t, 0.1)
    nb, cin, n1 = feat1.shape
    if knn12 is None:
        knn12 = k_nearest_neighbor(xyz2, xyz1, k)
    nbr_xyz = batch_indexing_channel_first(xyz2, knn12) - xyz1.unsqueeze(-1)
    nbr_feat = batch_indexing_channel_first(feat2, knn12)
    stacked = torch.cat([feat1.unsqueeze(-1).expand(nb, cin, n1, knn12.shape[2]), nbr_feat, nbr_xyz], dim=1)
    p2p = _mlp(stacked, [(wts['W1'], wts['b1']), (wts['W2'], wts['b2'])], lrelu)
    w2 = _mlp(nbr_xyz, [(wts['n2_Wa'], wts['n2_ba']), (wts['n2_Wb'], wts['n2_bb']),
                        (wts['n2_Wc'], wts['n2_bc'])], F.relu)
    p2n = (w2 * p2p).sum(dim=3)
    if knn11 is None:
        knn11 = k_nearest_neighbor(xyz1, xyz1, k)
    self_xyz = batch_indexing_channel_first(xyz1, knn11) - xyz1.unsqueeze(-1)
    w1 = _mlp(self_xyz, [(wts['n1_Wa'], wts['n1_ba']), (wts['n1_Wb'], wts['n1_bb']),
                         (wts['n1_Wc'], wts['n1_bc'])], F.relu)
    return (w1 * batch_indexing_channel_first(p2n, knn11)).sum(dim=3)


# ------------------------------------------------------------------ a9  event_utils.py:23-39,109-128,211-303
def events_to_voxel(events, num_bins, height, width, event_polarity, _accumulate=None):
    """events: numpy [n,4] float32 (x,y,t,p) -> numpy [bins*(2|1),H,W]."""
    xs = torch.from_numpy(events[:, 0].astype(np.int32)).long()
    ys = torch.from_numpy(events[:, 1].astype(np.int32)).long()
    ps = torch.from_numpy(events[:, 3].astype(np.int32))
    t = events[:, 2]
    t = torch.from_numpy((t - t[0]) / ((t[-1] - t[0]) + 1e-6))

    def one_grid(weights):
        tn = (t - t[0]) / (t[-1] - t[0]) * (num_bins - 1)
        out = []
        for b in range(num_bins):
            wb = weights * torch.clamp(1.0 - (tn - b).abs(), min=0)
            img = torch.zeros(height, width)
            img.index_put_((ys, xs), wb.float(), accumulate=True)
            out.append(img)
        return torch.stack(out)

    if not event_polarity:
        return one_grid(ps).numpy()
    one, zero = torch.ones(1), torch.zeros(1)
    return torch.cat([one_grid(torch.where(ps > 0, one, zero)),
                      one_grid(torch.where(ps <= 0, one, zero))], 0).numpy()


# ------------------------------------------------------------------ a10 dsec.py:536-604
def events_to_voxel_trilinear(x, y, t, p, num_bins, height, width, event_polarity):
    ts = (t - t[0]).astype('float32')
    ts = torch.from_numpy(ts / ts[-1])
    xs, ys, ps = (torch.from_numpy(np.asarray(a).astype('float32')) for a in (x, y, p))

    def splat(xs, ys, ts, value):
        grid = torch.zeros(num_bins, height, width)
        if xs.numel() == 0:
            return grid
        tn = (num_bins - 1) * (ts - ts[0]) / (ts[-1] - ts[0])
        x0, y0, t0 = xs.int(), ys.int(), tn.int()
        for xl in (x0, x0 + 1):
            for yl in (y0, y0 + 1):
                for tl in (t0, t0 + 1):
                    ok = (xl < width) & (xl >= 0) & (yl < height) & (yl >= 0) & (tl >= 0) & (tl < num_bins)
                    wgt = value * (1 - (xl - xs).abs()) * (1 - (yl - ys).abs()) * (1 - (tl - tn).abs())
                    lin = height * width * tl.long() + width * yl.long() + xl.long()
                    grid.put_(lin[ok], wgt[ok], accumulate=True)
        return grid

    if not event_polarity:
        return splat(xs, ys, ts, 2 * ps - 1).numpy()
    pos, neg = ps > 0, ps <= 0
    return torch.cat([splat(xs[pos], ys[pos], ts[pos], 2 * ps[pos] - 1),
                      splat(xs[neg], ys[neg], ts[neg], 1)], 0).numpy()


def knn_interpolation(input_xyz, input_features, query_xyz, k=3, knn_indices=None):
    """models/utils.py:140-156: inverse-distance interpolation over the k nearest inputs (channel-first tensors)."""
    if knn_indices is None:
        knn_indices = k_nearest_neighbor(input_xyz, query_xyz, k)
    near = batch_indexing_channel_first(input_xyz, knn_indices)                  # [B,3,Q,k]
    dist = torch.linalg.norm(near - query_xyz[..., None], dim=1).clamp(1e-8)      # [B,Q,k]
    w = 1.0 / dist
    w = w / torch.sum(w, dim=-1, keepdim=True)
    feats = batch_indexing_channel_first(input_features, knn_indices)             # [B,C,Q,k]
    return torch.sum(feats * w[:, None, :, :], dim=-1)


def backwarp_3d(xyz1, xyz2, flow12, k=3):
    """models/utils.py:159-169."""
    return xyz2 + knn_interpolation(xyz1 + flow12, -flow12, xyz2, k)


# ------------------------------------------------------------------ f3  models/utils.py:186-198, RPEFlow_core.py:351,362
def backwarp_2d(x, flow12, padding_mode='border'):
    """x [B,C,H,W] sampled at pixel + flow12 [B,2,H,W]; grid normalised exactly as norm_grid does."""
    nb, _, h, w = x.shape
    xs = torch.arange(0, w, dtype=torch.float32)[None, None, :].expand(nb, h, w)
    ys = torch.arange(0, h, dtype=torch.float32)[None, :, None].expand(nb, h, w)
    g = torch.stack([xs, ys], 1) + flow12
    gn = torch.zeros_like(g)
    gn[:, 0] = 2.0 * g[:, 0] / (w - 1) - 1.0
    gn[:, 1] = 2.0 * g[:, 1] / (h - 1) - 1.0
    return F.grid_sample(x, gn.permute(0, 2, 3, 1), padding_mode=padding_mode, align_corners=True)


def warp_correlate(feat1, feat2, flow, md=4, negative_slope=0.1):
    """leaky_relu(correlation2d(feat1, backwarp_2d(feat2, flow, 'border'), md), 0.1) — RPEFlow_core.py:351 + :362."""
    warped = feat2 if flow is None else backwarp_2d(feat2, flow, 'border')
    return F.leaky_relu(correlation2d(feat1, warped, md), negative_slope)


# ------------------------------------------------------------------ f4  models/utils.py:201-214
def convex_upsample(flow, mask, scale_factor=8):
    nb, _, h, w = flow.shape
    mask = torch.softmax(mask.view(nb, 1, 9, scale_factor, scale_factor, h, w), dim=2)
    up = F.unfold(flow * scale_factor, [3, 3], padding=1).view(nb, 2, 9, 1, 1, h, w)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(nb, 2, h * scale_factor, w * scale_factor)


def pointconv(xyz, features, sampled_xyz, wts, k=16, knn=None):
    """models/pointconv.py:33-61 (and :90-122 with sampled_xyz = xyz).  wts: Wa [8,3], ba, Wb [16,8], bb, L, bias."""
    lrelu = lambda t: F.leaky_relu(t, 0.1)
    nb, ns = xyz.shape[0], sampled_xyz.shape[2]
    if knn is None:
        knn = k_nearest_neighbor(xyz, sampled_xyz, k)
    rel = batch_indexing_channel_first(xyz, knn) - sampled_xyz[:, :, :, None]                    # [B,3,S,k]
    weights = _mlp(rel, [(wts['Wa'], wts['ba']), (wts['Wb'], wts['bb'])], lrelu).transpose(1, 2)   # [B,S,16,k]
    cl = torch.cat([xyz, features], dim=1).transpose(1, 2)                                       # [B,N,C+3]
    nbr = batch_indexing_channel_last(cl, knn)                                                   # [B,S,k,C+3]
    mixed = torch.matmul(weights, nbr).reshape(nb, ns, -1)                                       # [B,S,16*(C+3)]
    return lrelu(F.linear(mixed, wts['L'], wts['bias']).transpose(1, 2))
