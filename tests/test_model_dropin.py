"""The drop-in contract at the level it is stated: the UNMODIFIED reference model (models/RPEFlow.py:36-99 ->
RPEFlow_core.py:302-432) runs with this library underneath and produces the reference's results.

CPU (no GPU here): install() / uninstall() round trip on the real model — with CPU tensors every re-bound op hands the
call back to the reference, so the flows must be bit-identical to the untouched reference and uninstall() must put
every name back.

GPU (-m gpu): three arms of the same model instance, same weights, same inputs, at 960x540 / 8192 points:
  A  the reference's own torch path on the GPU (no extension: wrapper.py's fallbacks);
  C  the reference with ITS OWN CUDA extensions (oracle/_ref/ref_ext, built by its setup.py);
  B  install(): every hot op on this library's sm_100a kernels.
Checked: every hot op of arm B took the kernel route (install.stats()); the FPS index list is identical in all arms;
each of the 43 KNN results of arm B equals the CPU oracle bit for bit on the very inputs the model passed, and differs
from arm A / arm C on identical inputs only within the near-tie bound the reference's own KNN test tolerates
(k_nearest_neighbor_test.cpp:61-63 prints the mismatch count); flow_2d / flow_3d of arm B match arm A within the stated
tolerance.  The reference tree is the git-ignored copy staged by `make -C oracle reftree` (it travels to the GPU box).
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from refmodel_util import (ROOT, record_calls, reference_extensions_available, reference_extensions_bound,
                           reference_root)

REF = reference_root()
needs_ref = pytest.mark.skipif(REF is None, reason="no reference tree (run `make -C oracle reftree` where /root/reference exists)")


@needs_ref
def test_install_uninstall_round_trip_on_the_real_model_cpu():
    code = f"""
import sys, torch
sys.path.insert(0, {ROOT!r})
from rpeflow_b200 import refhost
import rpeflow_b200.install as inst
model = refhost.build_rpeflow({REF!r}, device='cpu', install=False)
inp = refhost.synthetic_model_inputs(1, height=128, width=192, n_points=8192)
a = refhost.forward(model, inp)
import models.utils as mu, models.csrc.wrapper as w, models.RPEFlow_core as core, models.pwc3d_core as p3
before = (core.grid_sample_wrapper, core.k_nearest_neighbor, p3.Correlation3D.forward, core.CorrFeatureFuser3D.forward)
inst.install()                                   # late install: wrapper.py already ran its import block
assert w._k_nearest_neighbor_cuda is not None and core.grid_sample_wrapper is not before[0]
b = refhost.forward(model, inp)
st = inst.stats()
assert st['k_nearest_neighbor'] == {{'b200': 0, 'reference': 43}}, st
assert st['furthest_point_sampling']['reference'] == 1 and st['correlation2d']['reference'] == 5
assert st['Correlation3D.forward']['reference'] == 5 and st['project_feat_with_nn_corr']['reference'] == 20
assert st['PointConvDownSampling.forward']['reference'] == 10 and st['PointConvNoSampling.forward']['reference'] == 10
assert torch.equal(a['flow_2d'], b['flow_2d']) and torch.equal(a['flow_3d'], b['flow_3d'])
inst.uninstall()
after = (core.grid_sample_wrapper, core.k_nearest_neighbor, p3.Correlation3D.forward, core.CorrFeatureFuser3D.forward)
assert all(x is y for x, y in zip(before, after)) and w._k_nearest_neighbor_cuda is None
assert not hasattr(p3.Correlation3D, '_b200_reference_forward')
# per-call dispatch (ADVICE r1): a CPU call on a re-bound op never raises, grads keep the torch graph
inst.install()
xy = torch.rand(1, 2, 7, requires_grad=True)
f = torch.rand(1, 3, 6, 6)
assert core.grid_sample_wrapper(f, xy).shape == (1, 3, 7)
assert mu.project_feat_with_nn_corr(torch.rand(1, 2, 7), f, torch.rand(1, 4, 7)).shape == (1, 7, 6, 6)
assert core.correlation2d(f, f, 5).shape == (1, 121, 6, 6)           # CPU tensors: the reference's own loop
print('ROUNDTRIP-OK')
"""
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert "ROUNDTRIP-OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def _knn_layout(x):
    """[B,D,N] (channel-first, what the model passes) or [B,N,D] -> numpy [B,N,D]."""
    x = x.detach().float().cpu()
    if x.shape[1] <= 3:
        x = x.transpose(1, 2)
    return np.ascontiguousarray(x.numpy())


def _near_tie_report(inp, qry, got, ref, k):
    """Rows of `got` and `ref` ([Q,k] index lists for one cloud) that differ: every differing row must be explained by
    near-ties — the fp64 distances of the two lists, position by position, differ by at most 4 ulp_fp32 of the squared
    coordinate magnitude (SURVEY §8a parity rule (ii))."""
    rows = np.nonzero((got != ref).any(axis=1))[0]
    if rows.size == 0:
        return 0, 0.0
    q = qry[rows].astype(np.float64)[:, None, :]
    dg = ((inp[got[rows]].astype(np.float64) - q) ** 2).sum(-1)
    dr = ((inp[ref[rows]].astype(np.float64) - q) ** 2).sum(-1)
    scale = np.maximum((inp.astype(np.float64) ** 2).sum(-1).max(), (qry.astype(np.float64) ** 2).sum(-1).max())
    bound = 4 * np.spacing(np.float32(scale)).astype(np.float64) + 1e-30
    worst = float(np.abs(dg - dr).max() / bound)
    return int(rows.size), worst


@needs_ref
@pytest.mark.gpu
def test_reference_model_forward_through_the_drop_in():
    from oracle import spec
    from rpeflow_b200 import refhost
    import rpeflow_b200.install as inst

    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = False            # fp32 everywhere outside the hot path, in every arm
    torch.backends.cuda.matmul.allow_tf32 = False
    B = 2
    model = refhost.build_rpeflow(REF, device=dev, install=False, seed=0)
    host = refhost.synthetic_model_inputs(B, height=540, width=960, n_points=8192, seed=0)
    inputs = {k: v.to(dev) for k, v in host.items()}
    names = ("k_nearest_neighbor", "furthest_point_sampling")

    # ---- arm A: the reference's torch fallbacks on the GPU
    log_a = []
    with record_calls(names, log_a):
        out_a = refhost.forward(model, inputs)
    out_a2 = refhost.forward(model, inputs)            # run-to-run noise floor of the arm itself
    noise2d = (out_a["flow_2d"] - out_a2["flow_2d"]).abs().max().item()
    noise3d = (out_a["flow_3d"] - out_a2["flow_3d"]).abs().max().item()

    # ---- arm C: the reference with its own CUDA extensions
    log_c, out_c = [], None
    if reference_extensions_available():
        with reference_extensions_bound(), record_calls(names, log_c):
            out_c = refhost.forward(model, inputs)

    # ---- arm B: this library underneath
    log_b = []
    inst.install()
    inst.stats(reset=True)
    inst.set_observer(lambda name, args, kwargs, out: log_b.append((name, args, kwargs, out)) if name in names else None)
    try:
        out_b = refhost.forward(model, inputs)
        torch.cuda.synchronize()
        st = inst.stats()
    finally:
        inst.set_observer(None)
        inst.uninstall()

    # every hot op ran a kernel of this library, none fell back
    expect = {"furthest_point_sampling": 1, "correlation2d": 5, "Correlation3D.forward": 5, "project_feat_with_nn_corr": 20,
              "PointConvDownSampling.forward": 10, "PointConvNoSampling.forward": 10, "knn_interpolation": 9,
              "backwarp_3d": 4, "backwarp_2d": 4, "convex_upsample": 1}
    for op, n in expect.items():
        assert st[op] == {"b200": n, "reference": 0}, (op, st[op])
    assert st["grid_sample_wrapper"]["reference"] == 0 and st["batch_indexing_channel_first"]["reference"] == 0
    assert st["k_nearest_neighbor"]["reference"] == 0

    def split(log):
        fps = [(a, o) for n, a, k, o in log if n == "furthest_point_sampling"]
        knn = [(a, k, o) for n, a, k, o in log if n == "k_nearest_neighbor"]
        return fps, knn
    fps_a, knn_a = split(log_a)
    fps_b, knn_b = split(log_b)
    assert len(fps_a) == len(fps_b) == 1 and len(knn_a) == len(knn_b) == 43

    # FPS: identical index lists in every arm, and equal to the oracle
    idx_b = fps_b[0][1]
    assert torch.equal(fps_a[0][1], idx_b), "FPS indices differ from the reference torch path"
    np.testing.assert_array_equal(idx_b.cpu().numpy(), spec.fps(fps_b[0][0][0].float().cpu().numpy(), idx_b.shape[1]))
    if out_c is not None:
        fps_c, knn_c = split(log_c)
        assert torch.equal(fps_c[0][1], idx_b), "FPS indices differ from the reference CUDA kernel"

    def knn_args(args, kwargs):
        x = kwargs.get("input_xyz", args[0] if len(args) > 0 else None)
        q = kwargs.get("query_xyz", args[1] if len(args) > 1 else None)
        k = kwargs.get("k", args[2] if len(args) > 2 else None)
        return x, q, int(k)

    # KNN, call site by call site
    report = {"exact_vs_oracle": 0, "same_inputs_as_A": 0, "rows_differing_vs_A": 0, "rows_total": 0, "worst_tie_ratio_vs_A": 0.0,
              "rows_differing_vs_C": 0, "worst_tie_ratio_vs_C": 0.0}
    for i, (args, kwargs, got) in enumerate(knn_b):
        x, q, k = knn_args(args, kwargs)
        xn, qn, gn = _knn_layout(x), _knn_layout(q), got.cpu().numpy()
        np.testing.assert_array_equal(gn, spec.knn(xn, qn, k), err_msg=f"KNN call {i} ({tuple(x.shape)} k={k}) differs from the oracle")
        report["exact_vs_oracle"] += 1
        for tag, other in (("A", knn_a), ("C", knn_c if out_c is not None else None)):
            if other is None:
                continue
            xo, qo, ko = knn_args(other[i][0], other[i][1])
            assert ko == k and xo.shape == x.shape and qo.shape == q.shape, f"KNN call {i}: call sites out of step"
            if not (torch.equal(xo, x) and torch.equal(qo, q)):
                continue                                   # inputs already carry float differences of earlier ops
            on = other[i][2].cpu().numpy()
            if tag == "A":
                report["same_inputs_as_A"] += 1
                report["rows_total"] += gn.shape[0] * gn.shape[1]
            for b in range(gn.shape[0]):
                rows, worst = _near_tie_report(xn[b], qn[b], gn[b], on[b], k)
                report[f"rows_differing_vs_{tag}"] += rows
                report[f"worst_tie_ratio_vs_{tag}"] = max(report[f"worst_tie_ratio_vs_{tag}"], worst)
    assert report["same_inputs_as_A"] >= 20, report          # pyramid, pixel-grid and self searches see identical inputs
    assert report["rows_differing_vs_A"] <= 0.01 * report["rows_total"], report
    assert report["worst_tie_ratio_vs_A"] <= 1.0 and report["worst_tie_ratio_vs_C"] <= 1.0, report

    # flows.  The reference has TWO KNN arithmetics of its own — the torch fallback's expanded -2qx+|q|^2+|x|^2 (arm A) and
    # its CUDA kernel's direct difference (arm C) — and a random-init model amplifies the handful of near-tie neighbour
    # swaps between them (profiles/r2_model_ablation.md: swapping ONLY k_nearest_neighbor moves flow_2d by 2e-2 of its
    # scale, identically for the reference's own extension and for this library).  Stated tolerances:
    #   vs arm C (the reference's GPU path, same direct-distance rule):  2e-3 of the flow scale
    #   vs arm A: no farther from the torch fallback than the reference's own CUDA path is (x2 + the bar above)
    def flow_err(x, y):
        return {k: (x[k] - y[k]).abs().max().item() for k in ("flow_2d", "flow_3d")}
    scale = {k: max(1.0, out_a[k].abs().max().item()) for k in ("flow_2d", "flow_3d")}
    report["scale"] = scale
    report["B_vs_A"] = flow_err(out_b, out_a)
    for key in ("flow_2d", "flow_3d"):
        assert out_a[key].shape == out_b[key].shape and bool(torch.isfinite(out_b[key]).all())
    if out_c is not None:
        report["C_vs_A"] = flow_err(out_c, out_a)
        report["B_vs_C"] = flow_err(out_b, out_c)
        for key in ("flow_2d", "flow_3d"):
            assert report["B_vs_C"][key] <= 2e-3 * scale[key], (key, report)
            assert report["B_vs_A"][key] <= 2.0 * report["C_vs_A"][key] + 2e-3 * scale[key], (key, report)
    else:
        for key in ("flow_2d", "flow_3d"):
            assert report["B_vs_A"][key] <= 5e-2 * scale[key], (key, report)
    report["noise_floor_A"] = {"flow_2d": noise2d, "flow_3d": noise3d}
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    import json
    with open(os.path.join(out_dir, "model_dropin_report.json"), "w") as f:
        json.dump({"stats": st, "report": report}, f, indent=1)
    print("model drop-in:", report)
