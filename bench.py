#!/usr/bin/env python
"""bench.py — frame-pairs/s through the RPEFlow cost-volume stack on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--workload things|dsec|hd]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on this box's host cores

A step = one pass of rpeflow_b200.stack.CostVolumeStack over a batch of B synthetic frame pairs per GPU
(960x540 images -> 576x960 pyramid, 8192 points, 1 M events: BASELINE.json configs[0] shapes; the work is
sharded by sample, weak scaling, no collective on the data path — NCCL is only used for the barrier, the
max-over-ranks time and the cross-rank verification all-gather).  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=74,
                    help="frame pairs per GPU per step (74 pairs = 148 clouds: one FPS cloud per SM of a B200)")
    ap.add_argument("--workload", default="things", choices=["things", "dsec", "hd", "tiny"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample-s", type=float, default=20.0, help="target seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--eager", action="store_true", help="no CUDA graphs: launch every kernel from Python (debugging)")
    return ap.parse_args()


def baseline_meta():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        return json.load(f)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU every 100 ms while the timed region runs (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4))}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                mask = get(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------------------------- CPU reference arm
def cpu_frame_pair(cfg, host, weights):
    """One frame pair through the oracle's torch-CPU restatement of the reference path (same census as the stack)."""
    from oracle import torch_ref as R
    from rpeflow_b200.stack import LEVEL_CHANNELS, PYRAMID_POINTS
    hs, ws = cfg.sensor
    if cfg.name == "dsec":
        R.events_to_voxel_trilinear(host["ev_x"][0].numpy(), host["ev_y"][0].numpy(), host["ev_t"][0].numpy(),
                                    host["ev_p"][0].numpy(), cfg.event_bins, cfg.height, cfg.width, True)
    else:
        R.events_to_voxel(host["events"][0].numpy(), cfg.event_bins, cfg.height, cfg.width, True)
    pc1, pc2 = host["pcs"][:1, :3], host["pcs"][:1, 3:]
    picked = R.furthest_point_sampling(torch.cat([pc1, pc2], 0).transpose(1, 2), max(PYRAMID_POINTS))
    xyzs1, xyzs2 = [pc1], [pc2]
    for n in PYRAMID_POINTS:
        xyzs1.append(R.batch_indexing_channel_first(pc1, picked[:1, :n]))
        xyzs2.append(R.batch_indexing_channel_first(pc2, picked[1:, :n]))
    for lvl in range(5):
        for xyzs in (xyzs1, xyzs2):
            R.k_nearest_neighbor(xyzs[lvl], xyzs[lvl + 1], cfg.k)
    for lvl in range(5, 0, -1):
        h, w = cfg.level_hw(lvl)
        xyz1, xyz2 = xyzs1[lvl], xyzs2[lvl]
        f1_2d, f2_2d = host["feat2d"][lvl][0][:1], host["feat2d"][lvl][1][:1]
        f1_3d, f2_3d = host["feat3d"][lvl][0][:1], host["feat3d"][lvl][1][:1]
        ef_2d = host["efeat2d"][lvl][:1]
        dec_2d, dec_3d = host["flowfeat"][lvl][0][:1], host["flowfeat"][lvl][1][:1]

        def to_pixels(xyz):
            return torch.cat([(xyz[:, 0:1] + (ws - 1) / 2) * ((w - 1) / (ws - 1)),
                              (xyz[:, 1:2] + (hs - 1) / 2) * ((h - 1) / (hs - 1))], dim=1)
        xy1, xy2 = to_pixels(xyz1), to_pixels(xyz2)
        grid = R.pixel_grid(1, h, w)
        nn1 = R.k_nearest_neighbor(xy1, grid, 1)[..., 0]
        nn2 = R.k_nearest_neighbor(xy2, grid, 1)[..., 0]
        knn11 = R.k_nearest_neighbor(xyz1, xyz1, cfg.k)
        R.project_feat_with_nn_corr(xy1, f1_2d, f1_3d, nn1)
        R.project_feat_with_nn_corr(xy2, f2_2d, f2_3d, nn2)
        R.grid_sample_wrapper(f1_2d, xy1)
        R.grid_sample_wrapper(f2_2d, xy2)
        if lvl < 5:
            R.k_nearest_neighbor(xyzs1[lvl + 1], xyz1, 3)
            R.k_nearest_neighbor(xyz1, xyz2, 3)
        cost3d = R.correlation3d(xyz1, f1_3d, xyz2, f2_3d, weights[lvl], k=cfg.k, knn11=knn11)
        cost2d = R.correlation2d(f1_2d, f2_2d, cfg.max_displacement)
        R.project_feat_with_nn_corr(xy1, cost2d, torch.cat([cost3d, xyz1[:, :2]], 1), nn1)
        R.grid_sample_wrapper(torch.cat([cost2d, f1_2d[:, :2]], 1), xy1)
        R.grid_sample_wrapper(ef_2d, xy1)
        R.project_feat_with_nn_corr(xy1, dec_2d, dec_3d, nn1)
        R.grid_sample_wrapper(dec_2d, xy1)
    for i in range(5):
        R.k_nearest_neighbor(xyzs1[i + 1], xyzs1[i], 3)


def cpu_weights(cfg):
    import torch.nn as nn  # noqa: F401
    from rpeflow_b200.stack import LEVEL_CHANNELS
    torch.manual_seed(0)
    out = {}
    for lvl, c in zip(range(1, 6), LEVEL_CHANNELS):
        shapes = {"W1": (c, 2 * c + 3), "b1": (c,), "W2": (c, c), "b2": (c,)}
        for t in ("n1", "n2"):
            shapes.update({f"{t}_Wa": (8, 3), f"{t}_ba": (8,), f"{t}_Wb": (8, 8), f"{t}_bb": (8,), f"{t}_Wc": (c, 8), f"{t}_bc": (c,)})
        out[lvl] = {n: torch.randn(*s) * (1.0 / max(s[-1], 1)) ** 0.5 for n, s in shapes.items()}
    return out


def cpu_threads():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, n)


def run_reference_arm(args, cfg, meta):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: oracle/torch_ref.py),
    all host threads, one frame pair per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rpeflow_b200.stack import make_host_inputs
    threads = cpu_threads()
    torch.set_num_threads(threads)
    host = make_host_inputs(cfg, 1)
    weights = cpu_weights(cfg)
    t_budget = 280.0
    t0 = time.perf_counter()
    for _ in range(max(1, args.warmup)):
        cpu_frame_pair(cfg, host, weights)
        if time.perf_counter() - t0 > 60:
            break
    per = (time.perf_counter() - t0) / max(1, args.warmup)
    steps = args.steps
    note = None
    if per * steps > t_budget:                      # keep the whole run within a few minutes
        steps = max(1, int(t_budget / per))
        note = f"steps reduced from {args.steps} to {steps} to bound the CPU run (~{per:.1f} s per frame pair)"
    t1 = time.perf_counter()
    for _ in range(steps):
        cpu_frame_pair(cfg, host, weights)
    dt = time.perf_counter() - t1
    value = steps / dt
    line = {
        "impl": "reference", "metric": meta["metric"], "value": value, "unit": "frame-pairs/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args.batch, args),
        "cpu_baseline": {"value": value, "unit": "frame-pairs/s", "cores": threads, "kind": "port",
                         "sample": "1 frame pair per step, full op census, oracle/torch_ref.py (torch CPU restatement of the reference fallbacks)"},
        "e2e": {"value": value, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if note:
        line["note"] = note
    _emit(line)


def workload_config(cfg, batch, args):
    hp, wp = cfg.padded
    return {"workload": f"RPEFlow cost-volume stack, {cfg.name}: {cfg.width}x{cfg.height} (-> {wp}x{hp} pyramid, 5 levels), "
                        f"{cfg.n_points} points, {cfg.n_events} events, md=4, k=16; per frame pair: 1 voxelisation, 1 FPS, "
                        f"43 KNN, 5 corr2d, 5 Correlation3D, 20 project_feat_with_nn_corr, 25 grid_sample_wrapper",
            "frame_pairs_per_gpu_per_step": batch, "sharding": "by frame pair, no data-path collective",
            "l2_policy": "inputs larger than L2 (per-step working set > 1 GB vs 126 MB L2)"}


# ----------------------------------------------------------------------------------------------- GPU arm
_JSON_OUT = None


def _claim_stdout():
    """rank 0 prints ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner with printf on
    file descriptor 1), so the real stdout is kept aside for the JSON line and descriptor 1 is pointed at stderr."""
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def _emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse()
    _claim_stdout()
    meta = baseline_meta()
    from rpeflow_b200.stack import CONFIGS
    cfg = CONFIGS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, cfg, meta)
        return

    import torch.distributed as dist
    from rpeflow_b200 import _lib
    from rpeflow_b200.stack import (CostVolumeStack, census_work, make_host_inputs, tensors_nbytes, to_device)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path in rpeflow_b200)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL's debug stream (the "NCCL version ..." banner is printed at every
        # level >= VERSION, WARN included) goes to stderr instead of its default, stdout
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    host = make_host_inputs(cfg, B, first_sample=rank * B, pin=True)
    stack = CostVolumeStack(cfg, dev)
    x = to_device(host, dev)
    torch.cuda.synchronize()

    # -------- device-resident timing: W warm-up + exactly K timed steps, CUDA events on the launching stream.
    # A step = one replay of the CUDA graph of CostVolumeStack.run (every kernel of the census, two-stream overlap
    # of the point chain with the point-independent 2-D ops); --eager launches the same kernels from Python.
    from rpeflow_b200.stack import GraphedStack, HostFeeder
    if args.eager:
        step_fn = lambda: stack.run(x, overlap=True)
        l0 = _lib.LAUNCHES
        step_fn()
        launches_per_step = _lib.LAUNCHES - l0
    else:
        l0 = _lib.LAUNCHES
        graphed = GraphedStack(stack, x, fused=True, with_checksum=False)
        launches_per_step = (_lib.LAUNCHES - l0) // 2          # one eager warm-up + one capture
        step_fn = graphed.replay
    for _ in range(max(3, args.warmup)):
        step_fn()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_fn()
    ev1.record()
    barrier()
    clocks = sampler.finish()
    launches = launches_per_step * args.steps
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * B / (ms_step * 1e-3)

    # per-op device time (rank 0): a separate serial, eager pass with CUDA-event brackets around every op
    timers = []
    stack.run(x, timed=True)                            # allocator warm-up outside the measurement
    torch.cuda.synchronize()
    for _ in range(3):
        _, T = stack.run(x, timed=True)
        timers.append(T)
    torch.cuda.synchronize()
    per_op = {}
    for T in timers:                                   # minimum over the passes: allocator growth lands in one of them
        for name, (ms, n) in T.totals_ms().items():
            cur = per_op.get(name)
            if cur is None or ms < cur["ms_per_step"]:
                per_op[name] = {"ms_per_step": ms, "calls_per_step": n}

    # -------- roofline of the dominant HBM-bound kernel: corr2d at pyramid level 1
    work = census_work(cfg)
    peak, peak_src = measured_peaks()
    c2 = per_op["corr2d_L1"]
    # correlation2d() on the NCHW level-1 maps = one launch of corr2d_fwd_diag_kernel (no permutes); timed alone here
    # with an L2 flush between launches, and inside the step by the per-op brackets.
    f1, f2 = x["feat2d"][1][0], x["feat2d"][1][1]
    from rpeflow_b200 import ops
    for _ in range(3):
        ops.correlation2d(f1, f2, cfg.max_displacement)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    kms = []
    for _ in range(10):
        flush.zero_()                                   # L2 flush between isolated launches (256 MB > 126 MB L2)
        k0.record()
        ops.correlation2d(f1, f2, cfg.max_displacement)
        k1.record()
        k1.synchronize()
        kms.append(k0.elapsed_time(k1))
    kms.sort()
    corr_ms = kms[len(kms) // 2]
    corr_bytes = work["corr2d_bytes"][1] * B
    achieved = corr_bytes / (corr_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")     # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tpath):
        with open(tpath) as f:
            per_pair = json.load(f).get(f"corr2d_fwd_L1_{cfg.name}_per_frame_pair")
            traffic = per_pair * B if per_pair else None
    roofline = {"kernel": "corr2d_fwd_diag_kernel (level 1: C=32, %dx%d, batch %d)" % (*cfg.level_hw(1), B),
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": corr_bytes,
                "launch_ms": corr_ms, "in_step_ms": c2["ms_per_step"]}
    del flush

    # secondary per-op figures (not HBM-bound ones are reported in their own unit)
    knn_ms = sum(v["ms_per_step"] for k, v in per_op.items() if k.startswith("knn"))
    ops_report = {
        "per_op_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(per_op.items())},
        "knn_gpairs_per_s": work["knn_pairs"] * B / (knn_ms * 1e-3) / 1e9 if knn_ms else None,
        "fps_ns_per_iteration": per_op["fps"]["ms_per_step"] * 1e6 / max(1, 4096 - 1),
        "gathers_gbs": work["gather_bytes"] * B / (1e-3 * (per_op["grid_sample"]["ms_per_step"] + per_op["project_nn_corr"]["ms_per_step"])) / 1e9,
        "event_voxel_gbs": work["event_voxel_bytes"] * B / (1e-3 * per_op["event_voxel"]["ms_per_step"]) / 1e9,
        "corr3d_tflops_dense_equiv": work["corr3d_flops"] * B / (1e-3 * per_op["corr3d"]["ms_per_step"]) / 1e12,
    }

    # -------- end to end: every step copies ALL of its inputs from pinned host memory (copy stream, group by group
    # in dependency order) into one of two device input sets, replays the per-group CUDA graphs as the groups land,
    # and reads the step's checksum vector back to the host.  Step s+1's copies overlap step s's compute.
    e2e = None
    if not args.no_e2e:
        if not args.eager:
            del graphed
        feeder = HostFeeder(host, dev, depth=2)
        if args.eager:
            class _Eager:
                def __init__(self, xs):
                    self.xs = xs

                def replay(self, wait=None):
                    out, _ = stack.run(self.xs, wait=wait)
                    self.checksum = stack.checksum(out)
            runners = [_Eager(sl) for sl in feeder.slots]
        else:
            runners = [GraphedStack(stack, sl, fused=False) for sl in feeder.slots]
        result_host = torch.zeros((2, 2, 64), dtype=torch.float64).pin_memory()
        main_stream = torch.cuda.current_stream()
        consumed = []

        def e2e_run(steps):
            done = [None, None]
            pending = feeder.issue(0)
            for s in range(steps):
                slot = s % 2
                evs = pending
                if s + 1 < steps:                       # next step's copies start as soon as its device set is free
                    pending = feeder.issue((s + 1) % 2, after=done[(s + 1) % 2])
                runners[slot].replay(wait=lambda g, evs=evs: main_stream.wait_event(evs[g]))
                ints, flts = runners[slot].checksum
                result_host[slot, 0, :ints.numel()].copy_(ints, non_blocking=True)
                result_host[slot, 1, :flts.numel()].copy_(flts, non_blocking=True)
                d = torch.cuda.Event()
                d.record(main_stream)
                done[slot] = d
                if s >= 1:                              # the host reads step s-1's result while step s runs
                    done[(s - 1) % 2].synchronize()
                    consumed.append(float(result_host[(s - 1) % 2, 1, 0]))
            main_stream.synchronize()
            consumed.append(float(result_host[(steps - 1) % 2, 1, 0]))
            return ints.numel() + flts.numel()
        nres = e2e_run(2)
        d2h_bytes = 8 * nres
        barrier()
        w0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(args.steps)
        e1.record()
        barrier()
        wall = time.perf_counter() - w0
        te = torch.tensor([max(e0.elapsed_time(e1) * 1e-3, 0.0)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * args.steps / float(te.item()), "unit": "frame-pairs/s",
               "h2d_bytes_per_step": feeder.nbytes, "d2h_bytes_per_step": d2h_bytes,
               "h2d_gbs": feeder.nbytes * args.steps / float(te.item()) / 1e9,
               "pipeline": "2 device input sets; copies of step s+1 (copy stream, 7 groups in dependency order) overlap the "
                           "graph replays of step s; result read back every step",
               "inputs_copied": "point clouds, raw events and every synthetic activation the ops read",
               "wall_s": wall}
        del runners, feeder

    # -------- cross-rank verification (NCCL all-gather of checksums of one common sample)
    verify = None
    if world > 1:
        common = to_device(make_host_inputs(cfg, 1, first_sample=0), dev)
        out, _ = stack.run(common)
        ints, flts = stack.checksum(out)
        gi = [torch.empty_like(ints) for _ in range(world)]
        gf = [torch.empty_like(flts) for _ in range(world)]
        dist.all_gather(gi, ints)
        dist.all_gather(gf, flts)
        ok_i = all(torch.equal(g, gi[0]) for g in gi)
        ok_f = all(torch.allclose(g, gf[0], rtol=1e-5, atol=1e-3) for g in gf)
        verify = {"index_checksums_identical": bool(ok_i), "float_checksums_close": bool(ok_f)}

    # -------- CPU baseline (rank 0, N=1 only): the oracle port on a bounded sample
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = cpu_threads()
        torch.set_num_threads(threads)
        host1 = make_host_inputs(cfg, 1)
        wts = cpu_weights(cfg)
        cpu_frame_pair(cfg, host1, wts)                # warm-up
        n, t0 = 0, time.perf_counter()
        while True:
            cpu_frame_pair(cfg, host1, wts)
            n += 1
            if time.perf_counter() - t0 > args.cpu_sample_s or n >= 8:
                break
        dtc = time.perf_counter() - t0
        cpu_baseline = {"value": n / dtc, "unit": "frame-pairs/s", "cores": threads, "kind": "port",
                        "sample": f"{n} frame pair(s) of the same workload through oracle/torch_ref.py (torch CPU restatement of "
                                  f"the reference fallbacks), {dtc:.1f} s"}

    if rank == 0:
        line = {
            "metric": meta["metric"], "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(cfg, B, args),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "ops": ops_report,
        }
        if verify:
            line["verify"] = verify
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
