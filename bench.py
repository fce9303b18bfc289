#!/usr/bin/env python
"""bench.py — frame-pairs/s through the RPEFlow cost-volume stack on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--workload things|dsec|hd|hd_scaled]
                    [--scaling weak|strong]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the reference's own CPU path on this box's host cores

A step = one pass of rpeflow_b200.stack.CostVolumeStack over a batch of B synthetic frame pairs per GPU
(960x540 images -> 576x960 pyramid, 8192 points, 1 M events: BASELINE.json configs[0] shapes; the work is
sharded by sample, no collective on the data path — NCCL is only used for the barrier, the max-over-ranks time and
the cross-rank verification all-gather).  Rank 0 prints ONE JSON line.  Keys beyond the driver's contract:
  roofline      the op with the largest share of the step among the HBM-bound ops + `per_op`: one entry for EVERY op
                (HBM-bound ones in GB/s against the measured copy peak; Correlation3D / KNN / FPS in their own units)
  ops           per-op device milliseconds (serial pass), the same kernels' reference-CUDA counterparts
                (`vs_ref_cuda`, timed through oracle/_ref on the same inputs) and the k = 32 searches of configs[2]
  model_e2e     the UNMODIFIED reference model (RPEFlow.forward at 960x540 / 8192 points) driven from host buffers
                with this library installed underneath, next to the same model on the reference's own CUDA kernels
"""
import argparse
import importlib.util
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
    ap.add_argument("--batch", type=int, default=None,
                    help="frame pairs per GPU per step (default workload.BENCH_BATCH = 148 pairs = 296 clouds: two FPS chains per "
                         "SM of a B200; 74 = one per SM); with --scaling strong it is the GLOBAL batch, split over the GPUs")
    ap.add_argument("--workload", default="things", choices=["things", "dsec", "hd", "hd_scaled", "tiny"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample-s", type=float, default=20.0, help="target seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-model", action="store_true", help="skip the model_e2e leg (reference model through the drop-in)")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference-CUDA-kernel column (ops.vs_ref_cuda)")
    ap.add_argument("--model-batch", type=int, default=4, help="frame pairs per forward in the model leg (conf/test/things.yaml:15)")
    ap.add_argument("--eager", action="store_true", help="no CUDA graphs: launch every kernel from Python (debugging / ncu)")
    return ap.parse_args()


def baseline_meta():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        return json.load(f)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p.get("bf16_tflops_sustained", p.get("bf16_tflops", 1366.4))),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def load_workload_module():
    """rpeflow_b200/workload.py loaded BY PATH: shapes + synthetic inputs without importing the package, so the
    reference arm's process never maps libb200flow.so."""
    name = "_b200_workload"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "rpeflow_b200", "workload.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def reference_tree():
    """The staged copy of the reference's python tree (oracle/_ref/reference, `make -C oracle reftree`); it travels to the
    GPU box git-ignored, like the prebuilt checker libraries."""
    path = os.path.join(ROOT, "oracle", "_ref", "reference")
    return path if os.path.isdir(os.path.join(path, "models")) else None


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
class _PortOps:
    """The oracle's torch-CPU restatement (oracle/torch_ref.py) — used only when no reference tree travelled."""
    kind = "port"
    what = "oracle/torch_ref.py (torch CPU restatement of the reference fallbacks)"

    def __init__(self, W, cfg):
        from oracle import torch_ref as R
        self.R, self.W = R, W
        torch.manual_seed(0)
        self.weights = {}
        for lvl, c in zip(range(1, 6), W.LEVEL_CHANNELS):
            shapes = {"W1": (c, 2 * c + 3), "b1": (c,), "W2": (c, c), "b2": (c,)}
            for t in ("n1", "n2"):
                shapes.update({f"{t}_Wa": (8, 3), f"{t}_ba": (8,), f"{t}_Wb": (8, 8), f"{t}_bb": (8,), f"{t}_Wc": (c, 8), f"{t}_bc": (c,)})
            self.weights[lvl] = {n: torch.randn(*s) * (1.0 / max(s[-1], 1)) ** 0.5 for n, s in shapes.items()}
        self.fps = R.furthest_point_sampling
        self.knn = R.k_nearest_neighbor
        self.gather_cf = R.batch_indexing_channel_first
        self.grid_sample = R.grid_sample_wrapper
        self.project = R.project_feat_with_nn_corr
        self.corr2d = R.correlation2d
        self.pixel_grid = R.pixel_grid

    def voxelise(self, cfg, host):
        R = self.R
        if cfg.name == "dsec":
            R.events_to_voxel_trilinear(host["ev_x"][0].numpy(), host["ev_y"][0].numpy(), host["ev_t"][0].numpy(),
                                        host["ev_p"][0].numpy(), cfg.event_bins, cfg.height, cfg.width, True)
        else:
            R.events_to_voxel(host["events"][0].numpy(), cfg.event_bins, cfg.height, cfg.width, True)

    def corr3d(self, lvl, xyz1, f1, xyz2, f2, knn11, k):
        return self.R.correlation3d(xyz1, f1, xyz2, f2, self.weights[lvl], k=k, knn11=knn11)


class _ReferenceOps:
    """The UNMODIFIED reference functions imported from the staged tree: models.csrc.wrapper fallbacks,
    models.pwc3d_core.Correlation3D, models.utils helpers, event_utils.eventsToVoxel / dsec.DSECTrain.eventsToVoxelInter."""
    kind = "reference"
    what = "the unmodified reference torch path (models.csrc.wrapper fallbacks, Correlation3D, models.utils, event_utils) from the staged tree"

    def __init__(self, W, cfg, tree):
        import types
        sys.dont_write_bytecode = True
        if tree not in sys.path:
            sys.path.insert(0, tree)
        for m in ("hdf5plugin", "h5py", "imageio", "skimage", "omegaconf", "matplotlib", "matplotlib.colors", "cv2"):
            try:
                __import__(m)
            except Exception:
                sys.modules.setdefault(m, types.ModuleType(m))
        if not hasattr(sys.modules["h5py"], "File"):
            sys.modules["h5py"].File = object
        om = sys.modules["omegaconf"]
        if not hasattr(om, "OmegaConf"):
            om.OmegaConf, om.DictConfig = object, dict
        if not hasattr(sys.modules["matplotlib.colors"], "hsv_to_rgb"):
            sys.modules["matplotlib.colors"].hsv_to_rgb = None
        import models.csrc.wrapper as wrapper
        import models.pwc3d_core as p3
        import models.utils as mu
        import event_utils
        self.mu, self.event_utils = mu, event_utils
        self.fps, self.knn, self.corr2d = wrapper.furthest_point_sampling, wrapper.k_nearest_neighbor, wrapper.correlation2d
        self.gather_cf, self.grid_sample, self.project = mu.batch_indexing_channel_first, mu.grid_sample_wrapper, mu.project_feat_with_nn_corr
        torch.manual_seed(0)
        self.mods = {lvl: p3.Correlation3D(c, c, k=cfg.k).eval() for lvl, c in zip(range(1, 6), W.LEVEL_CHANNELS)}
        self.dsec = None
        if cfg.name == "dsec":
            n_threads = torch.get_num_threads()
            import dsec                                  # dsec.py:7 pins torch to one thread at import: undo it
            torch.set_num_threads(n_threads)
            self.dsec = dsec

    def pixel_grid(self, b, h, w):
        return self.mu.mesh_grid(b, h, w, "cpu").reshape(b, 2, -1)

    def voxelise(self, cfg, host):
        if cfg.name == "dsec":
            ev = {"x": host["ev_x"][0].numpy(), "y": host["ev_y"][0].numpy(), "t": host["ev_t"][0].numpy(), "p": host["ev_p"][0].numpy()}
            self.dsec.DSECTrain.eventsToVoxelInter(self.dsec.DSECTrain.__new__(self.dsec.DSECTrain), ev, cfg.event_bins,
                                                   cfg.height, cfg.width, event_polarity=True)
        else:
            self.event_utils.eventsToVoxel(host["events"][0].numpy(), num_bins=cfg.event_bins, height=cfg.height,
                                           width=cfg.width, event_polarity=True)

    def corr3d(self, lvl, xyz1, f1, xyz2, f2, knn11, k):
        return self.mods[lvl](xyz1, f1, xyz2, f2, knn11)


def make_cpu_ops(W, cfg):
    tree = reference_tree()
    if tree is not None:
        try:
            return _ReferenceOps(W, cfg, tree)
        except Exception as e:                           # a half-staged tree must not cost the bench line
            print(f"bench.py: reference tree at {tree} unusable ({e!r}); using the oracle port", file=sys.stderr)
    return _PortOps(W, cfg)


@torch.no_grad()
def cpu_frame_pair(cfg, host, R):
    """One frame pair through the reference's CPU path, same op census as CostVolumeStack."""
    hs, ws = cfg.sensor
    R.voxelise(cfg, host)
    pc1, pc2 = host["pcs"][:1, :3], host["pcs"][:1, 3:]
    picked = R.fps(torch.cat([pc1, pc2], 0).transpose(1, 2), max(cfg.pyramid))
    xyzs1, xyzs2 = [pc1], [pc2]
    for n in cfg.pyramid:
        xyzs1.append(R.gather_cf(pc1, picked[:1, :n]))
        xyzs2.append(R.gather_cf(pc2, picked[1:, :n]))
    for lvl in range(5):
        for xyzs in (xyzs1, xyzs2):
            R.knn(xyzs[lvl], xyzs[lvl + 1], cfg.k)
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
        nn1 = R.knn(xy1, grid, 1)[..., 0]
        nn2 = R.knn(xy2, grid, 1)[..., 0]
        knn11 = R.knn(xyz1, xyz1, cfg.k)
        R.project(xy1, f1_2d, f1_3d, nn1)
        R.project(xy2, f2_2d, f2_3d, nn2)
        R.grid_sample(f1_2d, xy1)
        R.grid_sample(f2_2d, xy2)
        if lvl < 5:
            R.knn(xyzs1[lvl + 1], xyz1, 3)
            R.knn(xyz1, xyz2, 3)
        cost3d = R.corr3d(lvl, xyz1, f1_3d, xyz2, f2_3d, knn11, cfg.k)
        cost2d = R.corr2d(f1_2d, f2_2d, cfg.max_displacement)
        R.project(xy1, cost2d, torch.cat([cost3d, xyz1[:, :2]], 1), nn1)
        R.grid_sample(torch.cat([cost2d, f1_2d[:, :2]], 1), xy1)
        R.grid_sample(ef_2d, xy1)
        R.project(xy1, dec_2d, dec_3d, nn1)
        R.grid_sample(dec_2d, xy1)
    for i in range(5):
        R.knn(xyzs1[i + 1], xyzs1[i], 3)


def cpu_threads():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, n)


def _maps_product_library():
    try:
        with open("/proc/self/maps") as f:
            return "libb200flow" in f.read()
    except OSError:
        return None


def run_reference_arm(args, meta):
    """--impl reference: the reference's own CPU implementation of the path — the unmodified reference functions from the
    staged tree (kind "reference"), else the oracle port — all host threads, one frame pair per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    W = load_workload_module()
    cfg = W.CONFIGS[args.workload]
    threads = cpu_threads()
    torch.set_num_threads(threads)
    host = W.make_host_inputs(cfg, 1)
    R = make_cpu_ops(W, cfg)
    torch.set_num_threads(threads)
    t_budget = 280.0
    t0 = time.perf_counter()
    for _ in range(max(1, args.warmup)):
        cpu_frame_pair(cfg, host, R)
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
        cpu_frame_pair(cfg, host, R)
    dt = time.perf_counter() - t1
    value = steps / dt
    line = {
        "impl": "reference", "metric": meta["metric"], "value": value, "unit": "frame-pairs/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg, args.batch, args, 1),
        "cpu_baseline": {"value": value, "unit": "frame-pairs/s", "cores": threads, "kind": R.kind,
                         "sample": f"1 frame pair per step, full op census, {R.what}"},
        "e2e": {"value": value, "unit": "frame-pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "product_library_mapped": _maps_product_library(),      # must be false: this arm is the reference alone
    }
    if note:
        line["note"] = note
    _emit(line)


def workload_config(cfg, batch, args, world):
    hp, wp = cfg.padded
    return {"workload": f"RPEFlow cost-volume stack, {cfg.name}: {cfg.width}x{cfg.height} (-> {wp}x{hp} pyramid, 5 levels), "
                        f"{cfg.n_points} points (pyramid {list(cfg.pyramid)}), {cfg.n_events} events, md=4, k=16; per frame pair: "
                        f"1 voxelisation, 1 FPS, 43 KNN, 5 corr2d, 5 Correlation3D, 20 project_feat_with_nn_corr, 25 grid_sample_wrapper "
                        f"(PointConv bodies, convolutions and attention are outside the census: SURVEY §8d)",
            "frame_pairs_per_gpu_per_step": batch, "scaling": args.scaling,
            "sharding": "by frame pair, no data-path collective",
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


def _median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2]


class _Flush:
    """L2 flush between isolated launches: a 256 MB write (> 126 MB L2)."""

    def __init__(self, dev):
        self.buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def __call__(self):
        self.buf.zero_()


def time_isolated(fn, flush, reps=7, warm=2):
    """Median CUDA-event time (ms) of fn() on the current stream with an L2 flush before every launch."""
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(reps):
        flush()
        a.record()
        fn()
        b.record()
        b.synchronize()
        ms.append(a.elapsed_time(b))
    return _median(ms)


def per_op_times(stack, x):
    """Serial, eager pass with CUDA-event brackets around every op; minimum over 3 passes."""
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
    return per_op


def roofline_report(cfg, work, per_op, B, peaks, traffic_db):
    """One entry per op of the census.  HBM-bound ops: algorithmic bytes (SURVEY §8d formulas x B) / in-step device time vs
    the measured copy peak.  The others in their own unit (tensor pipe, pairs/s, ns per dependent iteration)."""
    peak = peaks["hbm_gbs"]
    total_ms = sum(v["ms_per_step"] for v in per_op.values())
    entries = []

    def hbm(op, key_prefix, nbytes, kernel):
        ms = sum(v["ms_per_step"] for k, v in per_op.items() if k.startswith(key_prefix))
        calls = sum(v["calls_per_step"] for k, v in per_op.items() if k.startswith(key_prefix))
        if ms <= 0:
            return
        gbs = nbytes / (ms * 1e-3) / 1e9
        t = traffic_db.get(f"{op}_{cfg.name}_per_frame_pair")
        entries.append({"op": op, "kernel": kernel, "bound": "hbm", "algorithmic_bytes": nbytes, "ms": round(ms, 4), "calls": calls,
                        "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 4),
                        "share_of_step": round(ms / total_ms, 4), "traffic": t * B if t else None})
    for lvl in range(1, 6):
        hbm(f"corr2d_L{lvl}", f"corr2d_L{lvl}", work["corr2d_bytes"][lvl] * B, "corr2d_fwd_{diag,nchw}_kernel")
    for lvl in range(1, 6):
        hbm(f"project_nn_corr_L{lvl}", f"project_nn_corr_L{lvl}", work["project_bytes"][lvl] * B,
            "project_prep_kernel + project_tile_kernel + project_post_kernel (maps of >= 4096 px TMA can address) / "
            "sample_point_major_kernel + project_nn_corr_kernel (small maps)")
    for lvl in range(1, 6):
        hbm(f"grid_sample_L{lvl}", f"grid_sample_L{lvl}", work["grid_sample_bytes"][lvl] * B,
            "grid_sample_pts_kernel (4 of the 5 calls per level are served by the projection's sampler)")
    hbm("event_voxel", "event_voxel", work["event_voxel_bytes"] * B, "event_voxel_*_kernel")
    hbm("gather_xyz", "gather_xyz", work["gather_xyz_bytes"] * B, "gather_cf_kernel")
    # families
    fam = {}
    for name, pref, nbytes in (("corr2d_all_levels", "corr2d_L", sum(work["corr2d_bytes"].values()) * B),
                               ("projection_gathers", ("project_nn_corr", "grid_sample"), work["gather_bytes"] * B)):
        ms = sum(v["ms_per_step"] for k, v in per_op.items() if k.startswith(pref))
        if ms > 0:
            gbs = nbytes / (ms * 1e-3) / 1e9
            fam[name] = {"algorithmic_bytes": nbytes, "ms": round(ms, 4), "achieved": round(gbs, 1), "frac": round(gbs / peak, 4),
                         "share_of_step": round(ms / total_ms, 4)}
    # not HBM-bound: own units
    c3 = per_op.get("corr3d")
    if c3:
        tf32_peak = peaks["bf16_tflops"] / 2.0
        mma = 3 * work["corr3d_mma_flops"] * B if cfg.precision == 2 else work["corr3d_mma_flops"] * B
        entries.append({"op": "corr3d", "kernel": "corr3d_stage1_tc_kernel + stage2 + pointwise linears", "bound": "tensor",
                        "ms": round(c3["ms_per_step"], 4), "calls": c3["calls_per_step"],
                        "achieved": round(mma / (c3["ms_per_step"] * 1e-3) / 1e12, 2), "peak": round(tf32_peak, 1), "unit": "TFLOP/s",
                        "frac": round(mma / (c3["ms_per_step"] * 1e-3) / 1e12 / tf32_peak, 4),
                        "executed_mma_flops": mma, "dense_equivalent_tflops": round(work["corr3d_flops"] * B / (c3["ms_per_step"] * 1e-3) / 1e12, 2),
                        "peak_source": "tf32 = measured dense bf16 / 2", "share_of_step": round(c3["ms_per_step"] / total_ms, 4)})
    for grp, pairs in work["knn_pairs_by_group"].items():
        v = per_op.get("knn_" + grp)
        if v:
            entries.append({"op": "knn_" + grp, "kernel": "knn_grid_build_kernel + knn_grid_query[_batched]_kernel", "bound": "issue",
                            "ms": round(v["ms_per_step"], 4), "calls": v["calls_per_step"],
                            "achieved": round(pairs * B / (v["ms_per_step"] * 1e-3) / 1e9, 1), "unit": "G brute-force-equivalent pairs/s",
                            "share_of_step": round(v["ms_per_step"] / total_ms, 4)})
    f = per_op.get("fps")
    if f:
        iters = max(cfg.pyramid) - 1
        entries.append({"op": "fps", "kernel": "fps_pruned_kernel / fps_kernel", "bound": "latency", "ms": round(f["ms_per_step"], 4),
                        "calls": 1, "achieved": round(f["ms_per_step"] * 1e6 / iters, 1), "unit": "ns per dependent iteration",
                        "clouds": 2 * B, "share_of_step": round(f["ms_per_step"] / total_ms, 4)})
    return entries, fam, total_ms


def ref_cuda_column(cfg, stack, x, B, per_op, flush):
    """The reference's own CUDA kernels (oracle/_ref/libref_kernels.so: its .cu files compiled for sm_100a) on the same
    device inputs: FPS, the KNN groups and correlation2d per level, new-vs-old milliseconds.  Checker leg, outside every
    timed region of the step."""
    from oracle import refcuda
    if not refcuda.available():
        return {"unavailable": "oracle/_ref/libref_kernels.so not built"}
    from rpeflow_b200 import ops
    out = {}
    ev = [torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)]

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(reps):
            flush()
            ev[0].record()
            fn()
            ev[1].record()
            ev[1].synchronize()
            ms.append(ev[0].elapsed_time(ev[1]))
        return _median(ms)
    # FPS
    pc1, pc2 = x["pcs"][:, :3].contiguous(), x["pcs"][:, 3:].contiguous()
    both = torch.cat([pc1, pc2], 0).transpose(1, 2).contiguous()
    n_s = max(cfg.pyramid)
    if both.shape[1] <= 65536:
        new = timed(lambda: ops.furthest_point_sampling(both, n_s))
        old = timed(lambda: refcuda.fps(both, n_s, sync=False))
        same = bool(torch.equal(ops.furthest_point_sampling(both, n_s), refcuda.fps(both, n_s)))
        out["fps"] = {"b200_ms": round(new, 4), "ref_cuda_ms": round(old, 4), "speedup": round(old / new, 2), "indices_identical": same,
                      "shape": list(both.shape)}
    # KNN: one representative call per group at level 1 shapes + the whole pyramid set
    picked = ops.furthest_point_sampling(both, n_s)
    from rpeflow_b200 import projection
    xyz0 = pc1
    lv = [xyz0] + [projection.batch_indexing_channel_first(pc1, picked[:B, :n]) for n in cfg.pyramid]
    cl = [t.transpose(1, 2).contiguous() for t in lv]
    h, w = cfg.level_hw(1)
    hs_, ws_ = cfg.sensor
    xy1 = torch.stack([(lv[1][:, 0] + (ws_ - 1) / 2) * ((w - 1) / (ws_ - 1)), (lv[1][:, 1] + (hs_ - 1) / 2) * ((h - 1) / (hs_ - 1))], 2).contiguous()
    grid = stack.pixel_grid(B, h, w).transpose(1, 2).contiguous()
    cases = {"pyramid_k16 (N0->N1)": (cl[0], cl[1], 16), "self_k16 (L1)": (cl[1], cl[1], 16), "interp_k3 (L2->L1)": (cl[2], cl[1], 3),
             "2d_k1 (L1 points -> pixel grid)": (xy1, grid, 1), "self_k32 (L1, configs[2])": (cl[1], cl[1], 32)}
    knn = {}
    for name, (inp, qry, k) in cases.items():
        new = timed(lambda: ops._k_nearest_neighbor_cuda(inp, qry, k))
        old = timed(lambda: refcuda.knn(inp, qry, k, sync=False), reps=2)
        a, b = ops._k_nearest_neighbor_cuda(inp, qry, k), refcuda.knn(inp, qry, k)
        knn[name] = {"b200_ms": round(new, 4), "ref_cuda_ms": round(old, 4), "speedup": round(old / new, 2),
                     "index_mismatch_fraction": float((a != b).float().mean().item()), "shape": [list(inp.shape), list(qry.shape), k]}
    out["knn"] = knn
    # correlation2d per level (reference: NHWC in, + the two permutes its wrapper does)
    c2 = {}
    for lvl in range(1, 6):
        f1, f2 = x["feat2d"][lvl]
        new = timed(lambda: ops.correlation2d(f1, f2, cfg.max_displacement))
        n1, n2 = f1.permute(0, 2, 3, 1).contiguous(), f2.permute(0, 2, 3, 1).contiguous()
        old_kernel = timed(lambda: refcuda.corr2d_fwd(n1, n2, cfg.max_displacement, sync=False))
        old_wrapper = timed(lambda: refcuda.corr2d_fwd(f1.permute(0, 2, 3, 1).contiguous(), f2.permute(0, 2, 3, 1).contiguous(),
                                                       cfg.max_displacement, sync=False))
        c2[f"L{lvl}"] = {"b200_ms": round(new, 4), "ref_cuda_kernel_ms": round(old_kernel, 4), "ref_cuda_with_wrapper_permutes_ms": round(old_wrapper, 4),
                         "speedup_vs_wrapper": round(old_wrapper / new, 2)}
        del n1, n2
    out["corr2d"] = c2
    out["note"] = ("ref_cuda = the reference's unmodified .cu files compiled for sm_100a (oracle/_ref), same device inputs, L2 flushed "
                   "before every launch; refcuda.* allocates/zero-fills its outputs like the reference's C++ wrappers do")
    return out


@torch.no_grad()
def pointconv_leg(cfg, x, B, dev):
    """The 20 PointConv bodies of one forward at the bench batch: 10 PointConvDownSampling of the feature
    pyramid (pwc3d_core.py:44-57; C = 32..192, N_l -> N_l+1, two clouds) and 10 PointConvNoSampling of FlowEstimator3D
    (pwc3d_core.py:118-143; 195 -> 128 -> 128 at every level).  Outside the census (and `value`), reported beside it."""
    from rpeflow_b200 import ops, pointconv, projection
    torch.manual_seed(0)
    pcs = [x["pcs"][:, :3].contiguous(), x["pcs"][:, 3:].contiguous()]
    both = torch.cat(pcs, 0).transpose(1, 2).contiguous()
    picked = ops.furthest_point_sampling(both, max(cfg.pyramid))
    pyr = [[pc] + [projection.batch_indexing_channel_first(pc, picked[i * B:(i + 1) * B, :n]) for n in cfg.pyramid] for i, pc in enumerate(pcs)]
    chans = [32, 64, 96, 128, 192]
    down = [pointconv.PointConvDownSampling(c, c).to(dev).eval() for c in chans]
    est1, est2 = pointconv.PointConvNoSampling(195, 128).to(dev).eval(), pointconv.PointConvNoSampling(128, 128).to(dev).eval()
    feats = [[torch.randn(B, chans[l], pyr[i][l].shape[2], device=dev) for l in range(5)] for i in range(2)]
    efeat = [torch.randn(B, 195, n, device=dev) for n in cfg.pyramid]
    knn = [ops.k_nearest_neighbor(pyr[0][l + 1], pyr[0][l + 1], 16) for l in range(5)]

    def run():
        for i in range(2):
            for l in range(5):
                down[l](pyr[i][l], feats[i][l], pyr[i][l + 1])
        for l in range(5):
            est2(pyr[0][l + 1], est1(pyr[0][l + 1], efeat[l], knn[l]), knn[l])
    run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = []
    for _ in range(3):
        a.record(); run(); b.record(); b.synchronize()
        ms.append(a.elapsed_time(b))
    return {"ms_per_step": round(min(ms), 4), "calls": 20, "includes": "the 10 pyramid KNN searches inside PointConvDownSampling",
            "note": "not part of the census / `value` (SURVEY §8d fixes the census); reported for visibility"}


def model_leg(args, dev):
    """model_e2e: the UNMODIFIED reference model (RPEFlow.forward, eval_withocc.py:54-63) at 960x540 / 8192 points driven
    from pinned host buffers (uint8 images, point clouds, intrinsics, events) to host flow buffers, every step:
       b200       install(): raw events -> GPU voxeliser -> forward on this library's kernels
       ref_cuda   the reference with ITS OWN CUDA extensions (oracle/_ref/ref_ext); event voxels precomputed on the host
                  (its datasets voxelise in DataLoader workers), so its H2D carries the [20,H,W] grid instead of raw events
    """
    tree = reference_tree()
    if tree is None:
        return {"unavailable": "no staged reference tree (make -C oracle reftree)"}
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from refmodel_util import reference_extensions_available, reference_extensions_bound
    from rpeflow_b200 import events as b200_events
    from rpeflow_b200 import refhost
    import rpeflow_b200.install as inst
    B, H, Wd, N, n_ev = args.model_batch, 540, 960, 8192, 1_000_000
    model = refhost.build_rpeflow(tree, device=dev, install=False, seed=0)
    host = refhost.synthetic_model_inputs(B, H, Wd, N, seed=3)
    g = torch.Generator().manual_seed(11)
    ev = torch.empty(B, n_ev, 4)
    ev[..., 0] = torch.randint(0, Wd, (B, n_ev), generator=g).float()
    ev[..., 1] = torch.randint(0, H, (B, n_ev), generator=g).float()
    ev[..., 2] = torch.sort(torch.rand(B, n_ev, generator=g), dim=1).values
    ev[..., 3] = torch.randint(0, 2, (B, n_ev), generator=g).float() * 2 - 1
    pin = {"images": host["images"].to(torch.uint8).pin_memory(), "pcs": host["pcs"].pin_memory(),
           "intrinsics": host["intrinsics"].pin_memory(), "events": ev.pin_memory()}
    out2d = torch.empty(B, 2, H, Wd).pin_memory()
    out3d = torch.empty(B, 3, N).pin_memory()

    def step_b200():
        d = {k: v.to(dev, non_blocking=True) for k, v in pin.items()}
        vox = torch.empty((B, 20, H, Wd), dtype=torch.float32, device=dev)
        for i in range(B):
            b200_events.events_to_voxel_device(d["events"][i], 10, H, Wd, True, check_range=False, out=vox[i])
        o = refhost.forward(model, {"images": d["images"], "pcs": d["pcs"], "intrinsics": d["intrinsics"], "event_voxel": vox})
        out2d.copy_(o["flow_2d"], non_blocking=True)
        out3d.copy_(o["flow_3d"], non_blocking=True)
        torch.cuda.synchronize()

    def run(step, steps, warm):
        for _ in range(warm):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        return (time.perf_counter() - t0) / steps

    res = {"batch": B, "shape": f"{Wd}x{H}, {N} points, {n_ev} events", "unit": "frame-pairs/s",
           "d2h_bytes_per_step": out2d.numel() * 4 + out3d.numel() * 4}
    steps = max(2, min(args.steps, 5))
    inst.install()
    inst.stats(reset=True)
    try:
        dt = run(step_b200, steps, 2)
        st = inst.stats()
        flows_b = (out2d.clone(), out3d.clone())
    finally:
        inst.uninstall()
    res["b200"] = {"value": B / dt, "ms_per_forward": 1e3 * dt,
                   "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in pin.values()),
                   "ops_on_b200_kernels": sum(v["b200"] for v in st.values()), "ops_on_reference_fallback": sum(v["reference"] for v in st.values())}
    res["value"] = res["b200"]["value"]
    # the reference arms get the voxel grid precomputed (this library's voxeliser, checked against the oracle in tests)
    vox_host = torch.empty(B, 20, H, Wd)
    for i in range(B):
        vox_host[i] = b200_events.events_to_voxel_device(pin["events"][i].to(dev), 10, H, Wd, True, check_range=False).cpu()
    pin_ref = {"images": pin["images"], "pcs": pin["pcs"], "intrinsics": pin["intrinsics"], "event_voxel": vox_host.pin_memory()}

    def step_ref():
        d = {k: v.to(dev, non_blocking=True) for k, v in pin_ref.items()}
        o = refhost.forward(model, d)
        out2d.copy_(o["flow_2d"], non_blocking=True)
        out3d.copy_(o["flow_3d"], non_blocking=True)
        torch.cuda.synchronize()
    if reference_extensions_available():
        with reference_extensions_bound():
            dt = run(step_ref, steps, 2)
        res["ref_cuda"] = {"value": B / dt, "ms_per_forward": 1e3 * dt,
                           "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in pin_ref.values()),
                           "flow_2d_max_abs_diff_vs_b200": float((out2d - flows_b[0]).abs().max()),
                           "flow_3d_max_abs_diff_vs_b200": float((out3d - flows_b[1]).abs().max())}
        res["speedup_vs_ref_cuda"] = res["b200"]["value"] / res["ref_cuda"]["value"]
    dt = run(step_ref, 2, 1)                              # wrapper.py's torch fallbacks on the GPU (no extension built)
    res["ref_torch_gpu"] = {"value": B / dt, "ms_per_forward": 1e3 * dt}
    res["note"] = ("whole forward incl. the convolutions / attention that are outside this library (cuDNN/cuBLAS in every arm, "
                   "TF32 convolutions = the reference's default); random-init weights; wall clock around H2D + forward + D2H")
    del model
    torch.cuda.empty_cache()
    return res


def main():
    args = parse()
    _claim_stdout()
    meta = baseline_meta()
    if args.batch is None:
        args.batch = load_workload_module().DEFAULT_BATCH[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, meta)
        return

    import torch.distributed as dist
    from rpeflow_b200 import _lib
    from rpeflow_b200.stack import (CONFIGS, CostVolumeStack, census_work, make_host_inputs, to_device)
    cfg = CONFIGS[args.workload]

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

    if args.scaling == "strong":                          # fixed global batch, contiguous slices
        lo, hi = rank * args.batch // world, (rank + 1) * args.batch // world
        B, first, global_batch = hi - lo, lo, args.batch
    else:
        B, first, global_batch = args.batch, rank * args.batch, world * args.batch
    host = make_host_inputs(cfg, B, first_sample=first, pin=True)
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
    value = global_batch / (ms_step * 1e-3)

    # -------- per-op device time (rank 0's numbers are reported) and the roofline entries built from it
    per_op = per_op_times(stack, x)
    work = census_work(cfg)
    peaks = measured_peaks()
    traffic_db = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")     # dram bytes per launch from the committed ncu --set full captures
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic_db = json.load(f)
    entries, families, serial_ms = roofline_report(cfg, work, per_op, B, peaks, traffic_db)
    hbm_entries = [e for e in entries if e["bound"] == "hbm"]
    top = max(hbm_entries, key=lambda e: e["ms"])              # the HBM-bound op with the largest share of the step
    flush = _Flush(dev)
    # that op's heaviest single call, timed alone with an L2 flush between launches (burst figure next to the in-step one)
    iso = None
    if top["op"].startswith("project_nn_corr"):
        from rpeflow_b200 import projection, ops as _ops
        from rpeflow_b200.workload import project_bytes
        lvl = int(top["op"][-1])
        h, w = cfg.level_hw(lvl)
        n = cfg.pyramid[lvl - 1]
        dec_2d, dec_3d = x["flowfeat"][lvl]
        xy = torch.rand(B, 2, n, device=dev) * torch.tensor([w - 1.0, h - 1.0], device=dev).view(1, 2, 1)
        nn = _ops.k_nearest_neighbor(xy, stack.pixel_grid(B, h, w), 1)[..., 0]
        ms = time_isolated(lambda: projection.project_feat_with_nn_corr(xy, dec_2d, dec_3d, nn), flush)
        nb = project_bytes(96, 64, n, h, w) * B
        iso = {"call": f"project_feat_with_nn_corr(C2=96, C3=64) at level {lvl}, batch {B} (3 launches)", "launch_ms": round(ms, 4),
               "algorithmic_bytes": nb, "achieved": round(nb / (ms * 1e-3) / 1e9, 1), "frac": round(nb / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)}
    elif top["op"] == "event_voxel":
        one = {k: v[:1] for k, v in x.items() if k in ("events", "ev_x", "ev_y", "ev_t", "ev_p")}
        ms = time_isolated(lambda: stack.voxelise(one), flush)
        nb = work["event_voxel_bytes"]
        iso = {"call": "event voxel grid of one sample (zero fill + scatter kernel), L2 flushed", "launch_ms": round(ms, 4),
               "algorithmic_bytes": nb, "achieved": round(nb / (ms * 1e-3) / 1e9, 1), "frac": round(nb / (ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)}
    roofline = {"kernel": f"{top['kernel']} ({top['op']}: the HBM-bound op with the largest share of the step, batch {B})",
                "bound": "hbm", "achieved": top["achieved"], "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": top["frac"],
                "traffic": top["traffic"], "peak_source": peaks["source"], "algorithmic_bytes_per_step": top["algorithmic_bytes"],
                "ms_per_step": top["ms"], "share_of_step": top["share_of_step"], "isolated": iso,
                "families": families, "per_op": entries, "serial_ms_all_ops": round(serial_ms, 4),
                "whole_step": {"algorithmic_bytes": (sum(work["corr2d_bytes"].values()) + work["gather_bytes"] + work["event_voxel_bytes"]) * B,
                               "graph_ms": ms_step}}
    roofline["whole_step"]["achieved"] = round(roofline["whole_step"]["algorithmic_bytes"] / (ms_step * 1e-3) / 1e9, 1)
    roofline["whole_step"]["frac"] = round(roofline["whole_step"]["achieved"] / peaks["hbm_gbs"], 4)

    ops_report = {"per_op_ms_per_step": {k: round(v["ms_per_step"], 4) for k, v in sorted(per_op.items())}}
    if rank == 0 and world == 1 and not args.no_ref_cuda:
        try:
            ops_report["vs_ref_cuda"] = ref_cuda_column(cfg, stack, x, B, per_op, flush)
        except Exception as e:
            ops_report["vs_ref_cuda"] = {"error": repr(e)}
    if rank == 0 and world == 1 and cfg.n_points <= 8192:
        try:
            ops_report["pointconv"] = pointconv_leg(cfg, x, B, dev)
        except Exception as e:
            ops_report["pointconv"] = {"error": repr(e)}

    # -------- end to end: every step copies ALL of its inputs from pinned host memory (copy stream, group by group
    # in dependency order) into one of two device input sets, replays the per-group CUDA graphs as the groups land,
    # and reads the step's checksum vector back to the host.  Step s+1's copies overlap step s's compute.
    e2e = None
    e2e_path = None
    if not args.no_e2e:
        if not args.eager:
            del graphed

        def e2e_leg(feeder):
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
            nres = e2e_run(max(2, args.warmup))
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
            del runners
            return {"value": global_batch * args.steps / float(te.item()), "unit": "frame-pairs/s",
                    "h2d_bytes_per_step": feeder.nbytes, "d2h_bytes_per_step": 8 * nres,
                    "h2d_gbs": feeder.nbytes * args.steps / float(te.item()) / 1e9, "wall_s": wall}

        feeder = HostFeeder(host, dev, depth=2)
        e2e = e2e_leg(feeder)
        e2e["pipeline"] = ("2 device input sets; copies of step s+1 (copy stream, 7 groups in dependency order) overlap the "
                           "graph replays of step s; result read back every step")
        e2e["inputs_copied"] = ("point clouds, raw events and every synthetic activation the ops read (activations the real model "
                                "produces on the device: this leg is bound by PCIe, see e2e_path_inputs and model_e2e)")
        del feeder
        # the same pipeline with only the path's EXTERNAL inputs crossing PCIe every step (point clouds + raw events); the
        # feature maps the real model computes on the device stay resident.  Reported beside `e2e`, not instead of it.
        try:
            feeder = HostFeeder(host, dev, depth=2, resident=x)
            e2e_path = e2e_leg(feeder)
            e2e_path["inputs_copied"] = "point clouds and raw events only; feature maps / point features resident on the device"
            del feeder
        except Exception as e:
            e2e_path = {"error": repr(e)}

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

    # -------- model level (rank 0, N=1): the unmodified reference model through the drop-in, host buffers in and out
    model_e2e = None
    if rank == 0 and world == 1 and not args.no_model and cfg.name == "things":
        del x, stack
        torch.cuda.empty_cache()
        try:
            model_e2e = model_leg(args, dev)
        except Exception as e:
            model_e2e = {"error": repr(e)}

    # -------- CPU baseline (rank 0, N=1 only): the reference's CPU path on a bounded sample, all threads + 1 thread
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        W = load_workload_module()
        threads = cpu_threads()
        torch.set_num_threads(threads)
        host1 = W.make_host_inputs(cfg, 1)
        R = make_cpu_ops(W, cfg)
        torch.set_num_threads(threads)
        cpu_frame_pair(cfg, host1, R)                  # warm-up
        n, t0 = 0, time.perf_counter()
        while True:
            cpu_frame_pair(cfg, host1, R)
            n += 1
            if time.perf_counter() - t0 > args.cpu_sample_s or n >= 8:
                break
        dtc = time.perf_counter() - t0
        cpu_baseline = {"value": n / dtc, "unit": "frame-pairs/s", "cores": threads, "kind": R.kind,
                        "sample": f"{n} frame pair(s) of the same workload through {R.what}, {dtc:.1f} s"}
        torch.set_num_threads(1)                        # the n = 1 figure BASELINE.md §3 asks for
        t0 = time.perf_counter()
        cpu_frame_pair(cfg, host1, R)
        dt1 = time.perf_counter() - t0
        torch.set_num_threads(threads)
        cpu_baseline["one_thread"] = {"value": 1.0 / dt1, "unit": "frame-pairs/s", "cores": 1, "sample": f"1 frame pair, {dt1:.1f} s"}

    if rank == 0:
        line = {
            "metric": meta["metric"], "value": value, "unit": "frame-pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(cfg, B, args, world),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "ops": ops_report, "model_e2e": model_e2e, "e2e_path_inputs": e2e_path,
        }
        if verify:
            line["verify"] = verify
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
