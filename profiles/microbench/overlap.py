"""Does the point-independent work (voxel grid, correlation2d) really run under the FPS latency chain?  Times FPS, the
voxeliser and the five correlation2d alone and together on two streams (batch = bench default)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from rpeflow_b200 import ops
from rpeflow_b200.stack import CONFIGS, CostVolumeStack, make_host_inputs, to_device

dev = torch.device("cuda", 0)
cfg = CONFIGS["things"]
B = int(os.environ.get("BATCH", "74"))
x = to_device(make_host_inputs(cfg, B), dev)
stack = CostVolumeStack(cfg, dev)
pc = torch.cat([x["pcs"][:, :3], x["pcs"][:, 3:]], 0).transpose(1, 2).contiguous()
side = torch.cuda.Stream()


def fps():
    return ops.furthest_point_sampling(pc, 4096)


def vox():
    return stack.voxelise(x)


def corr():
    return [ops.correlation2d(*x["feat2d"][l], 4) for l in range(5, 0, -1)]


def timed(main_fn, side_fns):
    """Both streams captured into one CUDA graph (two parallel branches), so host launch order plays no part."""
    def body():
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for f in side_fns:
                f()
        if main_fn:
            main_fn()
        main.wait_stream(side)
    cap = torch.cuda.Stream()
    with torch.cuda.stream(cap):
        body()                                            # warm-up (allocator, function attributes)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cap):
            body()
    best = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for name, m, s in (("fps", fps, []), ("vox", None, [vox]), ("corr2d x5", None, [corr]), ("vox+corr", None, [vox, corr]),
                   ("fps || vox", fps, [vox]), ("fps || corr", fps, [corr]), ("fps || vox+corr", fps, [vox, corr])):
    print(f"{name:18s} {timed(m, s):.3f} ms")
