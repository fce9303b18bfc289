"""torch.profiler kernel table of one RPEFlow.forward (B=4, 960x540, 8192 points): arm B (install()) and arm C (the
reference's own CUDA extensions).  Shows how much of the forward the hot path is and what is left around it."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from torch.profiler import profile, ProfilerActivity
from refmodel_util import reference_root, reference_extensions_bound
from rpeflow_b200 import refhost
import rpeflow_b200.install as inst

dev = torch.device("cuda", 0)
model = refhost.build_rpeflow(reference_root(), device=dev, install=False, seed=0)
host = refhost.synthetic_model_inputs(4, 540, 960, 8192, seed=3)
inputs = {k: v.to(dev) for k, v in host.items()}

def table(tag):
    for _ in range(2):
        refhost.forward(model, inputs)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        refhost.forward(model, inputs)
        torch.cuda.synchronize()
    ev = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
    ev.sort(key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in ev)
    lines = [f"## {tag}: {tot/1e3:.2f} ms of GPU kernel time, {sum(e.count for e in ev)} launches", "", "| kernel | calls | ms | share |", "|---|---|---|---|"]
    for e in ev[:45]:
        lines.append(f"| {e.key[:90]} | {e.count} | {e.device_time_total/1e3:.3f} | {100*e.device_time_total/tot:.1f}% |")
    ours = sum(e.device_time_total for e in ev if e.key.startswith("b200::") or "b200::" in e.key)
    lines.append(f"\nb200:: kernels: {ours/1e3:.2f} ms ({100*ours/tot:.1f}%)")
    import time
    t0 = time.perf_counter()
    for _ in range(3):
        refhost.forward(model, inputs)
    torch.cuda.synchronize()
    lines.append(f"wall per forward: {(time.perf_counter()-t0)/3*1e3:.1f} ms")
    return "\n".join(lines)

out = []
inst.install()
out.append(table("arm B (install(): this library)"))
inst.uninstall()
with reference_extensions_bound():
    out.append(table("arm C (the reference's own CUDA extensions)"))
open(os.path.join(ROOT, "gpurun_out", "r2_model_profile.md"), "w").write("\n\n".join(out) + "\n")
print("\n\n".join(out))
