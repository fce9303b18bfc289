"""Host-side (Python) time of one RPEFlow.forward with install(): cProfile, cumulative time of this library's wrappers vs the
whole forward (B=4, 960x540, 8192 points).  The forward is launch-bound, so wrapper overhead is wall-clock time."""
import cProfile, os, pstats, sys, time, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from refmodel_util import reference_root
from rpeflow_b200 import refhost
import rpeflow_b200.install as inst

dev = torch.device("cuda", 0)
model = refhost.build_rpeflow(reference_root(), device=dev, install=False, seed=0)
host = refhost.synthetic_model_inputs(4, 540, 960, 8192, seed=3)
inputs = {k: v.to(dev) for k, v in host.items()}
inst.install()
for _ in range(3):
    refhost.forward(model, inputs)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    refhost.forward(model, inputs)
torch.cuda.synchronize()
print(f"wall per forward: {(time.perf_counter() - t0) / 5 * 1e3:.1f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    refhost.forward(model, inputs)
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
st = pstats.Stats(pr, stream=s).sort_stats("cumulative")
st.print_stats(60)
txt = s.getvalue()
print("\n".join(l for l in txt.splitlines() if ("rpeflow_b200" in l or "forward" in l or "ncalls" in l or "function calls" in l))[:6000])
