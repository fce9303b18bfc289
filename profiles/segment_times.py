"""Device time of each input-group graph segment of the stack (points / events / lvl5..lvl1), batch 32, things."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rpeflow_b200.stack import CONFIGS, CostVolumeStack, GraphedStack, make_host_inputs, to_device

dev = torch.device("cuda", 0)
cfg = CONFIGS["things"]
B = int(os.environ.get("BATCH", "32"))
x = to_device(make_host_inputs(cfg, B), dev)
stack = CostVolumeStack(cfg, dev)
gs = GraphedStack(stack, x, fused=False, with_checksum=False)
for _ in range(3):
    gs.replay()
torch.cuda.synchronize()
tot = {}
for rep in range(5):
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(gs.graphs) + 1)]
    evs[0].record()
    for i, (group, g) in enumerate(gs.graphs):
        g.replay()
        evs[i + 1].record()
    torch.cuda.synchronize()
    for i, (group, g) in enumerate(gs.graphs):
        tot[group] = min(tot.get(group, 1e9), evs[i].elapsed_time(evs[i + 1]))
print({k: round(v, 3) for k, v in tot.items()}, "sum", round(sum(tot.values()), 3))
