"""Per-kernel CUDA time of one train step (torch profiler), grouped by kernel name."""
import os, sys, json, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlagg_unet_b200.trainer import SyntheticPlan, nnUNetTrainer_MLAgg_2D_dt_MS
from torch.profiler import profile, ProfilerActivity
tr = nnUNetTrainer_MLAgg_2D_dt_MS(SyntheticPlan()).initialize()
batch = tr.synthetic_batch(device="cuda")
for _ in range(3):
    tr.train_step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tr.train_step(batch)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        agg[e.name][0] += 1
        agg[e.name][1] += e.device_time
tot = sum(v[1] for v in agg.values())
print(f"total kernel time {tot/1e3:.2f} ms, {sum(v[0] for v in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:45]:
    print(f"{t/1e3:8.3f} ms {n:5d}x  {k[:150]}")
