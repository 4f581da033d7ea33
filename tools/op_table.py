"""Per-aten-op CUDA time of one train step grouped by input shapes (torch profiler) -- finds the glue copies/reductions."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlagg_unet_b200.trainer import SyntheticPlan, nnUNetTrainer_MLAgg_2D_dt_MS
from torch.profiler import profile, ProfilerActivity
tr = nnUNetTrainer_MLAgg_2D_dt_MS(SyntheticPlan()).initialize()
batch = tr.synthetic_batch(device="cuda")
for _ in range(3):
    tr.train_step(batch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    tr.train_step(batch)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    t = getattr(e, "self_device_time_total", 0) or getattr(e, "self_cuda_time_total", 0)
    if t > 0:
        rows.append((t, e.count, e.key, str(e.input_shapes)[:150]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total self device time {tot/1e3:.2f} ms")
for t, n, k, sh in rows[:70]:
    print(f"{t/1e3:8.3f} ms {n:4d}x {k[:44]:44s} {sh}")
