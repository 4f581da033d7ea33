"""Times one synthetic train step of the full network (BASELINE config 3 shape) and prints a kernel-time table."""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlagg_unet_b200.trainer import SyntheticPlan, nnUNetTrainer_MLAgg_2D_dt_MS

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=10)
ap.add_argument("--size", type=int, default=320)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--profile", type=int, default=1)
a = ap.parse_args()
tr = nnUNetTrainer_MLAgg_2D_dt_MS(SyntheticPlan(patch_size=(a.size, a.size), batch_size=a.B)).initialize()
batch = tr.synthetic_batch(device="cuda")
for _ in range(3):
    out = tr.train_step(batch)
torch.cuda.synchronize()
t = time.time()
for _ in range(a.steps):
    out = tr.train_step(batch)
torch.cuda.synchronize()
dt = (time.time() - t) / a.steps
print(json.dumps({"ms_per_step": dt * 1e3, "images_per_s": a.B / dt, "loss": float(out["loss"]),
                  "max_mem_GB": torch.cuda.max_memory_allocated() / 2**30}))
if a.profile:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        tr.train_step(batch)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=60))
