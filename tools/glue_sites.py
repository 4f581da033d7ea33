"""Where does the remaining torch glue of one train step come from?  One eagerly launched step under torch.profiler;
every aten op that launches copy / add / fill / cat kernels is listed by (op, input shapes, innermost frame inside this
repo) with its summed device time.  Backward ops have no Python stack: they are listed under the autograd node that ran
them.

    python tools/glue_sites.py [--top 60]
"""
import argparse
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mlagg_unet_b200.trainer import SyntheticPlan, nnUNetTrainer_MLAgg_2D_dt_MS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--top", type=int, default=60)
    a = ap.parse_args()
    os.environ["MLAGG_CUDA_GRAPH"] = "0"
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1234)
    tr = nnUNetTrainer_MLAgg_2D_dt_MS(SyntheticPlan(patch_size=(320, 320), batch_size=10), device=dev).initialize()
    batch = tr.synthetic_batch(seed=0, pin=False)
    batch = {"data": batch["data"].to(dev), "target": [t.to(dev) for t in batch["target"]]}
    for _ in range(3):
        tr.train_step(batch, sync=False)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA],
                                record_shapes=True, with_stack=True) as prof:
        tr.train_step(batch, sync=False)
        torch.cuda.synchronize()
    want = ("aten::copy_", "aten::add", "aten::add_", "aten::cat", "aten::fill_", "aten::zero_", "aten::mul", "aten::sum",
            "aten::contiguous", "aten::clone", "aten::_to_copy", "aten::index_select", "aten::slice_backward",
            "aten::select_backward", "aten::constant_pad_nd", "aten::div", "aten::sub", "aten::neg", "aten::mul_")
    rows = defaultdict(lambda: [0.0, 0])
    for ev in prof.events():
        if ev.name not in want or ev.device_time_total <= 0:
            continue
        # only leaves: skip ops whose children carry the same kernels
        if any(c.name in want and c.device_time_total > 0 for c in ev.cpu_children):
            continue
        site = "-"
        for fr in ev.stack or []:
            if "mlagg-unet_b200" in fr or "mlagg_unet_b200" in fr or "/bench.py" in fr:
                site = fr.split("/")[-1]
                break
        par = ev.cpu_parent
        chain = []
        while par is not None and len(chain) < 3:
            chain.append(par.name[:48])
            par = par.cpu_parent
        key = (ev.name, str(ev.input_shapes)[:70], site, " < ".join(chain))
        rows[key][0] += ev.device_time_total
        rows[key][1] += 1
    tot = sum(v[0] for v in rows.values())
    print(f"{tot / 1e3:.3f} ms of device time in {sum(v[1] for v in rows.values())} glue ops")
    for (name, shp, site, chain), (t, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print(f"{t:8.1f} us {n:3d}x  {name:18s} {shp:70s} {site:44s} {chain}")


if __name__ == "__main__":
    main()
