"""Every C-ABI call of ONE eagerly launched train step (config 3 by default) with its integer arguments and its device
time: each `mlagg_*` entry point of the loaded library is wrapped by a recorder that brackets the call with CUDA events
(synchronising before and after, so the number is the call alone, caches as the step leaves them).  Calls are grouped by
(entry point, integer / float arguments); the table is what the per-kernel bandwidth figures in DESIGN.md are computed
from (the CUPTI launch list has durations but no shapes).

    python tools/call_shapes.py [--size 320] [--batch 10] [--in-channels 1] [--top 80] [--json out.json]
"""
import argparse
import ctypes
import json
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mlagg_unet_b200 import _lib  # noqa: E402
from mlagg_unet_b200.trainer import SyntheticPlan, nnUNetTrainer_MLAgg_2D_dt_MS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=320)
    ap.add_argument("--batch", type=int, default=10)
    ap.add_argument("--in-channels", type=int, default=1)
    ap.add_argument("--top", type=int, default=80)
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    os.environ["MLAGG_CUDA_GRAPH"] = "0"
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1234)
    plan = SyntheticPlan(patch_size=(a.size, a.size), batch_size=a.batch, num_input_channels=a.in_channels)
    tr = nnUNetTrainer_MLAgg_2D_dt_MS(plan, device=dev).initialize()
    batch = tr.synthetic_batch(seed=0, pin=False)
    batch = {"data": batch["data"].to(dev), "target": [t.to(dev) for t in batch["target"]]}
    for _ in range(3):
        tr.train_step(batch, sync=False)
    torch.cuda.synchronize()

    L = _lib.lib()
    log = defaultdict(list)
    record = {"on": False}

    def wrap(name, fn, argtypes):
        scalar = [i for i, t in enumerate(argtypes) if t not in (ctypes.c_void_p,)]

        def call(*args):
            if not record["on"]:
                return fn(*args)
            key = (name,) + tuple(args[i] if i < len(args) else None for i in scalar)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            torch.cuda.synchronize()
            log[key].append(e0.elapsed_time(e1) * 1e3)
            return rc
        return call

    for name, (res, args) in _lib.SIGNATURES.items():
        if res is ctypes.c_int and args and name not in ("mlagg_version",):
            setattr(L, name, wrap(name, getattr(L, name), args))
    record["on"] = True
    tr.train_step(batch, sync=False)
    torch.cuda.synchronize()
    record["on"] = False

    rows = sorted(((sum(v), len(v), k) for k, v in log.items()), reverse=True)
    total = sum(r[0] for r in rows)
    print(f"{sum(r[1] for r in rows)} C-ABI calls, {total / 1e3:.2f} ms of device time (each call timed alone)")
    byname = defaultdict(float)
    for s, n, k in rows:
        byname[k[0]] += s
    for nm, s in sorted(byname.items(), key=lambda kv: -kv[1]):
        print(f"  {s / 1e3:8.3f} ms  {nm}")
    print()
    for s, n, k in rows[:a.top]:
        print(f"{s:9.1f} us {n:3d}x {s / n:8.1f} us/call  {k[0]} {k[1:]}")
    if a.json:
        with open(a.json, "w") as f:
            json.dump([{"name": k[0], "args": list(k[1:]), "calls": n, "us_total": s} for s, n, k in rows], f, indent=0)


if __name__ == "__main__":
    main()
