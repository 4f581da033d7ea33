"""bench.py -- BASELINE.json's metric: 2-D train images/s of nnUNetTrainer_MLAgg_2D_dt_MS on synthetic data of the
AbdomenMRI `2d_bs10` plan shape (10 x 1 x 320 x 320 per GPU, 14 classes), bf16 autocast, AdamW, grad-clip 12.

    python bench.py --gpus N --steps K --warmup W                (N > 1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" = one full train step (forward, Dice+CE deep-supervision loss, backward, clip, AdamW) -- the MLAgg + MSMM
hot path plus the cuDNN conv stages and the optimizer; nothing is skipped.  `value` is whole-job images/s with
the batch resident in HBM; `e2e` repeats the measurement through trainer.train_step with a PINNED HOST batch
(H2D copy of data + 5 targets and a D2H read of the loss every step).  `roofline` is the dominant kernel of the
named hot path (selective-scan backward): algorithmic bytes / CUDA-event time measured live in the timed region.
The timed steps replay the whole-step CUDA graph the trainer captures after three eager steps; the per-kernel events
(roofline, gpu_launches) come from one more pass over the same K steps run eagerly.
`cpu_baseline` / `--impl reference` time the reference's CPU path (oracle/: the reference module math restated and
pinned to the reference's own source via tests/golden; scan_ref.c with OpenMP) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "2D train images/s (nnUNetTrainer_MLAgg_2D_dt_MS, AbdomenMRI 2d_bs10 plan shape)"
WORKLOAD = "AbdomenMRI 2d_bs10 plan: 10x1x320x320 per GPU, 14 classes, full train step (fwd + DiceCE-DS loss + bwd + clip + AdamW)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def cpu_reference_run(steps, warmup, sample_batch, size, num_classes=14):
    """Reference CPU path: oracle full-network fwd + loss + analytic/autograd bwd + AdamW on `sample_batch` images."""
    from mlagg_unet_b200.trainer import DeepSupervisionDiceCE, SyntheticPlan, nnUNetTrainer_MLAgg_2D_dt_MS
    from oracle.network import mlla_uper_forward
    from oracle.scan import c_threads
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    plan = SyntheticPlan(patch_size=(size, size), batch_size=sample_batch, num_classes=num_classes)
    tr = nnUNetTrainer_MLAgg_2D_dt_MS(plan, device=torch.device("cpu"))
    net = tr.build_network_architecture(plan, {}, plan, 1, True)  # weights + conv stages; forward = oracle
    params = [p for p in net.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, 5e-4, weight_decay=3e-5, eps=1e-4)
    loss_fn = DeepSupervisionDiceCE(5, True, False)
    batch = tr.synthetic_batch(sample_batch)
    times = []
    for i in range(warmup + steps):
        t = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        l = loss_fn(mlla_uper_forward(net, batch["data"]), batch["target"])
        l.backward()
        torch.nn.utils.clip_grad_norm_(params, 12)
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t)
    sec = sum(times) / len(times)
    return {"value": sample_batch / sec, "unit": "images/s", "cores": max(torch.get_num_threads(), c_threads()),
            "kind": "port", "sec_per_step": sec,
            "sample": f"{steps} train steps (fwd+loss+bwd+clip+AdamW) of the whole network on {sample_batch} image(s) "
                      f"1x{size}x{size}, fp32, oracle hot path (scan_ref.c OpenMP) + torch conv stages"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=10)
    ap.add_argument("--size", type=int, default=320)
    ap.add_argument("--in-channels", type=int, default=1, help="3 for the Endovis17-shaped config 5 (3 x 512 x 512)")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: ONE global batch of --batch-per-gpu images split over the ranks exactly as the "
                         "reference does (nnUNetTrainer.py:295-307: [5, 5], [3, 3, 3, 1]; invalid at 8 ranks, SURVEY 8d)")
    ap.add_argument("--cpu-sample-batch", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-step", action="store_true",
                    help="one extra (eagerly launched) step between cudaProfilerStart/Stop, for "
                         "`ncu --profile-from-start off` launch lists (profiles/README.md)")
    ap.add_argument("--trace-step", default=None, metavar="CSV",
                    help="one extra eagerly launched step under torch.profiler (CUPTI activity records, no replay): "
                         "writes every kernel launch of the step with its duration, in launch order")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    a.warmup = max(a.warmup, 3)

    if a.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(a.steps, min(a.warmup, 1), a.cpu_sample_batch, a.size)
        print(json.dumps({
            "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": min(a.warmup, 1), "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "per_step_sample_images": a.cpu_sample_batch},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    from mlagg_unet_b200 import _lib
    from mlagg_unet_b200.trainer import SyntheticPlan, nnUNetTrainer_MLAgg_2D_dt_MS
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product path has no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(1234 + rank)

    cpu_base = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_reference_run(3, 1, a.cpu_sample_batch, a.size)
        cpu_base = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    global_batch = a.batch_per_gpu * world
    if a.strong:
        from mlagg_unet_b200.trainer import split_batch
        sizes = split_batch(a.batch_per_gpu, world)
        assert all(v > 0 for v in sizes), f"the reference's batch split {sizes} is invalid for {world} ranks (SURVEY.md 8d)"
        global_batch, a.batch_per_gpu = a.batch_per_gpu, sizes[rank]
    plan = SyntheticPlan(patch_size=(a.size, a.size), batch_size=a.batch_per_gpu, num_input_channels=a.in_channels)
    tr = nnUNetTrainer_MLAgg_2D_dt_MS(plan, device=dev).initialize()
    host = tr.synthetic_batch(seed=rank, pin=True)
    resident = {"data": host["data"].to(dev), "target": [t.to(dev) for t in host["target"]]}
    h2d = host["data"].numel() * host["data"].element_size() + sum(t.numel() * t.element_size() for t in host["target"])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(batch, sync_loss, steps, events=False):
        barrier()
        _lib.STATS["launches"] = 0
        _lib.STATS["events"] = {} if events else None
        _lib.STATS["bytes"] = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = tr.train_step(batch, sync=sync_loss)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ev, _lib.STATS["events"] = _lib.STATS["events"], None
        return float(ms.item()), _lib.STATS["launches"], ev, out

    # warm-up: >= 3 eager steps, then train_step captures the whole step in a CUDA graph (trainer._capture) and replays it
    for _ in range(a.warmup + 2):
        tr.train_step(resident, sync=False)
    graphed = tr._graph is not None
    sampler = ClockSampler(local).start() if rank == 0 else None
    ms, _, _, _ = timed(resident, False, a.steps)
    clocks = sampler.stop() if sampler else None
    for _ in range(2):
        tr.train_step(host, sync=True)
    ms_e2e, _, _, out = timed(host, True, a.steps)
    if a.profile_step:
        barrier()
        _lib.STATS["events"] = {}          # eager launches (ncu cannot replay some graph kernel nodes); same kernels
        torch.cuda.profiler.start()
        tr.train_step(resident, sync=False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        _lib.STATS["events"] = None
    if a.trace_step and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        barrier()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            tr.train_step(resident, sync=False)
            torch.cuda.synchronize()
        evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA
                      and not e.name.lower().startswith(("memcpy", "memset"))), key=lambda e: e.time_range.start)
        with open(a.trace_step, "w") as f:
            f.write('"ID","Kernel Name","Metric Unit","Metric Value"\n')
            for i, e in enumerate(evs):
                f.write('"%d","%s","us","%.3f"\n' % (i, e.name.replace('"', "'"), e.device_time))
    # per-kernel CUDA events need Python between the launches: the same K steps once more, eagerly (not part of `value`);
    # the graph replays exactly this launch sequence, so the launch count is taken here too
    ms_eager, launches, ev, _ = timed(resident, False, a.steps, events=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    imgs = global_batch * a.steps
    # roofline of the dominant hot-path kernel: selective-scan backward at the mamba interface
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    net = tr.network.module if hasattr(tr.network, "module") else tr.network
    hw = [(a.size // 2 // 2 ** i, a.size // 2 // 2 ** i) for i in range(4)]
    Lcat = sum(h * w for h, w in hw)
    D, G, N = 4 * net.mambaskip.blocks[0].self_attention.d_inner, 4, 16
    alg = {"scan_fwd": 4 * (3 * a.batch_per_gpu * D * Lcat + 2 * a.batch_per_gpu * G * N * Lcat) + 4 * D * (N + 2),
           "scan_bwd": 4 * (5 * a.batch_per_gpu * D * Lcat + 4 * a.batch_per_gpu * G * N * Lcat)}
    kern = {}
    for name, pairs in (ev or {}).items():
        ts = [p[0].elapsed_time(p[1]) for p in pairs]
        kern[name] = {"launches": len(ts), "ms_avg": sum(ts) / len(ts), "ms_total": sum(ts)}
    fam_bytes = dict(_lib.STATS.get("bytes") or {})       # algorithmic bytes per family over the eager K steps
    dom = "scan_bwd"
    roof = None
    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel (the newest committed
    # capture whose shape AND kernel revision match; null otherwise -- a stale figure is worse than none)
    traffic, traffic_src = None, None
    for name in ("scan_bwd_r02_ncu.json", "scan_bwd_r01_ncu.json"):
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", name)))
            if prof["shape"] == {"batch": a.batch_per_gpu, "dim": D, "groups": G, "dstate": N, "L": Lcat}:
                traffic, traffic_src = prof["dram_bytes_read"] + prof["dram_bytes_write"], "profiles/" + name
                break
        except Exception:
            continue
    if dom in kern:
        ach = alg[dom] / (kern[dom]["ms_avg"] * 1e-3) / 1e9
        roof = {"kernel": "mlagg::scan_bwd_kernel (selective-scan backward in the fused MSMM operand mode, fp32 I/O; algorithmic bytes counted at the mamba interface, SURVEY.md 8d)", "bound": "hbm",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": alg[dom], "ms_per_launch": kern[dom]["ms_avg"],
                "traffic": traffic, "traffic_source": traffic_src,
                # the scan is NOT HBM-bound on B200: one MUFU.EX2 per (channel, state, step) at 16 lanes/clk/SM (measured,
                # profiles/mufu_bench_r01.txt) puts the exponential floor above the HBM floor -- report both
                "xu_floor": {"exp_per_launch": a.batch_per_gpu * D * N * Lcat * (2 if dom == "scan_bwd" else 1),
                             "floor_ms": a.batch_per_gpu * D * N * Lcat * (2 if dom == "scan_bwd" else 1)
                                         / (16.0 * 148 * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6) * 1e3,
                             "note": "backward recomputes a_t from the 16-step checkpoints: 1 exponential per update "
                                     "in the recompute sweep, reused by the adjoint sweep; counted 2x with the "
                                     "softplus / sigmoid helpers"},
                # every kernel family of the library: ms per step, and -- where the wrapper states the algorithmic bytes of
                # its calls -- achieved GB/s over all of a step's calls and its fraction of the same HBM peak
                "also": {k: {"ms_avg": v["ms_avg"], "launches_per_step": v["launches"] / a.steps,
                             "ms_per_step": v["ms_total"] / a.steps,
                             **({"achieved_GBps": alg[k] / (v["ms_avg"] * 1e-3) / 1e9,
                                 "frac": alg[k] / (v["ms_avg"] * 1e-3) / 1e9 / peak} if k in alg else {}),
                             **({"achieved_GBps": fam_bytes[k] / (v["ms_total"] * 1e-3) / 1e9,
                                 "frac": fam_bytes[k] / (v["ms_total"] * 1e-3) / 1e9 / peak,
                                 "algorithmic_bytes_per_step": fam_bytes[k] / a.steps}
                                if (k not in alg and fam_bytes.get(k)) else {})}
                         for k, v in kern.items()}}
    line = {
        "metric": METRIC, "value": imgs / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong" if a.strong else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD if (a.size, a.in_channels, a.strong) == (320, 1, False) else
                   f"{'strong-scaling split of one global batch, ' if a.strong else ''}{a.in_channels}x{a.size}x{a.size} "
                   f"inputs, 14 classes, full train step (fwd + DiceCE-DS loss + bwd + clip + AdamW)",
                   "global_batch": global_batch, "per_gpu_batch": a.batch_per_gpu,
                   "peak_memory_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
                   "parallelism": f"dp{world}" if world > 1 else "single", "scan_state_dtype": "f32",
                   "cuda_graph": graphed, "eager_ms_per_step_with_kernel_events": ms_eager / a.steps,
                   "l2": "per-step working set (~14 GB of activations) >> 126 MB L2, no flush needed"},
        "clocks": clocks,
        "e2e": {"value": imgs / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / a.steps, "loss": float(out["loss"])},
        "gpu_launches": launches,
        "roofline": roof,
        "cpu_baseline": cpu_base,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
