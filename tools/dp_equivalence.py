"""Multi-GPU data plane == single process: one data-parallel run on [5, 5] (the reference's split of
the 2d_bs10 batch over 2 ranks, nnUNetTrainer.py:295-307) against a single-process run on the 10 concatenated images,
through the product's own path -- flat fp32 gradient all-reduce, packed batch-dice all-gather, two-graph replay
(trainer.py `_reduce_clip_step`, `_capture`, `_replay`).  Compares the loss of every step and all parameters after the
last one.  DropPath is switched off (its per-sample masks come from per-process RNG streams).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/dp_equivalence.py [--size 320] [--steps 6] [--out gpurun_out/dp_equivalence.json]
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mlagg_unet_b200.thirdparty_shims import DropPath  # noqa: E402
from mlagg_unet_b200.trainer import SyntheticPlan, nnUNetTrainer_MLAgg_2D_dt_MS, split_batch  # noqa: E402


def make(plan, dev, ddp):
    torch.manual_seed(1234)
    tr = nnUNetTrainer_MLAgg_2D_dt_MS(plan, device=dev)
    tr.is_ddp = ddp
    tr.initialize()
    for m in tr.network.modules():
        if isinstance(m, DropPath):
            m.drop_prob = 0.0
    return tr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=320)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--global-batch", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "dp_equivalence.json"))
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sizes = split_batch(a.global_batch, world)
    assert all(v > 0 for v in sizes), sizes
    lo = sum(sizes[:rank])

    gplan = SyntheticPlan(patch_size=(a.size, a.size), batch_size=a.global_batch)
    full = nnUNetTrainer_MLAgg_2D_dt_MS(gplan, device=torch.device("cpu")).synthetic_batch(a.global_batch, seed=7)
    mine = {"data": full["data"][lo:lo + sizes[rank]].to(dev), "target": [t[lo:lo + sizes[rank]].to(dev) for t in full["target"]]}

    tr = make(SyntheticPlan(patch_size=(a.size, a.size), batch_size=sizes[rank]), dev, True)
    dp_losses = []
    for _ in range(a.steps):
        l = torch.as_tensor(float(tr.train_step(mine)["loss"]), device=dev)
        # the reference logs the mean over ranks of the per-rank losses (on_train_epoch_end, :866-876)
        dist.all_reduce(l)
        dp_losses.append(float(l) / world)
    graphs = 0 if tr._graph is None else len(tr._graph)
    dp_params = {n: p.detach().clone() for n, p in tr.network.named_parameters()}
    # every rank holds the same parameters
    worst_rank_diff = 0.0
    for n, p in dp_params.items():
        q = p.clone()
        dist.broadcast(q, 0)
        worst_rank_diff = max(worst_rank_diff, float((p - q).abs().max()))
    del tr
    torch.cuda.empty_cache()
    dist.barrier()

    res = None
    if rank == 0:
        one = make(gplan, dev, False)
        whole = {"data": full["data"].to(dev), "target": [t.to(dev) for t in full["target"]]}
        sp_losses = [float(one.train_step(whole)["loss"]) for _ in range(a.steps)]
        worst, worst_name, upd = 0.0, None, 0.0
        init = make(gplan, dev, False)
        p0 = dict(init.network.named_parameters())
        num = den = 0.0
        for n, p in one.network.named_parameters():
            d = float((p.detach() - dp_params[n]).abs().max())
            upd = max(upd, float((p.detach() - p0[n].detach()).abs().max()))
            num += float((p.detach() - dp_params[n]).double().pow(2).sum())
            den += float((p.detach() - p0[n].detach()).double().pow(2).sum())
            if d > worst:
                worst, worst_name = d, n
        res = {"world": world, "split": sizes, "size": a.size, "steps": a.steps, "graphs_per_step_dp": graphs,
               "single_process_graph": one._graph is not None,
               "loss_dp": dp_losses, "loss_single": sp_losses,
               "max_abs_loss_diff": max(abs(x - y) for x, y in zip(dp_losses, sp_losses)),
               "max_abs_param_diff": worst, "param_with_max_diff": worst_name, "max_abs_param_update": upd,
               # AdamW's update is ~ lr * sign(g) for |g| >> eps: an element whose gradient is ~0 can step the other way
               # when the bf16 conv-stage gradients are rounded per rank, so the element-wise maximum is one lr step;
               # the meaningful figure is the distance between the two UPDATE VECTORS relative to their length
               "rel_l2_update_diff": (num / max(den, 1e-30)) ** 0.5,
               "max_abs_diff_between_ranks": worst_rank_diff}
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        json.dump(res, open(a.out, "w"), indent=1)
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        # bf16 compute with fp32 atomics: the two runs sum the same per-image gradients in a different order
        assert res["max_abs_diff_between_ranks"] == 0.0, res
        assert res["max_abs_loss_diff"] < 2e-3, res
        assert res["max_abs_param_diff"] <= 2.5 * res["max_abs_param_update"] / res["steps"], res   # <= ~2 lr steps
        assert res["rel_l2_update_diff"] < 0.05, res


if __name__ == "__main__":
    main()
