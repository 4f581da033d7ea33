"""CPU, world_size 2, gloo: the N > 1 host logic -- batch-dice statistics gathered across ranks with
_AllGatherGrad (reference training/loss/dice.py:104-107, utilities/ddp_allgather.py:25-49) give the same loss and
gradients as one process on the concatenated batch; DDP over a module that carries the reference's unused
`dummy_tensor` (frozen here) survives more than one iteration; deep-supervision flag reaches `.module`."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Tiny(torch.nn.Module):
    """stand-in with the two properties that matter for DDP: used conv weights and an unused frozen parameter"""

    def __init__(self, ncls=4):
        super().__init__()
        self.conv = torch.nn.Conv2d(1, ncls, 3, padding=1)
        self.dummy_tensor = torch.nn.Parameter(torch.tensor([1.0]), requires_grad=False)
        self.deep_supervision = True

    def forward(self, x):
        y = self.conv(x)
        return [y, torch.nn.functional.avg_pool2d(y, 2)] if self.deep_supervision else y


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mlagg_unet_b200.trainer import DeepSupervisionDiceCE, nnUNetTrainer_MLAgg_2D_dt_MS
    torch.manual_seed(0)
    net = _Tiny()
    ddp = torch.nn.parallel.DistributedDataParallel(net)
    g = torch.Generator().manual_seed(7)
    data = torch.randn(4, 1, 8, 8, generator=g)
    tgt = [torch.randint(0, 4, (4, 1, 8, 8), generator=g).float(), torch.randint(0, 4, (4, 1, 4, 4), generator=g).float()]
    lo, hi = rank * 2, rank * 2 + 2
    loss_fn = DeepSupervisionDiceCE(2, batch_dice=True, ddp=True)
    losses = []
    for it in range(3):  # > 1 iteration: the reference's requires_grad dummy_tensor breaks DDP on the second one
        ddp.zero_grad()
        l = loss_fn(ddp(data[lo:hi]), [t[lo:hi] for t in tgt])
        l.backward()
        losses.append(float(l.detach()))
    grad = net.conv.weight.grad.clone()
    tr = nnUNetTrainer_MLAgg_2D_dt_MS.__new__(nnUNetTrainer_MLAgg_2D_dt_MS)
    tr.network = ddp
    tr.set_deep_supervision_enabled(False)
    flag_ok = net.deep_supervision is False and not isinstance(ddp(data[lo:hi]), list)
    if rank == 0:
        q.put((losses, grad.tolist(), flag_ok))  # plain lists: tensors would travel by fd and outlive the worker
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_ddp_batch_dice_matches_single_process():
    from mlagg_unet_b200.trainer import DeepSupervisionDiceCE
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    losses, grad, flag_ok = q.get(timeout=100)
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    assert flag_ok
    # single process, whole batch: dice statistics are batch-wide, CE is a per-rank mean -> same value for equal splits
    torch.manual_seed(0)
    net = _Tiny()
    g = torch.Generator().manual_seed(7)
    data = torch.randn(4, 1, 8, 8, generator=g)
    tgt = [torch.randint(0, 4, (4, 1, 8, 8), generator=g).float(), torch.randint(0, 4, (4, 1, 4, 4), generator=g).float()]
    loss_fn = DeepSupervisionDiceCE(2, batch_dice=True, ddp=False)
    out = net(data)
    # DDP averages gradients over ranks and each rank's loss holds the GLOBAL dice + its LOCAL CE
    dice_ce = [loss_fn.one(o, t) for o, t in zip(out, tgt)]
    ref = sum(w * l for w, l in zip(loss_fn.weights, dice_ce))
    lo_ce = [torch.nn.functional.cross_entropy(o[:2], t[:2, 0].long()) for o, t in zip(out, tgt)]
    all_ce = [torch.nn.functional.cross_entropy(o, t[:, 0].long()) for o, t in zip(out, tgt)]
    rank0 = ref + sum(w * (a - b) for w, a, b in zip(loss_fn.weights, lo_ce, all_ce))
    assert losses[0] == pytest.approx(float(rank0), rel=1e-5)
    assert losses[0] == losses[1] == losses[2]  # no optimizer step: identical iterations, and no DDP error
    ref.backward()
    assert torch.allclose(torch.tensor(grad), net.conv.weight.grad, rtol=1e-4, atol=1e-6)
