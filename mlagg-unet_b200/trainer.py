"""Trainer-level boundary: `nnUNetTrainer_MLAgg_2D_dt_MS` with the reference's static
`build_network_architecture(plans_manager, dataset_json, configuration_manager, num_input_channels,
enable_deep_supervision)` (reference nnUNetTrainer_MLAgg_2D_dt_MS.py:62-92), `configure_optimizers` (:137-147),
`_get_deep_supervision_scales` (:101-104), `_build_loss` (:106-129), `set_deep_supervision_enabled` (:94-99) and
a `train_step` that mirrors nnUNetTrainer.train_step (nnUNetTrainer.py:833-863).

nnunetv2 itself cannot be imported in this image, so this class stands alone: the plans /
configuration / label managers are duck-typed (`.patch_size`, `.batch_dice`, `.get_label_manager(...)`), and a
synthetic stand-in (`SyntheticPlan`) provides them for benchmarks.  With a real nnunetv2 install, the overlay
`nnunetv2/training/nnUNetTrainer/nnUNetTrainer_MLAgg_2D_dt_MS.py` subclasses the real
nnUNetTrainer and delegates here.

Deviations from the reference, all from SURVEY.md F5/F6: bf16 autocast without GradScaler (fp16+GradScaler is
the reference default; selectable), `dummy_tensor` frozen, deep-supervision flag set on `.module` under DDP.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import os

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from .mlagg import MLLA_Uper
from .thirdparty_shims import CosineLRScheduler


# ------------------------------------------------------------------ loss (reference training/loss/*, off the hot path)
class _AllGatherGrad(torch.autograd.Function):
    """all_gather forward, all_reduce backward (reference utilities/ddp_allgather.py:25-49)."""

    @staticmethod
    def forward(ctx, t):
        out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t.contiguous())
        return torch.stack(out, dim=0)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        return g[dist.get_rank()]


def soft_dice_loss(logits, target, batch_dice=True, do_bg=False, smooth=1e-5, ddp=False):
    """MemoryEfficientSoftDiceLoss with softmax non-linearity (reference training/loss/dice.py:58-112)."""
    x = torch.softmax(logits.float(), dim=1)
    with torch.no_grad():
        onehot = torch.zeros_like(x, dtype=torch.bool).scatter_(1, target.long(), 1)
    if not do_bg:
        x, onehot = x[:, 1:], onehot[:, 1:]
    axes = tuple(range(2, x.dim()))
    inter, spred, sgt = (x * onehot).sum(axes), x.sum(axes), onehot.sum(axes).float()
    if ddp and batch_dice:
        # the reference gathers the three statistics separately (dice.py:104-107); one packed collective per scale
        # gives the same sums with a third of the launches / stream hand-offs
        inter, spred, sgt = _AllGatherGrad.apply(torch.stack((inter, spred, sgt), dim=0)).sum(0).unbind(0)
    if batch_dice:
        inter, spred, sgt = inter.sum(0), spred.sum(0), sgt.sum(0)
    return -((2 * inter + smooth) / torch.clip(sgt + spred + smooth, 1e-8)).mean()


class _DiceCEStats(torch.autograd.Function):
    """C ABI: mlagg_dice_ce_stats_fwd / _bwd (csrc/loss.cu).  (logits (B, K, H, W), target (B, 1, H, W)) ->
    (stats (B, K, 3) = per-image, per-class (sum p [t = k], sum p, sum [t = k]) with p = softmax(logits), ce_sum (1))."""

    @staticmethod
    def addressable(logits, target):
        if not (logits.is_cuda and logits.dim() == 4 and logits.dtype in (torch.float32, torch.bfloat16)
                and target.dtype in (torch.float32, torch.int64) and logits.shape[1] <= 32 and target.shape[1] == 1):
            return False
        W = logits.shape[3]
        return logits.stride(2) == W * logits.stride(3) and logits.stride(3) >= 1 and logits.stride(1) >= 1

    @staticmethod
    def forward(ctx, logits, target):
        from . import _lib
        Bn, K, H, W = logits.shape
        tgt = target.contiguous()
        stats, ce = _lib.zeros((Bn, K, 3), logits.device), _lib.zeros(1, logits.device)
        dt, tdt = (0 if logits.dtype == torch.float32 else 1), (0 if tgt.dtype == torch.float32 else 1)
        # heads padded to 16 channels (mlagg._conv1x1_cl): every pixel row owns 16 contiguous bf16 -> vector row access
        rows16 = (dt == 1 and K <= 16 and logits.stride() == (H * W * 16, 1, W * 16, 16) and logits.data_ptr() % 32 == 0
                  and logits.untyped_storage().nbytes() >= (logits.storage_offset() + Bn * H * W * 16) * 2)
        dt |= 2 if rows16 else 0
        with torch.cuda.device(logits.device), _lib.timed("dice_ce_fwd"):
            rc = _lib.lib().mlagg_dice_ce_stats_fwd(logits.data_ptr(), tgt.data_ptr(), stats.data_ptr(), ce.data_ptr(), Bn,
                                                    H * W, K, logits.stride(0), logits.stride(1), logits.stride(3), dt, tdt,
                                                    _lib.stream_ptr())
        _lib.check(rc, "mlagg_dice_ce_stats_fwd")
        ctx.save_for_backward(logits, tgt)
        ctx.meta = (dt, tdt)
        return stats, ce

    @staticmethod
    def backward(ctx, g_stats, g_ce):
        from . import _lib
        logits, tgt = ctx.saved_tensors
        dt, tdt = ctx.meta
        Bn, K, H, W = logits.shape
        g_stats = torch.zeros(Bn, K, 3, device=logits.device) if g_stats is None else g_stats.float().contiguous()
        g_ce = torch.zeros(1, device=logits.device) if g_ce is None else g_ce.float().contiguous()
        if dt & 2:      # same strides as the logits, with the 16-wide rows owned by the gradient tensor (padding <- zeros)
            dl = torch.empty(Bn, H, W, 16, device=logits.device, dtype=logits.dtype)[..., :K].permute(0, 3, 1, 2)
        else:
            dl = torch.empty_strided(logits.shape, logits.stride(), device=logits.device, dtype=logits.dtype)
        with torch.cuda.device(logits.device), _lib.timed("dice_ce_bwd"):
            rc = _lib.lib().mlagg_dice_ce_stats_bwd(logits.data_ptr(), tgt.data_ptr(), g_stats.data_ptr(), g_ce.data_ptr(),
                                                    dl.data_ptr(), Bn, H * W, K, logits.stride(0), logits.stride(1),
                                                    logits.stride(3), dt, tdt, _lib.stream_ptr())
        _lib.check(rc, "mlagg_dice_ce_stats_bwd")
        return dl, None


def dice_ce_loss_fused(logits, target, batch_dice=True, do_bg=False, smooth=1e-5, ddp=False):
    """cross_entropy + soft_dice_loss of one scale from one pass over the logits (same formulas as the two functions it
    replaces; the dice arithmetic runs on the (B, K, 3) statistics)."""
    stats, ce = _DiceCEStats.apply(logits, target)
    k0 = 0 if do_bg else 1
    inter, spred, sgt = stats[:, k0:, 0], stats[:, k0:, 1], stats[:, k0:, 2].detach()
    if ddp and batch_dice:
        inter, spred, sgt = _AllGatherGrad.apply(torch.stack((inter, spred, sgt), dim=0)).sum(0).unbind(0)
    if batch_dice:
        inter, spred, sgt = inter.sum(0), spred.sum(0), sgt.sum(0)
    dice = -((2 * inter + smooth) / torch.clip(sgt + spred + smooth, 1e-8)).mean()
    return ce[0] / (logits.shape[0] * logits.shape[2] * logits.shape[3]) + dice


class DeepSupervisionDiceCE(nn.Module):
    """DeepSupervisionWrapper(DC_and_CE_loss(weight_ce=1, weight_dice=1)) with weights 1/2**i, normalised."""

    def __init__(self, n_scales=5, batch_dice=True, ddp=False):
        super().__init__()
        w = np.array([1 / (2 ** i) for i in range(n_scales)])
        self.weights = (w / w.sum()).tolist()
        self.batch_dice, self.ddp = batch_dice, ddp

    def one(self, logits, target):
        if _DiceCEStats.addressable(logits, target) and os.environ.get("MLAGG_LOSS_TORCH") is None:
            return dice_ce_loss_fused(logits, target, self.batch_dice, False, 1e-5, self.ddp)
        ce = F.cross_entropy(logits.float(), target[:, 0].long())
        return ce + soft_dice_loss(logits, target, self.batch_dice, False, 1e-5, self.ddp)

    def forward(self, outputs, targets):
        if not isinstance(outputs, (list, tuple)):
            return self.one(outputs, targets[0] if isinstance(targets, (list, tuple)) else targets)
        if (os.environ.get("MLAGG_LOSS_TORCH") is None and len(outputs) > 1
                and all(_DiceCEStats.addressable(o, t) for o, t in zip(outputs, targets))
                and len({(o.shape[0], o.shape[1]) for o in outputs}) == 1):
            return self.all_scales(outputs, targets)
        return sum(w * self.one(o, t) for w, o, t in zip(self.weights, outputs, targets))

    def all_scales(self, outputs, targets, smooth=1e-5):
        """The same loss with the dice arithmetic of ALL scales done at once on the stacked (S, B, K, 3) statistics:
        one set of tiny element-wise kernels (and, data parallel, ONE packed all-gather) per step instead of one per
        scale -- ~150 fewer launches.  Per scale the formulas are exactly those of `dice_ce_loss_fused`."""
        pairs = [_DiceCEStats.apply(o, t) for o, t in zip(outputs, targets)]
        stats = torch.stack([st for st, _ in pairs], dim=0)                       # (S, B, K, 3)
        ce = torch.cat([c for _, c in pairs], dim=0)                              # (S,)
        key = (ce.device, tuple(tuple(o.shape) for o in outputs))
        if getattr(self, "_consts", (None,))[0] != key:      # built from device-side fills: safe inside a graph capture
            mk = lambda vals: torch.stack([torch.full((), float(v), device=ce.device) for v in vals])
            self._consts = (key, mk([o.shape[0] * o.shape[2] * o.shape[3] for o in outputs]), mk(self.weights[:len(outputs)]))
        _, npix, w = self._consts
        inter, spred, sgt = stats[:, :, 1:, 0], stats[:, :, 1:, 1], stats[:, :, 1:, 2].detach()
        if self.ddp and self.batch_dice:
            inter, spred, sgt = _AllGatherGrad.apply(torch.stack((inter, spred, sgt), dim=0)).sum(0).unbind(0)
        if self.batch_dice:
            inter, spred, sgt = inter.sum(1), spred.sum(1), sgt.sum(1)            # (S, K - 1)
            dice = -((2 * inter + smooth) / torch.clip(sgt + spred + smooth, 1e-8)).mean(1)
        else:
            dice = -((2 * inter + smooth) / torch.clip(sgt + spred + smooth, 1e-8)).mean((1, 2))
        return (w * (ce / npix + dice)).sum()


# ------------------------------------------------------------------ synthetic plans
@dataclass
class _LabelManager:
    num_segmentation_heads: int
    has_regions: bool = False
    ignore_label: int | None = None


@dataclass
class SyntheticPlan:
    """Stands in for PlansManager + ConfigurationManager + dataset_json of a `2d_bs10`-style plan."""
    patch_size: tuple = (320, 320)
    batch_size: int = 10
    num_classes: int = 14
    num_input_channels: int = 1
    batch_dice: bool = True
    dataset_json: dict = field(default_factory=dict)

    def get_label_manager(self, dataset_json=None):
        return _LabelManager(self.num_classes)


class TrainingLogger:
    """One value per epoch and key, `ema_fg_dice` derived from `mean_fg_dice` (0.9 / 0.1) -- the bookkeeping of the
    reference's nnUNetLogger (training/logging/nnunet_logger.py:17-51) without its matplotlib plots."""
    KEYS = ("mean_fg_dice", "ema_fg_dice", "dice_per_class_or_region", "train_losses", "val_losses", "lrs",
            "epoch_start_timestamps", "epoch_end_timestamps")

    def __init__(self):
        self.my_fantastic_logging = {k: [] for k in self.KEYS}

    def log(self, key, value, epoch: int):
        lst = self.my_fantastic_logging[key]
        assert len(lst) in (epoch, epoch + 1), "exactly one value per epoch and key"
        if len(lst) < epoch + 1:
            lst.append(value)
        else:
            lst[epoch] = value
        if key == "mean_fg_dice":
            ema = self.my_fantastic_logging["ema_fg_dice"]
            self.log("ema_fg_dice", ema[epoch - 1] * 0.9 + 0.1 * value if len(ema) > 0 else value, epoch)


class SyntheticLoader:
    """Endless iterator of synthetic batches in the shape the reference's augmenter yields (pattern:
    nnUNetTrainerBenchmark_5epochs_noDataLoading.py:16-22).  `pool` distinct batches are generated once and cycled."""

    def __init__(self, trainer, batch_size=None, pool=2, seed=0, pin=True):
        pin = pin and trainer.device.type == "cuda"
        self.batches = [trainer.synthetic_batch(batch_size, seed=seed + i, pin=pin) for i in range(pool)]
        self.i = 0

    def __iter__(self):
        return self

    def __next__(self):
        b = self.batches[self.i % len(self.batches)]
        self.i += 1
        return b


def split_batch(global_batch: int, world_size: int):
    """Per-rank batch sizes exactly as nnUNetTrainer._set_batch_size_and_oversample computes them
    (reference nnUNetTrainer.py:295-307).  NOTE: for 10 images on 8 ranks this yields
    [2,2,2,2,2,0,-2,-4] -- callers must reject non-positive entries."""
    assert global_batch >= world_size, "Cannot run DDP if the batch size is smaller than the number of GPUs"
    per = int(np.ceil(global_batch / world_size))
    return [per - ((r + 1) * per - global_batch) if (r + 1) * per > global_batch else per for r in range(world_size)]


class nnUNetTrainer_MLAgg_2D_dt_MS:
    def __init__(self, plans: SyntheticPlan | None = None, configuration: str = "2d_bs10", fold: int = 0,
                 dataset_json: dict | None = None, unpack_dataset: bool = True,
                 device: torch.device = torch.device("cuda"), amp_dtype=torch.bfloat16):
        self.plans = plans or SyntheticPlan()
        self.plans_manager = self.configuration_manager = self.plans
        self.dataset_json = dataset_json or self.plans.dataset_json
        self.label_manager = self.plans.get_label_manager(self.dataset_json)
        self.device = torch.device(device)
        self.initial_lr, self.weight_decay = 5e-4, 3e-5
        self.oversample_foreground_percent = 0.33
        self.num_iterations_per_epoch, self.num_val_iterations_per_epoch = 250, 50
        self.num_epochs, self.current_epoch = 500, 0
        self.is_ddp = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.amp_dtype = amp_dtype
        self.grad_scaler = torch.amp.GradScaler("cuda") if (amp_dtype == torch.float16 and self.device.type == "cuda") else None
        self.network = self.optimizer = self.lr_scheduler = self.loss = None
        self.use_cuda_graph = os.environ.get("MLAGG_CUDA_GRAPH", "1") != "0"
        self._graph = self._graph_key = self._static = self._flat_grad = self._params = self._flat_views = None
        self._flat_slices = None
        self._eager_steps = 0
        self.logger = TrainingLogger()
        self.output_folder = None            # set it to get checkpoint_latest / _best / _final like the reference
        self.save_every, self._best_ema = 50, None
        self.dataloader_train = self.dataloader_val = None
        self.local_rank = dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0

    # ---- reference static API
    @staticmethod
    def build_network_architecture(plans_manager, dataset_json, configuration_manager, num_input_channels,
                                   enable_deep_supervision: bool = True) -> nn.Module:
        label_manager = plans_manager.get_label_manager(dataset_json)
        return MLLA_Uper(img_size=configuration_manager.patch_size, patch_size=2, in_channels=num_input_channels,
                         out_channels=label_manager.num_segmentation_heads, embed_dim=96, depths=[2, 2, 2, 2],
                         num_heads=[2, 4, 8, 16], mlp_ratio=2, qkv_bias=True, drop_rate=0., dropout_path_rate=0.1,
                         sr_ratio=[16, 8, 4, 2], norm_layer=nn.LayerNorm, ape=False, use_checkpoint=False,
                         deep_supervision=enable_deep_supervision)

    def set_deep_supervision_enabled(self, enabled: bool):
        net = self.network.module if hasattr(self.network, "module") else self.network
        net.deep_supervision = enabled

    def _get_deep_supervision_scales(self):
        return [list(i) for i in 1 / np.cumprod(np.vstack([[1, 1], [2, 2], [2, 2], [2, 2], [2, 2]]), axis=0)]

    def _build_loss(self):
        return DeepSupervisionDiceCE(len(self._get_deep_supervision_scales()), self.plans.batch_dice, self.is_ddp)

    def configure_optimizers(self):
        # ALL parameters, in registration order, like the reference (`AdamW(self.network.parameters(), ...)`, :137-140):
        # the frozen `dummy_tensor` keeps its slot, so `optimizer_state` indices are interchangeable with the
        # reference's checkpoints in both directions (it never gets a gradient there either, so it has no state)
        params = list(self.network.parameters())
        cuda = self.device.type == "cuda"
        # fused + capturable: the step (and the learning rate, kept as a device scalar) can live inside a CUDA graph
        lr = torch.tensor(self.initial_lr, device=self.device, dtype=torch.float32) if cuda else self.initial_lr
        opt = torch.optim.AdamW(params, lr, weight_decay=self.weight_decay, eps=1e-4, fused=cuda, capturable=cuda)
        sched = CosineLRScheduler(opt, t_initial=self.num_epochs, lr_min=1e-6, warmup_t=10, warmup_lr_init=1e-4)
        return opt, sched

    # ---- lifecycle (nnUNetTrainer.initialize, :193-212)
    def initialize(self):
        self.network = self.build_network_architecture(self.plans_manager, self.dataset_json,
                                                       self.configuration_manager, self.plans.num_input_channels,
                                                       True).to(self.device)
        if self.device.type == "cuda":
            self.network = self.network.to(memory_format=torch.channels_last)
        if self.is_ddp:
            # data parallel without the DDP wrapper: identical start (rank 0's weights), one flat gradient all-reduce per
            # step (see _step_math).  The reference wraps in DDP (nnUNetTrainer.py:205-207), which cannot survive its own
            # unused `dummy_tensor` parameter; the arithmetic -- averaged gradients -- is the same.
            for t in list(self.network.parameters()) + list(self.network.buffers()):
                dist.broadcast(t.data, 0)
        self.optimizer, self.lr_scheduler = self.configure_optimizers()
        self._bind_flat_grads()
        if self.device.type == "cuda" and self.grad_scaler is None:
            from . import ops
            ops.build_cast_cache(self._params, self.amp_dtype)
        self.loss = self._build_loss()
        return self

    def synthetic_batch(self, batch_size=None, seed=0, device=None, pin=False):
        """{'data': (B,C,H,W) float, 'target': [ (B,1,H/s,W/s) ] x 5} -- the shape the reference's augmenter yields
        (pattern: nnUNetTrainerBenchmark_5epochs_noDataLoading.py:16-22)."""
        g = torch.Generator().manual_seed(seed)
        B = batch_size or self.plans.batch_size
        H, W = self.plans.patch_size
        data = torch.randn(B, self.plans.num_input_channels, H, W, generator=g)
        target = [torch.round(torch.rand(B, 1, int(H * s[0]), int(W * s[1]), generator=g) * (self.plans.num_classes - 1))
                  for s in self._get_deep_supervision_scales()]
        if pin:
            data, target = data.pin_memory(), [t.pin_memory() for t in target]
        if device is not None:
            data, target = data.to(device), [t.to(device) for t in target]
        return {"data": data, "target": target}

    # ---- checkpoints (nnUNetTrainer.save_checkpoint / load_checkpoint, nnUNetTrainer.py:1007-1054): same dict keys, so a
    # reference `checkpoint_final.pth` loads here and one written here loads in the reference
    def save_checkpoint(self, filename: str) -> None:
        if dist.is_available() and dist.is_initialized() and dist.get_rank() != 0:
            return
        net = self.network.module if hasattr(self.network, "module") else self.network
        torch.save({"network_weights": net.state_dict(), "optimizer_state": self._portable_optimizer_state(),
                    "grad_scaler_state": self.grad_scaler.state_dict() if self.grad_scaler is not None else None,
                    "logging": self.logger.my_fantastic_logging, "_best_ema": self._best_ema,
                    "current_epoch": self.current_epoch + 1,
                    "init_args": {"configuration": "2d_bs10", "fold": 0}, "trainer_name": self.__class__.__name__,
                    "inference_allowed_mirroring_axes": (0, 1)}, filename)

    def load_checkpoint(self, filename_or_checkpoint, load_optimizer: bool = True) -> None:
        if self.network is None:
            self.initialize()
        ck = filename_or_checkpoint
        if isinstance(ck, str):
            ck = torch.load(ck, map_location=self.device, weights_only=False)
        net = self.network.module if hasattr(self.network, "module") else self.network
        own = net.state_dict().keys()
        state = {(k[7:] if k not in own and k.startswith("module.") else k): v for k, v in ck["network_weights"].items()}
        net.load_state_dict(state)            # strict: names and shapes are the reference's
        self.current_epoch = ck.get("current_epoch", 0)
        if isinstance(ck.get("logging"), dict) and ck["logging"]:
            self.logger.my_fantastic_logging.update({k: list(v) for k, v in ck["logging"].items()
                                                     if k in self.logger.my_fantastic_logging})
        self._best_ema = ck.get("_best_ema", None)
        if load_optimizer and ck.get("optimizer_state") is not None:
            self.optimizer.load_state_dict(ck["optimizer_state"])
            self._restore_optimizer_flags()
        if self.grad_scaler is not None and ck.get("grad_scaler_state") is not None:
            self.grad_scaler.load_state_dict(ck["grad_scaler_state"])
        self._graph = None                    # captured graphs hold the old optimizer state tensors
        self._eager_steps = 0

    def _portable_optimizer_state(self):
        """optimizer.state_dict() in the form the reference's plain AdamW writes and expects: python-float lr, the stock
        foreach / fused / capturable settings, `step` as a CPU scalar tensor (a CUDA `step` trips the assertion of the
        non-capturable implementation)."""
        sd = self.optimizer.state_dict()
        groups = []
        for g in sd["param_groups"]:
            g = dict(g)
            g["lr"] = float(g["lr"])
            if "initial_lr" in g:
                g["initial_lr"] = float(g["initial_lr"])
            g.update(fused=None, capturable=False, foreach=None)
            groups.append(g)
        state = {k: {n: (v.detach().float().cpu() if n == "step" and torch.is_tensor(v) else v) for n, v in st.items()}
                 for k, st in sd["state"].items()}
        return {"state": state, "param_groups": groups}

    def _restore_optimizer_flags(self):
        """`load_state_dict` replaces the param groups with the checkpoint's: a reference checkpoint brings a python
        float `lr` and fused / capturable = False, which a captured step would bake in as constants.  Put back what
        configure_optimizers set up: lr as a device scalar (the scheduler writes into it), fused + capturable, and
        `step` counters as device tensors."""
        if self.device.type != "cuda":
            return
        for g in self.optimizer.param_groups:
            lr = g["lr"]
            g["lr"] = (lr.to(self.device, torch.float32) if torch.is_tensor(lr)
                       else torch.tensor(float(lr), device=self.device, dtype=torch.float32))
            g["fused"], g["capturable"], g["foreach"] = True, True, False
        for p, st in self.optimizer.state.items():
            if "step" in st and not (torch.is_tensor(st["step"]) and st["step"].is_cuda):
                st["step"] = torch.as_tensor(float(st["step"]), dtype=torch.float32, device=self.device)
            for k in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
                # the fused kernel pairs elements by memory order: moments take the parameter's strides (conv weights are
                # channels_last here, contiguous in a reference checkpoint)
                if k in st and st[k].stride() != p.stride():
                    st[k] = torch.empty_like(p).copy_(st[k])

    # ---- one training step (nnUNetTrainer.train_step, :833-863)
    def _forward_loss(self, data, target):
        from . import _lib, ops
        if self.device.type == "cuda":
            _lib.arena_begin(self.device)   # one memset instead of ~350 zero-fill kernels for the backward accumulators
        ops.refresh_cast_cache()      # one multi-tensor fp32 -> bf16 copy of the parameters instead of a cast per layer
        with torch.autocast(self.device.type, dtype=self.amp_dtype, enabled=self.device.type == "cuda"):
            return self.loss(self.network(data), target)

    def _drop_grads(self):
        for p in self._params:        # undefined .grad: autograd hands its gradient tensors over instead of adding
            p.grad = None

    @staticmethod
    def _attach_side_grads():
        """join the side stream and attach the parameter gradients that were produced on it (`_lib.stash_grad`)"""
        from . import _lib
        _lib.side_join()
        for p, g in _lib.take_stashed_grads():
            p.grad = g if p.grad is None else p.grad + g

    def _reduce_clip_step(self, grads=None):
        """Gather the per-parameter gradients into ONE flat fp32 buffer (a handful of batched-copy launches), all-reduce
        it once over NCCL / NVSwitch (108 MB: well under a millisecond, no DDP hooks or buckets), clip the global norm to
        12 with two kernels on the flat buffer, and let AdamW read its gradients as views of that buffer."""
        grads = [p.grad for p in self._params] if grads is None else grads
        # a parameter the step did not use (e.g. out_1..out_4 after set_deep_supervision_enabled(False)) has no gradient:
        # its slice of the flat buffer is zero (the reference's clip_grad_norm_ / AdamW skip it; AdamW with a zero
        # gradient still applies weight decay, so such parameters are ALSO dropped from the step below)
        missing = [i for i, g in enumerate(grads) if g is None]
        have = [(g, p, s) for g, p, s in zip(grads, self._params, self._flat_slices) if g is not None]
        parts = [(g if g.stride() == p.stride() else torch.empty_like(p).copy_(g)).as_strided((g.numel(),), (1,))
                 for g, p, _ in have]
        torch._foreach_copy_([s for _, _, s in have], parts)   # multi-tensor copy: 0.15 ms for 108 MB (torch.cat(out=): 0.55 ms)
        for i in missing:
            self._flat_slices[i].zero_()
        if self.is_ddp:
            dist.all_reduce(self._flat_grad, op=dist.ReduceOp.SUM)
            self._flat_grad.div_(dist.get_world_size())
        # torch.nn.utils.clip_grad_norm_(params, 12): total 2-norm over all gradients == norm of the flat buffer
        coef = torch.clamp(12.0 / (torch.linalg.vector_norm(self._flat_grad) + 1e-6), max=1.0)
        self._flat_grad.mul_(coef)
        skip = set(missing)
        for i, (p, v) in enumerate(zip(self._params, self._flat_views)):
            p.grad = None if i in skip else v
        self.optimizer.step()

    def _step_math(self, data, target):
        """forward -> DiceCE-DS loss -> backward -> gradient all-reduce -> clip(12) -> AdamW; capturable in a CUDA graph."""
        self._drop_grads()
        from . import _lib, ops
        l = self._forward_loss(data, target)
        if self.device.type == "cuda":
            _lib.side_begin(self.device)      # weight-gradient GEMMs run next to the main backward chain
        l.backward()
        self._attach_side_grads()
        ops.release_cast_cache()
        self._reduce_clip_step()
        _lib.arena_end()
        return l.detach()

    def _bind_flat_grads(self):
        self._params = [p for p in self.network.parameters() if p.requires_grad]
        self._flat_grad = torch.zeros(sum(p.numel() for p in self._params), device=self.device, dtype=torch.float32)
        self._flat_views, self._flat_slices, off = [], [], 0
        for p in self._params:
            dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
            assert p.dtype == torch.float32 and dense, "flat gradient views need dense fp32 parameters"
            # same strides as the parameter (channels_last conv weights): autograd's gradient layout contract gives
            # .grad the parameter's strides, and fused AdamW pairs elements by memory order
            self._flat_views.append(self._flat_grad[off:off + p.numel()].as_strided(p.size(), p.stride()))
            self._flat_slices.append(self._flat_grad[off:off + p.numel()])
            off += p.numel()

    def _alloc_static(self, batch):
        self._static = {"data": torch.empty_like(batch["data"], device=self.device),
                        "target": [torch.empty_like(t, device=self.device) for t in batch["target"]]}
        self._static["data"].copy_(batch["data"])
        for d, t in zip(self._static["target"], batch["target"]):
            d.copy_(t)

    def _capture_stream_kw(self):
        """The graphs are captured on a high-priority stream: the kernel nodes of the main chain then carry a higher
        priority than the side-stream nodes (weight gradients, default = lowest priority) and the block scheduler places
        their CTAs first -- measured +0.3 .. 0.6 ms per step on top of the side stream itself.  MLAGG_HP_STREAM=0: off."""
        if os.environ.get("MLAGG_HP_STREAM", "1") == "0":
            return {}
        if getattr(self, "_hp_stream", None) is None:
            self._hp_stream = torch.cuda.Stream(device=self.device, priority=-1)
        return {"stream": self._hp_stream}

    def _side_stream_warmup(self, fn, n=2):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):                       # warm-up on a side stream, as graph capture requires
            for _ in range(n):
                fn()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)

    def _capture(self, batch):
        """CUDA graphs over static input buffers.
        One process: ONE graph holds the whole step (forward, loss, backward, clip, AdamW).
        Data parallel: collectives stay outside the graphs -- graph A = zero grads + network forward, eager = loss (its
        batch-dice all-gathers) and d loss / d logits, graph B = network backward, eager = flat gradient all-reduce,
        clip, AdamW."""
        self._alloc_static(batch)
        st = self._static
        # the side-stream warm-up runs real steps: snapshot weights + optimizer state and put them back afterwards, so
        # that capturing never changes the training trajectory
        params = [p for p in self.network.parameters()]
        snap_p = [p.detach().clone() for p in params]
        snap_o = [(v, v.clone()) for stt in self.optimizer.state.values() for v in stt.values() if torch.is_tensor(v)]
        if not self.is_ddp:
            self._side_stream_warmup(lambda: self._step_math(st["data"], st["target"]))
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, **self._capture_stream_kw()):
                st["loss"] = self._step_math(st["data"], st["target"])
            self._graph = (graph,)
        else:
            self._side_stream_warmup(lambda: self._step_math(st["data"], st["target"]))
            ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            from . import ops
            from . import _lib
            with torch.cuda.graph(ga, **self._capture_stream_kw()):
                _lib.arena_begin(self.device)
                ops.refresh_cast_cache()
                with torch.autocast("cuda", dtype=self.amp_dtype):
                    outs = self.network(st["data"])
            outs = list(outs) if isinstance(outs, (list, tuple)) else [outs]
            st["outs"] = outs
            st["douts"] = [torch.zeros_like(o) for o in outs]
            self._drop_grads()
            with torch.cuda.graph(gb, pool=ga.pool(), **self._capture_stream_kw()):
                _lib.side_begin(self.device)
                torch.autograd.backward(outs, grad_tensors=st["douts"])
                self._attach_side_grads()
            ops.release_cast_cache()
            _lib.arena_end()
            st["grads"] = [p.grad for p in self._params]     # written in place by every replay of graph B
            self._graph = (ga, gb)
        with torch.no_grad():
            for p, c in zip(params, snap_p):
                p.copy_(c)
            for v, c in snap_o:
                v.copy_(c)
        self._graph_key = self._batch_key(batch)

    def _replay(self):
        st = self._static
        if len(self._graph) == 1:
            self._graph[0].replay()
            return st["loss"]
        ga, gb = self._graph
        ga.replay()
        heads = [o.detach().requires_grad_() for o in st["outs"]]
        with torch.autocast("cuda", dtype=self.amp_dtype):
            l = self.loss(heads if len(heads) > 1 else heads[0], st["target"])
        for d, g in zip(st["douts"], torch.autograd.grad(l, heads)):
            d.copy_(g)
        gb.replay()
        self._reduce_clip_step(st["grads"])
        return l.detach()

    @staticmethod
    def _batch_key(batch):
        return (tuple(batch["data"].shape), batch["data"].dtype, tuple(tuple(t.shape) for t in batch["target"]))

    def train_step(self, batch: dict, sync: bool = True) -> dict:
        from . import _lib
        if self.grad_scaler is not None:                    # fp16 + GradScaler (the reference's recipe): plain eager step
            return self._train_step_scaled(batch, sync)
        graphable = (self.use_cuda_graph and self.device.type == "cuda" and _lib.STATS["events"] is None
                     and self.network.training)
        if graphable and self._graph is None and self._eager_steps >= 3:
            self._capture(batch)
        if graphable and self._graph is not None and self._graph_key == self._batch_key(batch):
            self._static["data"].copy_(batch["data"], non_blocking=True)
            for d, t in zip(self._static["target"], batch["target"]):
                d.copy_(t, non_blocking=True)
            l = self._replay()
        else:
            data = batch["data"].to(self.device, non_blocking=True)
            target = [t.to(self.device, non_blocking=True) for t in batch["target"]]
            l = self._step_math(data, target)
            self._eager_steps += 1
        return {"loss": l.cpu().numpy() if sync else l}

    # ---- epoch loop (nnUNetTrainer.run_training and its hooks, nnUNetTrainer.py:784-1223), minus file logging / plots
    def print_to_log_file(self, *args):
        if self.local_rank == 0 and os.environ.get("MLAGG_QUIET", "0") != "1":
            print(*args, flush=True)

    def on_train_start(self):
        if self.network is None:
            self.initialize()
        if self.dataloader_train is None:                 # no batchgenerators in this image: synthetic plan data
            self.dataloader_train = SyntheticLoader(self, seed=1000 * self.local_rank)
            self.dataloader_val = SyntheticLoader(self, seed=1000 * self.local_rank + 500)
        if self.output_folder is not None:
            os.makedirs(self.output_folder, exist_ok=True)

    def on_epoch_start(self):
        import time
        self.logger.log("epoch_start_timestamps", time.time(), self.current_epoch)

    def on_train_epoch_start(self):
        self.network.train()
        self.lr_scheduler.step(self.current_epoch)        # once per epoch (:825); writes the device-resident lr
        lr = float(self.optimizer.param_groups[0]["lr"])
        self.print_to_log_file(f"Epoch {self.current_epoch}  lr {lr:.5g}")
        self.logger.log("lrs", lr, self.current_epoch)

    def on_train_epoch_end(self, train_outputs):
        losses = np.array([float(o["loss"]) for o in train_outputs])
        if self.is_ddp:
            gathered = [None] * dist.get_world_size()
            dist.all_gather_object(gathered, losses)
            loss_here = float(np.vstack(gathered).mean())
        else:
            loss_here = float(losses.mean())
        self.logger.log("train_losses", loss_here, self.current_epoch)

    def on_validation_epoch_start(self):
        self.network.eval()

    def validation_step(self, batch: dict) -> dict:
        """nnUNetTrainer.validation_step (:865-926): loss with deep supervision, then hard tp / fp / fn of the argmax mask
        of the full-resolution head per class (background dropped) for the online pseudo dice."""
        data = batch["data"].to(self.device, non_blocking=True)
        target = [t.to(self.device, non_blocking=True) for t in batch["target"]]
        with torch.no_grad(), torch.autocast(self.device.type, dtype=self.amp_dtype, enabled=self.device.type == "cuda"):
            output = self.network(data)
            l = self.loss(output, target)
        out0 = output[0] if isinstance(output, (list, tuple)) else output
        tgt0 = target[0].long()
        C = out0.shape[1]
        pred = out0.argmax(1, keepdim=True)
        axes = (0, 2, 3)
        p1 = torch.zeros(out0.shape, device=out0.device, dtype=torch.float32).scatter_(1, pred, 1)
        g1 = torch.zeros(out0.shape, device=out0.device, dtype=torch.float32).scatter_(1, tgt0, 1)
        tp, fp, fn = (p1 * g1).sum(axes), (p1 * (1 - g1)).sum(axes), ((1 - p1) * g1).sum(axes)
        assert tp.shape == (C,)
        return {"loss": l.detach().float().cpu().numpy(), "tp_hard": tp.cpu().numpy()[1:], "fp_hard": fp.cpu().numpy()[1:],
                "fn_hard": fn.cpu().numpy()[1:]}

    def on_validation_epoch_end(self, val_outputs):
        tp, fp, fn = (np.sum([o[k] for o in val_outputs], 0) for k in ("tp_hard", "fp_hard", "fn_hard"))
        losses = np.array([float(o["loss"]) for o in val_outputs])
        if self.is_ddp:
            ws = dist.get_world_size()
            parts = [None] * ws
            dist.all_gather_object(parts, (tp, fp, fn, losses))
            tp, fp, fn = (np.sum([q[i] for q in parts], 0) for i in range(3))
            losses = np.concatenate([q[3] for q in parts])
        with np.errstate(divide="ignore", invalid="ignore"):
            dc = [float(2 * i / (2 * i + j + k)) if (2 * i + j + k) > 0 else float("nan") for i, j, k in zip(tp, fp, fn)]
        self.logger.log("mean_fg_dice", float(np.nanmean(dc)), self.current_epoch)
        self.logger.log("dice_per_class_or_region", dc, self.current_epoch)
        self.logger.log("val_losses", float(losses.mean()), self.current_epoch)

    def on_epoch_end(self):
        import time
        lg = self.logger.my_fantastic_logging
        self.logger.log("epoch_end_timestamps", time.time(), self.current_epoch)
        self.print_to_log_file(f"train_loss {lg['train_losses'][-1]:.4f}  val_loss {lg['val_losses'][-1]:.4f}  "
                               f"pseudo dice {lg['mean_fg_dice'][-1]:.4f}  "
                               f"epoch time {lg['epoch_end_timestamps'][-1] - lg['epoch_start_timestamps'][-1]:.2f} s")
        if self.output_folder is not None:
            if (self.current_epoch + 1) % self.save_every == 0 and self.current_epoch != self.num_epochs - 1:
                self.save_checkpoint(os.path.join(self.output_folder, "checkpoint_latest.pth"))
            if self._best_ema is None or lg["ema_fg_dice"][-1] > self._best_ema:
                self._best_ema = lg["ema_fg_dice"][-1]
                self.save_checkpoint(os.path.join(self.output_folder, "checkpoint_best.pth"))
        elif self._best_ema is None or lg["ema_fg_dice"][-1] > self._best_ema:
            self._best_ema = lg["ema_fg_dice"][-1]
        self.current_epoch += 1

    def on_train_end(self):
        if self.output_folder is not None:
            self.save_checkpoint(os.path.join(self.output_folder, "checkpoint_final.pth"))
            latest = os.path.join(self.output_folder, "checkpoint_latest.pth")
            if self.local_rank == 0 and os.path.isfile(latest):
                os.remove(latest)

    def run_training(self):
        """nnUNetTrainer.run_training (:1202-1223), hook for hook."""
        self.on_train_start()
        for _ in range(self.current_epoch, self.num_epochs):
            self.on_epoch_start()
            self.on_train_epoch_start()
            train_outputs = [self.train_step(next(self.dataloader_train)) for _ in range(self.num_iterations_per_epoch)]
            self.on_train_epoch_end(train_outputs)
            with torch.no_grad():
                self.on_validation_epoch_start()
                val_outputs = [self.validation_step(next(self.dataloader_val))
                               for _ in range(self.num_val_iterations_per_epoch)]
                self.on_validation_epoch_end(val_outputs)
            self.on_epoch_end()
        self.on_train_end()

    def _train_step_scaled(self, batch: dict, sync: bool = True) -> dict:
        data = batch["data"].to(self.device, non_blocking=True)
        target = [t.to(self.device, non_blocking=True) for t in batch["target"]]
        self._drop_grads()
        l = self._forward_loss(data, target)
        params = self._params
        self.grad_scaler.scale(l).backward()
        from . import _lib
        _lib.arena_end()
        if self.is_ddp:
            for p in params:
                dist.all_reduce(p.grad, op=dist.ReduceOp.AVG)
        self.grad_scaler.unscale_(self.optimizer)
        torch.nn.utils.clip_grad_norm_(params, 12)
        self.grad_scaler.step(self.optimizer)
        self.grad_scaler.update()
        return {"loss": l.detach().cpu().numpy() if sync else l.detach()}
