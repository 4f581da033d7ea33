"""OVERLAY for the reference tree: copy to mlagg/nnunetv2/training/nnUNetTrainer/nnUNetTrainer_MLAgg_2D_dt_MS.py.

`nnUNetv2_train ... -tr nnUNetTrainer_MLAgg_2D_dt_MS` finds the trainer BY CLASS NAME under this directory
(run/run_training.py:39-40, recursive_find_python_class), so the class below subclasses the real nnUNetTrainer and
keeps the reference's hyper-parameters (:52-59) and static network factory (:62-92), while the network it builds is
mlagg_unet_b200's MLLA_Uper: same state_dict keys and shapes, hot path on the sm_100a kernels.

What is different from the reference file, all documented in DESIGN.md / INTEGRATION.md:
  * train_step delegates to the B200 step (bf16 autocast, flat-buffer gradient all-reduce instead of the DDP wrapper,
    CUDA graphs after three eager steps);  set MLAGG_REFERENCE_STEP=1 to keep nnUNetTrainer.train_step (fp16 + GradScaler);
  * set_deep_supervision_enabled reaches through `.module` under DDP;
  * `dummy_tensor` stays in the optimizer's parameter list but is frozen, so DDP never waits for its gradient.
Data loading, augmentation, validation export and the evaluation scripts are the reference's own code, untouched.
"""
import os

import numpy as np
import torch
from torch import nn

from nnunetv2.training.nnUNetTrainer.nnUNetTrainer import nnUNetTrainer
from nnunetv2.training.loss.compound_losses import DC_and_CE_loss
from nnunetv2.training.loss.deep_supervision import DeepSupervisionWrapper
from nnunetv2.training.loss.dice import MemoryEfficientSoftDiceLoss
from nnunetv2.utilities.plans_handling.plans_handler import ConfigurationManager, PlansManager

from mlagg_unet_b200 import trainer as _b200
from mlagg_unet_b200.mlagg import MLLA_Uper  # noqa: F401  (re-exported: inference code imports it from this module)


class nnUNetTrainer_MLAgg_2D_dt_MS(nnUNetTrainer):
    def __init__(self, plans: dict, configuration: str, fold: int, dataset_json: dict, unpack_dataset: bool = True,
                 device: torch.device = torch.device("cuda")):
        super().__init__(plans, configuration, fold, dataset_json, unpack_dataset, device)
        self.initial_lr = 5e-4
        self.weight_decay = 3e-5
        self.oversample_foreground_percent = 0.33
        self.num_iterations_per_epoch = 250
        self.num_val_iterations_per_epoch = 50
        self.num_epochs = 500
        self.current_epoch = 0
        self._b200 = None

    build_network_architecture = staticmethod(_b200.nnUNetTrainer_MLAgg_2D_dt_MS.build_network_architecture)

    def set_deep_supervision_enabled(self, enabled: bool):
        (self.network.module if hasattr(self.network, "module") else self.network).deep_supervision = enabled

    def _get_deep_supervision_scales(self):
        return list(list(i) for i in 1 / np.cumprod(np.vstack([[1, 1], [2, 2], [2, 2], [2, 2], [2, 2]]), axis=0))

    def _build_loss(self):
        loss = DC_and_CE_loss({"batch_dice": self.configuration_manager.batch_dice, "smooth": 1e-5, "do_bg": False,
                               "ddp": self.is_ddp}, {}, weight_ce=1, weight_dice=1,
                              ignore_label=self.label_manager.ignore_label, dice_class=MemoryEfficientSoftDiceLoss)
        scales = self._get_deep_supervision_scales()
        weights = np.array([1 / (2 ** i) for i in range(len(scales))])
        return DeepSupervisionWrapper(loss, weights / weights.sum())

    def configure_optimizers(self):
        return _b200.nnUNetTrainer_MLAgg_2D_dt_MS.configure_optimizers(self)

    # ---- the B200 step: adopt this trainer's network / optimizer / loss into the standalone step implementation
    def _adopt(self):
        t = _b200.nnUNetTrainer_MLAgg_2D_dt_MS.__new__(_b200.nnUNetTrainer_MLAgg_2D_dt_MS)
        _b200.nnUNetTrainer_MLAgg_2D_dt_MS.__init__(t, None, device=self.device)
        net = self.network.module if hasattr(self.network, "module") else self.network
        t.network, t.optimizer, t.lr_scheduler, t.loss = net, self.optimizer, self.lr_scheduler, self.loss
        t.is_ddp = self.is_ddp
        t._bind_flat_grads()
        from mlagg_unet_b200 import ops
        ops.build_cast_cache(t._params, t.amp_dtype)
        return t

    def train_step(self, batch: dict) -> dict:
        if os.environ.get("MLAGG_REFERENCE_STEP", "0") == "1":
            return super().train_step(batch)
        if self._b200 is None:
            self._b200 = self._adopt()
        return self._b200.train_step(batch)
