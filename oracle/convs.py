"""oracle/convs.py -- TEST INFRASTRUCTURE ONLY (parity oracle / CPU baseline).

Depthwise convolutions of the hot path:
  * depthwise 3x3 (+ optional SiLU) on tokens-major data -- the op behind `dwc`, `lepe`,
    `SS2D_skip.conv2d[i]`, `DWConv`, `cpe1/2` (reference
    nnUNetTrainer_MLAgg_2D_dt_MS.py:851,890,680,782; MambaSkip.py:302-312,545-556).
  * causal_conv1d -- named by north_star; not called by the shipped trainer.  The
    arithmetic lives in the un-vendored `causal-conv1d` package (reference README.md:49); its
    published definition y[b,c,t] = bias[c] + sum_j w[c,j] x[b,c,t-(k-1)+j] is restated.
    PARITY UNPINNED against that package (absent everywhere); pinned instead to the causal conv1d + SiLU inside
    Hugging Face transformers' `MambaMixer.slow_forward` (tests/test_oracle_pin_hf.py, forward and gradients).
Nothing under mlagg-unet_b200/ imports this module.
"""
from __future__ import annotations

import torch.nn.functional as F

from .msmm import dwconv3x3_tokens  # noqa: F401  (re-export)


def dwconv3x3_tokens_act(x, weight, bias, H, W, silu=False):
    y = dwconv3x3_tokens(x, weight, bias, H, W)
    return F.silu(y) if silu else y


def causal_conv1d(x, weight, bias=None, silu=False):
    """x (B,C,L), weight (C,k) -> (B,C,L): left-padded depthwise conv, optional SiLU."""
    C, k = weight.shape
    y = F.conv1d(x, weight[:, None, :], bias, padding=k - 1, groups=C)[..., : x.shape[-1]]
    return F.silu(y) if silu else y
