"""oracle/network.py -- TEST INFRASTRUCTURE ONLY (parity oracle / CPU baseline).

Whole-network CPU forward of MLLA_Uper (reference nnUNetTrainer_MLAgg_2D_dt_MS.py:1370-1407) for a given set of
weights: the named hot path (8 MLAgg blocks, MSMM) runs through the oracle's functional restatements
(oracle/mlagg.py, oracle/msmm.py, oracle/scan_ref.c); the conv stages, which are plain torch modules kept on
PyTorch in the product too, are called on the `net` instance that carries the weights.  Differentiable
(torch autograd + the analytic scan backward), so it also serves as the reference's CPU fwd+bwd for bench.py's
`cpu_baseline` / `--impl reference` legs.  The product never imports this module.
"""
from __future__ import annotations

import torch

from .mlagg import mlla_block_forward
from .msmm import vss_conv_block_forward
from .scan import selective_scan_oracle


def mlla_uper_forward(net, x, scan=selective_scan_oracle):
    """net: an MLLA_Uper-shaped module on CPU (weights + conv-stage submodules); x (B, C, H, W) on CPU."""
    p = dict(net.named_parameters())
    enc = net.mlla
    hs = [x]
    t = enc.patch_embed(x)
    for i, layer in enumerate(enc.layers):
        for j, blk in enumerate(layer.blocks):
            t = mlla_block_forward(p, t, blk.num_heads, blk.sr_ratio, prefix=f"mlla.layers.{i}.blocks.{j}.")
        hs.append(t)
        if i < enc.num_layers - 1:
            t = enc.downs[i](t)
    hidden = net.mambaskip.hidden_dim
    skips = hs[1:]
    for b in range(len(net.mambaskip.blocks)):
        skips = vss_conv_block_forward(p, skips, hidden, scan=scan, prefix=f"mambaskip.blocks.{b}.")
    hs[1:] = skips
    ds = net.deep_supervision
    o4 = net.out_4(hs[4]) if ds else None
    y = net.dec_block_2(hs[3] + net.up_2(hs[4]))
    o3 = net.out_3(y) if ds else None
    y = net.dec_block_1(hs[2] + net.up_1(y))
    o2 = net.out_2(y) if ds else None
    y = net.dec_block_0(hs[1] + net.up_0(y))
    o1 = net.out_1(y) if ds else None
    y = net.out_0(net.decoder0(y, net.encoder0(hs[0])))
    return [y, o1, o2, o3, o4] if ds else y
