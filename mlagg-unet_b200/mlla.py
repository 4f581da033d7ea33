"""elu+1 linear attention block of MLLA (the ops BASELINE.json:north_star names) -- drop-in for
`RoPE` (:169-195), `LinearAttention` (:198-253) and `MLLABlock` (:256-319) of the reference's
`mlagg/nnunetv2/training/nnUNetTrainer/nnUNetTrainer_MLLA_UNet.py`, same ctor arguments and parameter names
(`rope.rotations` stays a buffer so reference checkpoints load with strict=True).

The depthwise convs (cpe1, dwc + SiLU, lepe, cpe2) run as the sm_100a tokens-major stencil kernel; the
attention core is attention.linear_attention_core.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import attention as att
from .mlagg import Mlp
from .ops import dwconv3x3_tokens, layer_norm_tokens, linear_tokens, residual_drop_path, silu_gate
from .thirdparty_shims import DropPath


class RoPE(nn.Module):
    def __init__(self, shape, base=10000):
        super().__init__()
        channel_dims, feature_dim = shape[:-1], shape[-1]
        k_max = feature_dim // (2 * len(channel_dims))
        assert feature_dim % k_max == 0 and len(channel_dims) == 2
        self.H, self.W, self.C = channel_dims[0], channel_dims[1], feature_dim
        cs, sn = att.rope_tables(self.H, self.W, self.C, torch.device("cpu"), float(base))
        self.register_buffer("rotations", torch.stack([cs, sn], dim=-1).reshape(self.H, self.W, self.C // 2, 2))

    def forward(self, x):
        """x (B, H, W, C) -> rotated, fp32"""
        Bn = x.shape[0]
        return att.rope_apply(x.reshape(Bn, self.H * self.W, self.C), self.H, self.W).reshape(x.shape)


class LinearAttention(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, qkv_bias=True, **kwargs):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads = dim, input_resolution, num_heads
        self.qk = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.elu = nn.ELU()
        self.lepe = nn.Conv2d(dim, dim, 3, padding=1, groups=dim)
        self.rope = RoPE(shape=(input_resolution[0], input_resolution[1], dim))

    def forward(self, x):
        """x (B, N, C) -> (B, N, C)"""
        H, W = self.input_resolution
        o = att.linear_attention_qk(linear_tokens(x, self.qk), x, H, W, self.num_heads)
        return dwconv3x3_tokens(x, self.lepe.weight, self.lepe.bias, H, W, residual=o)

    def extra_repr(self):
        return f"dim={self.dim}, num_heads={self.num_heads}"


class MLLABlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, mlp_ratio=4., qkv_bias=True, drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, **kwargs):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads, self.mlp_ratio = dim, input_resolution, num_heads, mlp_ratio
        self.cpe1 = nn.Conv2d(dim, dim, 3, padding=1, groups=dim)
        self.norm1 = norm_layer(dim)
        self.in_proj = nn.Linear(dim, dim)
        self.act_proj = nn.Linear(dim, dim)
        self.dwc = nn.Conv2d(dim, dim, 3, padding=1, groups=dim)
        self.act = nn.SiLU()
        self.attn = LinearAttention(dim=dim, input_resolution=input_resolution, num_heads=num_heads, qkv_bias=qkv_bias)
        self.out_proj = nn.Linear(dim, dim)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.cpe2 = nn.Conv2d(dim, dim, 3, padding=1, groups=dim)
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x):
        H, W = self.input_resolution
        Bn, L, C = x.shape
        assert L == H * W, "input feature has wrong size"
        x = dwconv3x3_tokens(x, self.cpe1.weight, self.cpe1.bias, H, W, residual=x)
        shortcut = x
        t = layer_norm_tokens(x, self.norm1)
        gate = linear_tokens(t, self.act_proj)                  # SiLU applied inside the gate kernel below
        t = dwconv3x3_tokens(linear_tokens(t, self.in_proj), self.dwc.weight, self.dwc.bias, H, W, silu=True)
        t = self.attn(t)
        x = residual_drop_path(shortcut, linear_tokens(silu_gate(t.to(gate.dtype), gate), self.out_proj), self.drop_path)
        x = dwconv3x3_tokens(x, self.cpe2.weight, self.cpe2.bias, H, W, residual=x)
        return residual_drop_path(x, self.mlp(layer_norm_tokens(x, self.norm2)), self.drop_path)

    def extra_repr(self):
        return (f"dim={self.dim}, input_resolution={self.input_resolution}, num_heads={self.num_heads}, "
                f"mlp_ratio={self.mlp_ratio}")
