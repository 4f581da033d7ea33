"""Minimal torch restatements of third-party pieces the reference network imports but this image
does not have (SURVEY.md 8b "Missing third-party at the boundary", App. E):

  timm.models.layers.DropPath / to_2tuple / trunc_normal_
  timm.scheduler.CosineLRScheduler      (nnUNetTrainer_MLAgg_2D_dt_MS.py:137-147)
  monai.networks.blocks.UnetrBasicBlock (nnUNetTrainer_MLAgg_2D_dt_MS.py:1339-1347)
  monai.networks.blocks.UnetrUpBlock    (nnUNetTrainer_MLAgg_2D_dt_MS.py:1349-1357)

They are OFF the named hot path (conv stages stay on cuDNN) and exist so that the full network
builds with the reference's parameter names (`layer.conv1.conv.weight`, `transp_conv.conv.weight`,
`conv_block.*`).  If the real packages are importable they are not needed.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def _inst_norm(norm, x, act=None):
    """nn.InstanceNorm2d / nn.GroupNorm(num_groups=C) [+ LeakyReLU(0.01) | SiLU] of a channels_last map: the sm_100a
    kernel when it applies (CUDA, C % 4 == 0, fp32 / bf16), the torch module otherwise (CPU oracle runs, odd widths).
    torch's own kernels force an NCHW copy of every channels_last activation."""
    from .ops import instance_norm_cl, supports_instance_norm_cl
    ok = supports_instance_norm_cl(x) and not getattr(norm, "track_running_stats", False)
    if isinstance(norm, nn.GroupNorm):
        ok = ok and norm.num_groups == norm.num_channels
    if ok:
        return instance_norm_cl(x, norm.weight, norm.bias, norm.eps, act, 0.01)
    y = norm(x)
    return F.leaky_relu(y, 0.01) if act == "leaky_relu" else F.silu(y) if act == "silu" else y


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def trunc_normal_(t, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b)


class DropPath(nn.Module):
    """Stochastic depth: per-sample Bernoulli(1-p) mask scaled by 1/(1-p); identity in eval."""

    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = float(drop_prob)
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask

    def extra_repr(self):
        return f"drop_prob={self.drop_prob:0.3f}"


class CosineLRScheduler:
    """timm semantics with cycle_limit=1, warmup_prefix=False: `step(epoch)` sets every group's lr."""

    def __init__(self, optimizer, t_initial, lr_min=0.0, warmup_t=0, warmup_lr_init=0.0):
        self.optimizer = optimizer
        self.t_initial, self.lr_min = t_initial, lr_min
        self.warmup_t, self.warmup_lr_init = warmup_t, warmup_lr_init
        self.base_values = [float(g["lr"]) for g in optimizer.param_groups]
        for g in optimizer.param_groups:
            g.setdefault("initial_lr", g["lr"])
        if warmup_t:
            self._apply([warmup_lr_init] * len(self.base_values))

    def _values(self, t):
        if t < self.warmup_t:
            return [self.warmup_lr_init + t * (v - self.warmup_lr_init) / self.warmup_t for v in self.base_values]
        if t >= self.t_initial:
            return [self.lr_min for _ in self.base_values]
        return [self.lr_min + 0.5 * (v - self.lr_min) * (1 + math.cos(math.pi * t / self.t_initial))
                for v in self.base_values]

    def _apply(self, values):
        for g, v in zip(self.optimizer.param_groups, values):
            if isinstance(g["lr"], torch.Tensor):
                g["lr"].fill_(v)     # capturable optimizers keep lr on the device (CUDA-graph safe)
            else:
                g["lr"] = v

    def step(self, epoch, metric=None):
        self._apply(self._values(epoch))

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != "optimizer"}

    def load_state_dict(self, sd):
        self.__dict__.update(sd)


class _Conv(nn.Module):
    """monai `Convolution(conv_only=True)` naming: the torch conv sits at `.conv`."""

    def __init__(self, cin, cout, k, stride=1, transposed=False):
        super().__init__()
        if transposed:
            self.conv = nn.ConvTranspose2d(cin, cout, k, stride=stride, padding=(k - stride + 1) // 2,
                                           output_padding=2 * ((k - stride + 1) // 2) + stride - k, bias=False)
        else:
            self.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=(k - 1) // 2, bias=False)

    def forward(self, x):
        if isinstance(self.conv, nn.Conv2d) and self.conv.bias is None and x.is_cuda:
            from .ops import conv2d_split
            return conv2d_split(self.conv, x)        # NHWC strides for one-channel inputs; weight gradient on the side stream
        return self.conv(x)


class UnetResBlock2d(nn.Module):
    """monai.networks.blocks.dynunet_block.UnetResBlock, 2-D, norm 'instance' (no affine), LeakyReLU(0.01)."""

    def __init__(self, cin, cout, kernel_size=3, stride=1):
        super().__init__()
        self.conv1 = _Conv(cin, cout, kernel_size, stride)
        self.conv2 = _Conv(cout, cout, kernel_size, 1)
        self.lrelu = nn.LeakyReLU(0.01, inplace=True)
        self.norm1 = nn.InstanceNorm2d(cout)
        self.norm2 = nn.InstanceNorm2d(cout)
        self.downsample = cin != cout or stride != 1
        if self.downsample:
            self.conv3 = _Conv(cin, cout, 1, stride)
            self.norm3 = nn.InstanceNorm2d(cout)

    def forward(self, x):
        res = x
        y = _inst_norm(self.norm1, self.conv1(x), "leaky_relu")
        if self.downsample:
            res = _inst_norm(self.norm3, self.conv3(res))
        c2 = self.conv2(y)
        if not getattr(self.norm2, "track_running_stats", False) and self.lrelu.negative_slope == 0.01:
            from .ops import instance_norm_res_cl
            out = instance_norm_res_cl(c2, res, self.norm2.weight, self.norm2.bias, self.norm2.eps, "leaky_relu", 0.01)
            if out is not None:           # lrelu(norm2(.) + res): add and activation inside the normalisation's apply pass
                return out
        return self.lrelu(_inst_norm(self.norm2, c2) + res)


class UnetrBasicBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name, res_block=False):
        super().__init__()
        assert spatial_dims == 2 and res_block and norm_name == "instance"
        self.layer = UnetResBlock2d(in_channels, out_channels, kernel_size, stride)

    def forward(self, x):
        return self.layer(x)


class UnetrUpBlock(nn.Module):
    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, upsample_kernel_size, norm_name,
                 res_block=False):
        super().__init__()
        assert spatial_dims == 2 and res_block and norm_name == "instance"
        self.transp_conv = _Conv(in_channels, out_channels, upsample_kernel_size, upsample_kernel_size, True)
        self.conv_block = UnetResBlock2d(2 * out_channels, out_channels, kernel_size, 1)

    def forward(self, inp, skip):
        conv = self.transp_conv.conv
        a, b = inp.permute(0, 2, 3, 1), skip.permute(0, 2, 3, 1)
        if (inp.is_cuda and a.is_contiguous() and b.is_contiguous() and conv.kernel_size == (2, 2) and conv.stride == (2, 2)
                and conv.padding == (0, 0) and conv.bias is None and conv.weight.shape[0] % 8 == 0
                and conv.weight.shape[1] % 8 == 0 and skip.shape[-2:] == (2 * inp.shape[2], 2 * inp.shape[3])):
            # kernel 2 / stride 2: every output pixel has ONE contributing input pixel -- a per-token GEMM against the
            # (4 Co, Ci) matrix W[ci, co, di, dj] -> [(di, dj, co), ci], then a pixel shuffle fused with the skip concat
            from .ops import UpShuffleJoin, _Linear
            Bn, Ci, H, W = inp.shape
            w4 = conv.weight.permute(2, 3, 1, 0).reshape(-1, Ci)
            y = _Linear.apply(a.reshape(Bn, H * W, Ci), w4, None)
            x = UpShuffleJoin.apply(y, b, H, W).permute(0, 3, 1, 2)
        else:
            up = self.transp_conv(inp)
            a = up.permute(0, 2, 3, 1)
            if up.is_cuda and a.is_contiguous() and b.is_contiguous() and up.dtype == skip.dtype:
                from .ops import JoinLastDense
                x = JoinLastDense.apply(a, b).permute(0, 3, 1, 2)
            else:
                x = torch.cat((up, skip), dim=1)
        return self.conv_block(x)
