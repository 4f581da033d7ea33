"""MLAgg-UNet network modules -- drop-in for the classes of the reference's
`mlagg/nnunetv2/training/nnUNetTrainer/nnUNetTrainer_MLAgg_2D_dt_MS.py` with identical constructor
arguments, parameter names and shapes (a reference checkpoint loads with strict=True):

  hot path:  RMSNorm :592-610, AggregatedAttention :625-784, MLLABlock :824-915, Mlp :176-192,
                            BasicLayer :918-969, MLLA_Enc :1046-1179, and the MSMM (mamba_skip.VSS_Conv_Layer)
  conv stages (kept on PyTorch/cuDNN, off the named path): project / PatchEmbed :972-1043, MedNeXtBlock :230-324,
                            MedNeXtDownBlock :327-366, PatchExpand :479-546, OutBlock :549-561, MLLA_Uper :1183-1407

B200-first differences inside the hot path: the block keeps activations tokens-major from norm1 to the MLP
(the reference permutes to NCHW and back around every depthwise conv); `dwc`+SiLU and LePE are the sm_100a
stencil kernel; q's head_dim**-0.5 is applied once on the projection output; the attention cores are the
functions in attention.py.
"""
from __future__ import annotations

from collections.abc import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import attention as att
from .mamba_skip import VSS_Conv_Layer
from .ops import (Conv2dCL, ConvTranspose2dCL, GradContiguous, PadTopLeftAdd, SplitKV, avgpool_tokens, dwconv3x3_tokens, layer_norm_fork, layer_norm_tokens, linear_tokens,
                  mlp_gelu_tokens, residual_drop_path, silu_gate)
from .thirdparty_shims import DropPath, UnetrBasicBlock, UnetrUpBlock, _inst_norm


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        if isinstance(self.act, nn.GELU) and self.act.approximate == "none" and (self.drop.p == 0. or not self.training):
            return mlp_gelu_tokens(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)
        return self.drop(linear_tokens(self.drop(self.act(linear_tokens(x, self.fc1))), self.fc2))


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-6, elementwise_affine=True, memory_efficient=False):
        super().__init__()
        self.dim, self.eps, self.elementwise_affine = dim, eps, elementwise_affine
        if elementwise_affine:
            self.weight = nn.Parameter(torch.ones(dim))
        else:
            self.register_parameter("weight", None)

    def forward(self, x):
        y = x.float()
        y = (y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + self.eps)).type_as(x)
        return y if self.weight is None else y * self.weight

    def extra_repr(self):
        return f"dim={self.dim}, eps={self.eps}, elementwise_affine={self.elementwise_affine}"


@torch.no_grad()
def get_seqlen_and_mask(input_resolution, window_size):
    """(N,1) count of in-image neighbours and (N, window**2) bool mask, True = outside the image (:616-622)."""
    H, W = input_resolution
    h = window_size // 2
    r = torch.arange(H).view(H, 1, 1, 1) + torch.arange(-h, h + 1).view(1, 1, -1, 1)
    c = torch.arange(W).view(1, W, 1, 1) + torch.arange(-h, h + 1).view(1, 1, 1, -1)
    outside = ~((r >= 0) & (r < H) & (c >= 0) & (c < W)).reshape(H * W, window_size ** 2)
    return (~outside).sum(-1, keepdim=True).float(), outside


class _Linear1x1:
    """nn.Conv2d(C, C, 1) on a tokens-major map == per-token Linear with the conv's (C, C, 1, 1) weight viewed (C, C)."""

    @staticmethod
    def apply_conv(x, conv):
        from .ops import _Linear
        return _Linear.apply(x, conv.weight.view(conv.out_channels, conv.in_channels), conv.bias)


class AggregatedAttention(nn.Module):
    def __init__(self, dim, input_resolution, num_heads=8, local=True, window_size=3, qkv_bias=True, attn_drop=0.,
                 proj_drop=0., sr_ratio=1, fixed_pool_size=None):
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} should be divided by num_heads {num_heads}."
        self.dim, self.num_heads, self.local = dim, num_heads, local
        self.head_dim = dim // num_heads // 2
        self.scale = self.head_dim ** -0.5
        self.lambda_init = att.LAMBDA_INIT
        mk = lambda: nn.Parameter(torch.zeros(self.head_dim, dtype=torch.float32).normal_(mean=0, std=0.1))
        self.lambda_q1, self.lambda_k1, self.lambda_q2, self.lambda_k2 = mk(), mk(), mk(), mk()
        self.subln = RMSNorm(2 * self.head_dim, eps=1e-5, elementwise_affine=True)
        if local:
            assert window_size == 3, "the local branch is a 3x3 window"
            self.window_size, self.local_len = window_size, window_size ** 2
            _, mask = get_seqlen_and_mask(input_resolution, window_size)
            self.register_buffer("padding_mask", mask, persistent=False)
        else:
            self.sr_ratio = sr_ratio
            if fixed_pool_size is None:
                self.pool_H, self.pool_W = input_resolution[0] // sr_ratio, input_resolution[1] // sr_ratio
            else:
                assert fixed_pool_size < min(input_resolution)
                self.pool_H = self.pool_W = fixed_pool_size
            self.pool_len = self.pool_H * self.pool_W
            self.pool = nn.AdaptiveAvgPool2d((self.pool_H, self.pool_W))
            self.sr = nn.Conv2d(dim, dim, kernel_size=1, stride=1, padding=0)
            self.norm = nn.LayerNorm(dim)
            self.act = nn.GELU()
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.kv = nn.Linear(dim, dim * 2, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.lepe = nn.Conv2d(dim, dim, 3, padding=1, groups=dim)

    def forward(self, x, H, W):
        Bn, N, C = x.shape
        assert N == H * W
        h, hd = self.num_heads, self.head_dim
        q = linear_tokens(x, self.q)
        kv, v_local = SplitKV.apply(linear_tokens(x, self.kv), C)      # v_local = kv[..., C:]; gradients re-joined in place
        lam = att.diff_lambda(self.lambda_q1, self.lambda_k1, self.lambda_q2, self.lambda_k2)
        if self.local:
            o = att.local_diff_attention(q, kv, lam, self.subln.weight, H, W, h, hd, self.scale)
        else:
            # pooled tokens: 1x1 conv == per-token Linear; pooling on the tokens-major image view
            t = _Linear1x1.apply_conv(x, self.sr)
            if C % 4 == 0 and isinstance(self.act, nn.GELU):
                t = avgpool_tokens(t, H, W, self.pool_H, self.pool_W, gelu=True)     # GELU folded into the pooling read
            else:
                t = self.pool(self.act(t).transpose(1, 2).reshape(Bn, C, H, W)).flatten(2).transpose(1, 2)
            o = att.pooled_diff_attention(q, linear_tokens(layer_norm_tokens(t, self.norm), self.kv), lam,
                                          self.subln.weight, h, hd, self.scale)
        # LePE on the v half of the kv projection, read in place, with the attention output added in the same pass
        return dwconv3x3_tokens(v_local, self.lepe.weight, self.lepe.bias, H, W, residual=o)


class Attention(nn.Module):
    """Plain softmax attention + LePE; only built when sr_ratio == 1 (never in the shipped config, :787-821)."""

    def __init__(self, dim, input_resolution, num_heads=8, qkv_bias=True, attn_drop=0., proj_drop=0.):
        super().__init__()
        assert dim % num_heads == 0
        self.dim, self.num_heads, self.head_dim = dim, num_heads, dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.lepe = nn.Conv2d(dim, dim, 3, padding=1, groups=dim)

    def forward(self, x, H, W):
        Bn, N, C = x.shape
        q, k, v = self.qkv(x).reshape(Bn, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        o = F.scaled_dot_product_attention(q, k, v, dropout_p=self.attn_drop.p if self.training else 0.0)
        o = o.transpose(1, 2).reshape(Bn, N, C)
        vt = v.transpose(1, 2).reshape(Bn, N, C).contiguous()
        return o + dwconv3x3_tokens(vt, self.lepe.weight, self.lepe.bias, H, W)


class MLLABlock(nn.Module):
    def __init__(self, dim, input_resolution, num_heads, mlp_ratio=4., qkv_bias=True, drop=0., drop_path=0., sr_ratio=1,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, **kwargs):
        super().__init__()
        self.dim, self.input_resolution, self.num_heads, self.mlp_ratio = dim, input_resolution, num_heads, mlp_ratio
        self.norm1 = norm_layer(dim)
        self.in_proj = nn.Linear(dim, dim)
        self.act_proj = nn.Linear(dim, dim)
        self.dwc = nn.Conv2d(dim, dim, 3, padding=1, groups=dim)
        self.act = nn.SiLU()
        self.sr_ratio = sr_ratio
        if sr_ratio == 1:
            self.attn = Attention(dim=dim, input_resolution=input_resolution, num_heads=num_heads, qkv_bias=qkv_bias)
        else:
            self.attn = nn.ModuleList([
                AggregatedAttention(dim=dim // 2, input_resolution=input_resolution, num_heads=num_heads // 2,
                                    local=True, qkv_bias=qkv_bias, sr_ratio=sr_ratio),
                AggregatedAttention(dim=dim // 2, input_resolution=input_resolution, num_heads=num_heads // 2,
                                    local=False, qkv_bias=qkv_bias, sr_ratio=sr_ratio)])
        self.out_proj = nn.Linear(dim, dim)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward_tokens(self, t, H, W):
        """tokens-major (B, N, C) -> (B, N, C)"""
        t, shortcut = layer_norm_fork(t, self.norm1)          # shortcut == input; its gradient is added inside LN's backward
        gate = linear_tokens(t, self.act_proj)                  # SiLU applied inside the gate kernel below
        t = dwconv3x3_tokens(linear_tokens(t, self.in_proj), self.dwc.weight, self.dwc.bias, H, W, silu=True)
        if self.sr_ratio == 1:
            t = self.attn(t, H, W)
        else:
            a, b = torch.chunk(t, 2, dim=-1)
            t = torch.cat([self.attn[0](a, H, W), self.attn[1](b, H, W)], dim=-1)
        t = silu_gate(t, gate) if isinstance(self.act, nn.SiLU) else t * self.act(gate)
        t = residual_drop_path(shortcut, linear_tokens(t, self.out_proj), self.drop_path)
        n2, t = layer_norm_fork(t, self.norm2)
        return residual_drop_path(t, self.mlp(n2), self.drop_path)

    def forward(self, x):
        H, W = self.input_resolution
        Bn, C, h_, w_ = x.shape
        assert (H == h_) and (W == w_), "input feature has wrong size"
        t = self.forward_tokens(x.permute(0, 2, 3, 1).reshape(Bn, H * W, C), H, W)   # a view when x is channels_last
        return t.reshape(Bn, H, W, C).permute(0, 3, 1, 2)

    def extra_repr(self):
        return (f"dim={self.dim}, input_resolution={self.input_resolution}, num_heads={self.num_heads}, "
                f"mlp_ratio={self.mlp_ratio}")


class BasicLayer(nn.Module):
    def __init__(self, dim, input_resolution, depth, num_heads, mlp_ratio=4., qkv_bias=True, drop=0., drop_path=0.,
                 sr_ratio=1, norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False):
        super().__init__()
        self.dim, self.input_resolution, self.depth, self.use_checkpoint = dim, input_resolution, depth, use_checkpoint
        self.blocks = nn.ModuleList([
            MLLABlock(dim=dim, input_resolution=input_resolution, num_heads=num_heads, mlp_ratio=mlp_ratio,
                      qkv_bias=qkv_bias, drop=drop, drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                      sr_ratio=sr_ratio, norm_layer=norm_layer) for i in range(depth)])
        self.downsample = None if downsample is None else downsample(
            [input_resolution[0] * 2, input_resolution[1] * 2], dim=dim // 2)

    def forward(self, x):
        """NCHW -> NCHW; the blocks of a stage run back to back on the tokens-major layout."""
        if self.downsample is not None:
            x = self.downsample(x)
        H, W = self.input_resolution
        Bn, C, h_, w_ = x.shape
        assert (H == h_) and (W == w_), "input feature has wrong size"
        t = x.permute(0, 2, 3, 1).reshape(Bn, H * W, C)   # tokens-major == NHWC: a view when x is channels_last
        for blk in self.blocks:
            if self.use_checkpoint:
                t = torch.utils.checkpoint.checkpoint(blk.forward_tokens, t, H, W, use_reentrant=False)
            else:
                t = blk.forward_tokens(t, H, W)
        if t.requires_grad:
            t = GradContiguous.apply(t)      # the stage's backward runs on a tokens-major contiguous gradient
        return t.reshape(Bn, H, W, C).permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------------------------------------
# Conv stages: PyTorch / cuDNN, not the named hot path.  Parameter names follow the reference.
# ----------------------------------------------------------------------------------------------------------
class _TokensLN(nn.Module):
    """LayerNorm over channels of an NCHW map (the reference flattens, norms, and reshapes back).  On a channels_last
    CUDA map this is the tokens-major LayerNorm kernel on the same memory (no fp32 round trip under autocast)."""

    @staticmethod
    def apply(norm, x):
        t = x.permute(0, 2, 3, 1)
        if x.is_cuda and t.is_contiguous() and x.shape[1] % 4 == 0:
            return layer_norm_tokens(t, norm).permute(0, 3, 1, 2)
        return norm(t).permute(0, 3, 1, 2)


class project(nn.Module):
    def __init__(self, in_dim, out_dim, stride, padding, activate, norm, last=False):
        super().__init__()
        self.out_dim, self.last = out_dim, last
        self.conv1 = Conv2dCL(in_dim, out_dim, kernel_size=3, stride=stride, padding=padding)
        self.conv2 = Conv2dCL(out_dim, out_dim, kernel_size=3, stride=1, padding=1)
        self.activate = activate()
        self.norm1 = norm(out_dim)
        if not last:
            self.norm2 = norm(out_dim)

    def forward(self, x):
        x = _TokensLN.apply(self.norm1, self.activate(self.conv1(x)))
        x = self.conv2(x)
        if not self.last:
            x = _TokensLN.apply(self.norm2, self.activate(x))
        return x


class PatchEmbed(nn.Module):
    def __init__(self, patch_size=(2, 2), in_chans=4, embed_dim=96, norm_layer=None):
        super().__init__()
        self.patch_size, self.in_chans, self.embed_dim = patch_size, in_chans, embed_dim
        self.proj1 = project(in_chans, embed_dim // 2, [2, 2], 1, nn.GELU, nn.LayerNorm, False)
        self.proj2 = project(embed_dim // 2, embed_dim, [patch_size[0] // 2, patch_size[1] // 2], 1, nn.GELU,
                             nn.LayerNorm, True)
        self.norm = norm_layer(embed_dim) if norm_layer is not None else None

    def forward(self, x):
        _, _, H, W = x.size()
        if W % self.patch_size[1] != 0:
            x = F.pad(x, (0, self.patch_size[1] - W % self.patch_size[1]))
        if H % self.patch_size[0] != 0:
            x = F.pad(x, (0, 0, 0, self.patch_size[0] - H % self.patch_size[0]))
        x = self.proj2(self.proj1(x))
        if self.norm is not None:
            x = _TokensLN.apply(self.norm, x)
        return x


def _conv1x1_cl(conv, x):
    """1x1 (transposed) convolution of a channels_last map == per-token Linear: cuBLAS GEMM on the (B*H*W, C) view and
    the bias gradient through mlagg_colsum (torch's conv backward reduces the bias with its generic reduce kernel)."""
    from .ops import _Linear
    Bn, C, H, W = x.shape
    t = x.permute(0, 2, 3, 1)
    if not (x.is_cuda and t.is_contiguous() and conv.kernel_size == (1, 1) and conv.stride == (1, 1)
            and conv.padding == (0, 0) and conv.groups == 1):
        return conv(x)
    wgt = conv.weight.view(conv.weight.shape[0], conv.weight.shape[1])
    if isinstance(conv, nn.ConvTranspose2d):
        wgt = wgt.t()                                        # (in, out, 1, 1) -> (out, in)
    Co, bias = wgt.shape[0], conv.bias
    if Co % 8 != 0 and C % 8 == 0 and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16:
        # the 14-class segmentation heads: an output width that is not a multiple of 8 sends all three GEMMs to cuBLAS'
        # 2-byte-aligned kernels (0.9 ms per step for the two largest heads, tools/glue_sites.py / the launch list).
        # Zero rows up to the next multiple of 8 keep them on the tcgen05 kernels; the logits are the first Co columns of
        # the padded result (a strided view the loss kernel reads in place), and autograd's slice / pad gradients put
        # zeros into the padding on the way back.
        pad = 8 - Co % 8
        wgt = F.pad(wgt, (0, 0, 0, pad))
        bias = None if bias is None else F.pad(bias, (0, pad))
    y = _Linear.apply(t.reshape(Bn, H * W, C), wgt, bias)
    return y.reshape(Bn, H, W, -1)[..., :Co].permute(0, 3, 1, 2)


class MedNeXtBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, exp_r: int = 4, kernel_size: int = 7, do_res: int = True,
                 norm_type: str = "group", n_groups=None, dim="3d", grn=False):
        super().__init__()
        assert dim == "2d" and norm_type == "group" and not grn, "MLAgg-UNet builds the 2-D GroupNorm variant"
        self.do_res, self.dim, self.grn = do_res, dim, grn
        self.conv1 = Conv2dCL(in_channels, in_channels, kernel_size, stride=1, padding=kernel_size // 2,
                               groups=in_channels if n_groups is None else n_groups)
        self.norm = nn.GroupNorm(num_groups=in_channels, num_channels=in_channels)
        self.conv2 = nn.Conv2d(in_channels, exp_r * in_channels, kernel_size=1)
        self.act = nn.GELU()
        self.conv3 = nn.Conv2d(exp_r * in_channels, out_channels, kernel_size=1)

    def forward(self, x, dummy_tensor=None):
        t = _inst_norm(self.norm, self.conv1(x))
        tk = t.permute(0, 2, 3, 1)
        if t.is_cuda and tk.is_contiguous() and isinstance(self.act, nn.GELU):
            # the 1x1-conv pair on the channels_last map == fc1 -> GELU -> fc2 on its tokens (one fused node)
            Bn, C, H, W = t.shape
            w2, w3 = self.conv2.weight, self.conv3.weight
            y = mlp_gelu_tokens(tk.reshape(Bn, H * W, C), w2.view(w2.shape[0], w2.shape[1]), self.conv2.bias,
                                w3.view(w3.shape[0], w3.shape[1]), self.conv3.bias)
            y = y.reshape(Bn, H, W, -1).permute(0, 3, 1, 2)
        else:
            y = _conv1x1_cl(self.conv3, self.act(_conv1x1_cl(self.conv2, t)))
        return x + y if self.do_res else y


class MedNeXtDownBlock(MedNeXtBlock):
    def __init__(self, in_channels, out_channels, exp_r=4, kernel_size=7, do_res=False, norm_type="group", dim="3d",
                 grn=False):
        super().__init__(in_channels, out_channels, exp_r, kernel_size, do_res=False, norm_type=norm_type, dim=dim,
                         grn=grn)
        self.resample_do_res = do_res
        if do_res:
            self.res_conv = Conv2dCL(in_channels, out_channels, kernel_size=1, stride=2)
        self.conv1 = Conv2dCL(in_channels, in_channels, kernel_size, stride=2, padding=kernel_size // 2,
                               groups=in_channels)

    def forward(self, x, dummy_tensor=None):
        y = super().forward(x)
        return y + self.res_conv(x) if self.resample_do_res else y


class PatchExpand(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 7, norm_type: str = "group", dim="3d",
                 do_res=False):
        super().__init__()
        assert dim == "2d" and norm_type == "group"
        self.resample_do_res, self.dim = do_res, dim
        if do_res:
            self.res_conv = ConvTranspose2dCL(in_channels, out_channels, kernel_size=1, stride=2)
        self.conv1 = ConvTranspose2dCL(in_channels, out_channels, kernel_size, stride=2, padding=kernel_size // 2)
        self.norm = nn.GroupNorm(num_groups=in_channels, num_channels=in_channels)

    def forward(self, x, dummy_tensor=None):
        y = self.conv1(_inst_norm(self.norm, x))
        r = self.res_conv(x) if self.resample_do_res else None
        if y.is_cuda and y.is_contiguous(memory_format=torch.channels_last) and (
                r is None or (r.is_contiguous(memory_format=torch.channels_last) and r.dtype == y.dtype)):
            return PadTopLeftAdd.apply(y, r)
        y = F.pad(y, (1, 0, 1, 0))
        return y if r is None else y + F.pad(r, (1, 0, 1, 0))


class OutBlock(nn.Module):
    def __init__(self, in_channels, n_classes, dim):
        super().__init__()
        assert dim == "2d"
        self.conv_out = nn.ConvTranspose2d(in_channels, n_classes, kernel_size=1)

    def forward(self, x, dummy_tensor=None):
        return _conv1x1_cl(self.conv_out, x)


class MLLA_Enc(nn.Module):
    def __init__(self, img_size=224, patch_size=4, in_chans=3, num_classes=1000, embed_dim=96, depths=[2, 2, 6, 2],
                 num_heads=[3, 6, 12, 24], mlp_ratio=4., qkv_bias=True, drop_rate=0., drop_path_rate=0.1,
                 sr_ratio=[8, 4, 2, 1], norm_layer=nn.LayerNorm, ape=False, use_checkpoint=False, **kwargs):
        super().__init__()
        self.num_classes, self.num_layers, self.embed_dim, self.ape = num_classes, len(depths), embed_dim, ape
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.mlp_ratio = mlp_ratio
        self.patch_size = [patch_size, patch_size]
        self.patch_norm = False
        self.patch_embed = PatchEmbed(patch_size=self.patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                      norm_layer=norm_layer if self.patch_norm else None)
        res = [img_size // patch_size] * 2 if isinstance(img_size, int) else [i // patch_size for i in img_size]
        self.patches_resolution = res
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList([
            BasicLayer(dim=int(embed_dim * 2 ** i), input_resolution=(res[0] // (2 ** i), res[1] // (2 ** i)),
                       depth=depths[i], num_heads=num_heads[i], mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, drop=drop_rate,
                       drop_path=dpr[sum(depths[:i]):sum(depths[:i + 1])], sr_ratio=sr_ratio[i], norm_layer=norm_layer,
                       downsample=None, use_checkpoint=use_checkpoint) for i in range(self.num_layers)])
        self.downs = nn.ModuleList([
            MedNeXtDownBlock(in_channels=int(embed_dim * 2 ** i), out_channels=int(embed_dim * 2 ** (i + 1)),
                             exp_r=mlp_ratio, kernel_size=3, do_res=True, norm_type="group", dim="2d")
            for i in range(self.num_layers - 1)])

    def forward_features(self, x, normalize=True):
        outs = [x]
        x = self.patch_embed(x)
        for i, layer in enumerate(self.layers):
            x = layer(x)
            outs.append(x)
            if i < self.num_layers - 1:
                x = self.downs[i](x)
        return outs

    def forward(self, x, normalize=True):
        return self.forward_features(x, normalize=normalize)


class MLLA_Uper(nn.Module):
    def __init__(self, img_size, patch_size, in_channels: int, out_channels: int, embed_dim: int = 96,
                 depths: Sequence[int] = (2, 2, 2, 2), num_heads: Sequence[int] = (3, 6, 12, 24), mlp_ratio=4,
                 qkv_bias=True, drop_rate: float = 0.0, attn_drop_rate: float = 0.0, dropout_path_rate: float = 0.0,
                 sr_ratio=[8, 4, 2, 1], normalize: bool = True, norm_layer=nn.LayerNorm, ape=False,
                 use_checkpoint: bool = False, spatial_dims: str = "2d", norm_type: str = "group", do_res: bool = True,
                 deep_supervision: bool = True):
        super().__init__()
        self.normalize, self.deep_supervision = normalize, deep_supervision
        E = embed_dim
        self.mlla = MLLA_Enc(img_size=img_size, patch_size=patch_size, in_chans=in_channels, num_classes=out_channels,
                             embed_dim=E, depths=depths, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                             drop_rate=drop_rate, drop_path_rate=dropout_path_rate, sr_ratio=sr_ratio,
                             norm_layer=norm_layer, ape=ape, use_checkpoint=use_checkpoint)
        self.mambaskip = VSS_Conv_Layer([E, E * 2, E * 4, E * 8], E // 2, depth=1, drop_path=0.1, use_checkpoint=False)
        up = lambda cin, cout: PatchExpand(in_channels=cin, out_channels=cout, kernel_size=3, do_res=do_res,
                                           norm_type=norm_type, dim=spatial_dims)
        dec = lambda c, n: nn.Sequential(*[MedNeXtBlock(in_channels=c, out_channels=c, exp_r=mlp_ratio, kernel_size=3,
                                                        do_res=do_res, norm_type=norm_type, dim=spatial_dims, grn=False)
                                           for _ in range(n)])
        self.up_2, self.dec_block_2 = up(8 * E, 4 * E), dec(4 * E, depths[-2])
        self.up_1, self.dec_block_1 = up(4 * E, 2 * E), dec(2 * E, depths[-3])
        self.up_0, self.dec_block_0 = up(2 * E, E), dec(E, depths[-4])
        self.encoder0 = UnetrBasicBlock(spatial_dims=2, in_channels=in_channels, out_channels=E // 2, kernel_size=3,
                                        stride=1, norm_name="instance", res_block=True)
        self.decoder0 = UnetrUpBlock(spatial_dims=2, in_channels=E, out_channels=E // 2, kernel_size=3,
                                     upsample_kernel_size=2, norm_name="instance", res_block=True)
        self.out_0 = OutBlock(in_channels=E // 2, n_classes=out_channels, dim=spatial_dims)
        # kept for checkpoint compatibility; never used in forward, so frozen (DDP would otherwise
        # fail with "expected to have finished reduction" on the second iteration)
        self.dummy_tensor = nn.Parameter(torch.tensor([1.]), requires_grad=False)
        if deep_supervision:
            self.out_1 = OutBlock(in_channels=E, n_classes=out_channels, dim=spatial_dims)
            self.out_2 = OutBlock(in_channels=E * 2, n_classes=out_channels, dim=spatial_dims)
            self.out_3 = OutBlock(in_channels=E * 4, n_classes=out_channels, dim=spatial_dims)
            self.out_4 = OutBlock(in_channels=E * 8, n_classes=out_channels, dim=spatial_dims)

    def forward(self, x_in):
        # the whole network runs channels_last (NHWC): the conv stages get cuDNN's native layout and the token blocks
        # read the same memory as (B, N, C) without a copy
        x_in = x_in.contiguous(memory_format=torch.channels_last)
        hs = self.mlla(x_in, self.normalize)
        hs[1:] = self.mambaskip(hs[1:])
        ds = self.deep_supervision
        x_ds_4 = self.out_4(hs[4]) if ds else None
        x = self.dec_block_2(hs[3] + self.up_2(hs[4]))
        x_ds_3 = self.out_3(x) if ds else None
        x = self.dec_block_1(hs[2] + self.up_1(x))
        x_ds_2 = self.out_2(x) if ds else None
        x = self.dec_block_0(hs[1] + self.up_0(x))
        x_ds_1 = self.out_1(x) if ds else None
        x = self.out_0(self.decoder0(x, self.encoder0(hs[0])))
        return [x, x_ds_1, x_ds_2, x_ds_3, x_ds_4] if ds else x
