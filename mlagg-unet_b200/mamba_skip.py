"""Multi-Scale Mamba Module (MSMM) of the skip connections -- drop-in for the reference's
`mlagg/nnunetv2/training/nnUNetTrainer/variants/mamba/MambaSkip.py`:

    SS2D_skip        (:266-543)   same ctor arguments, parameter names and shapes (state_dict compatible)
    DWConv           (:545-556)
    ConvolutionalGLU (:559-577)
    VSS_Conv_Block   (:669-753)
    VSS_Conv_Layer   (:756-804)

What differs is HOW the forward runs (B200-first):
  * activations stay tokens-major (B, L, C); the per-stage depthwise conv + SiLU is one sm_100a stencil kernel
    on that layout (ops.dwconv3x3_tokens) instead of permute -> cuDNN -> permute;
  * the four direction-specific x_proj matrices are applied ONCE to the un-permuted tokens as a single
    (L x 96) @ (96 x 140) GEMM -- a permutation along L commutes with a per-token projection (App. A.3);
  * the 4-direction multi-scale cross-scan / cross-merge are index maps (gather / inverse gather);
  * the selective scan is the sm_100a kernel behind `selective_scan_fn` (fp32 state, like the reference).
"""
from __future__ import annotations

import math
from functools import lru_cache, partial
from typing import Callable

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .ops import CatStages, Conv2dCL, JoinLast, SplitLast, SplitStages, _Linear, conv_glu_core, dwconv3x3_stages, dwconv3x3_tokens, layer_norm_fork, layer_norm_tokens, linear_tokens, residual_drop_path
from .selective_scan_interface import msmm_scan, msmm_scan_tokens, selective_scan_fn, xdbl_pad
from .thirdparty_shims import DropPath, _inst_norm


@lru_cache(maxsize=64)
def _scan_maps_cpu(hw: tuple):
    """(idx, inv): LongTensors (4, L_cat).  Position l of direction k reads token idx[k, l] of the row-major,
    stage-concatenated sequence; inv[k] is the inverse permutation (reference MambaSkip.py:414-422, :454-471)."""
    cols = [[], [], [], []]
    off = 0
    for (H, W) in hw:
        n = H * W
        l = torch.arange(n)
        rm = l
        cm = (l % H) * W + torch.div(l, H, rounding_mode="floor")
        for k, m in enumerate((rm, cm, rm.flip(0), cm.flip(0))):
            cols[k].append(off + m)
        off += n
    idx = torch.stack([torch.cat(c) for c in cols])
    inv = torch.empty_like(idx)
    ar = torch.arange(idx.shape[1])
    for k in range(4):
        inv[k, idx[k]] = ar
    return idx, inv


_MAP_CACHE = {}


def cross_scan_maps(hw, device):
    key = (tuple((int(h), int(w)) for h, w in hw), str(device))
    if key not in _MAP_CACHE:
        idx, inv = _scan_maps_cpu(key[0])
        _MAP_CACHE[key] = (idx.to(device), inv.to(device))
    return _MAP_CACHE[key]


class SS2D_skip(nn.Module):
    def __init__(self, stage_num, d_model, d_state=16, d_conv=3, expand=2, dt_rank="auto", dt_min=0.001,
                 dt_max=0.1, dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, dropout=0., conv_bias=True,
                 bias=False, device=None, dtype=None, **kwargs):
        fk = {"device": device, "dtype": dtype}
        super().__init__()
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        assert d_conv == 3, "the tokens-major stencil kernel is 3x3"
        K, R, N, Di = 4, self.dt_rank, d_state, self.d_inner

        self.in_proj = nn.Linear(d_model, Di, bias=bias, **fk)
        self.conv2d = nn.ModuleList([nn.Conv2d(Di, Di, d_conv, padding=(d_conv - 1) // 2, groups=Di, bias=conv_bias,
                                               **fk) for _ in range(stage_num)])
        self.act = nn.ModuleList([nn.SiLU() for _ in range(stage_num)])
        self.x_proj_weight = nn.Parameter(torch.stack(
            [nn.Linear(Di, R + 2 * N, bias=False, **fk).weight for _ in range(K)], dim=0))      # (4, R+2N, Di)
        dts = [self.dt_init(R, Di, dt_scale, dt_init, dt_min, dt_max, dt_init_floor, **fk) for _ in range(K)]
        self.dt_projs_weight = nn.Parameter(torch.stack([t.weight for t in dts], dim=0))         # (4, Di, R)
        self.dt_projs_bias = nn.Parameter(torch.stack([t.bias for t in dts], dim=0))             # (4, Di)
        self.A_logs = self.A_log_init(N, Di, copies=K, merge=True)                               # (4*Di, N)
        self.Ds = self.D_init(Di, copies=K, merge=True)                                          # (4*Di,)
        self.out_norm = nn.LayerNorm(Di)
        self.out_proj = nn.Linear(Di, d_model, bias=bias, **fk)
        self.dropout = nn.Dropout(dropout) if dropout > 0. else None

    # ---- initialisers (reference :348-403)
    @staticmethod
    def dt_init(dt_rank, d_inner, dt_scale=1.0, dt_init="random", dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4, **fk):
        proj = nn.Linear(dt_rank, d_inner, bias=True, **fk)
        std = dt_rank ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(proj.weight, std)
        elif dt_init == "random":
            nn.init.uniform_(proj.weight, -std, std)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(d_inner, **fk) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min))
        dt = dt.clamp(min=dt_init_floor)
        with torch.no_grad():
            proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))  # softplus^-1(dt)
        proj.bias._no_reinit = True
        return proj

    @staticmethod
    def A_log_init(d_state, d_inner, copies=1, device=None, merge=True):
        A_log = torch.log(torch.arange(1, d_state + 1, dtype=torch.float32, device=device)).repeat(d_inner, 1)
        if copies > 1:
            A_log = A_log.unsqueeze(0).repeat(copies, 1, 1)
            if merge:
                A_log = A_log.flatten(0, 1)
        A_log = nn.Parameter(A_log.contiguous())
        A_log._no_weight_decay = True
        return A_log

    @staticmethod
    def D_init(d_inner, copies=1, device=None, merge=True):
        D = torch.ones(d_inner, device=device)
        if copies > 1:
            D = D.unsqueeze(0).repeat(copies, 1)
            if merge:
                D = D.flatten(0, 1)
        D = nn.Parameter(D.contiguous())
        D._no_weight_decay = True
        return D

    # ---- core: tokens (B, L, Di) -> merged scan output (B, L, Di) fp32
    @staticmethod
    def _stagewise_transpose(t, hw, to_col):
        """(B, D, L): per stage, row-major (H, W) order <-> column-major order (the W x H transposed image)."""
        parts, off = [], 0
        Bn, Dn, _ = t.shape
        for (H, W) in hw:
            a, b_ = (H, W) if to_col else (W, H)
            parts.append(t[:, :, off:off + H * W].reshape(Bn, Dn, a, b_).transpose(2, 3).reshape(Bn, Dn, H * W))
            off += H * W
        return torch.cat(parts, dim=-1)

    def forward_core_tokens(self, xc, hw):
        """Fused path: the cross-scan, dt projection and un-permutation of forward_corev0 (reference :405-473) happen
        inside the scan kernels' operand addressing; the walk orders of x / x_dbl and the cross-merge are single
        tile-transpose kernels (csrc/walk.cu); torch supplies ONE x_proj GEMM on the tokens (a permutation along L
        commutes with a per-token projection, App. A.3)."""
        R, N = self.dt_rank, self.d_state
        Wx = self.x_proj_weight                                                 # (4, R+2N, Di)
        pad = Wx.new_zeros(xdbl_pad(R + 2 * N) - 2 * (R + 2 * N), Wx.shape[2])
        W_all = torch.cat([Wx[0], Wx[2], pad, Wx[1], Wx[3], pad], dim=0)        # [row walk: dirs 0, 2 | column walk: 1, 3]
        xdbl = _Linear.apply(xc, W_all, None)                                   # (B, L, 2 P), autocast dtype
        return msmm_scan_tokens(xc, xdbl, self.dt_projs_weight.reshape(-1, R), self.dt_projs_bias.reshape(-1),
                                -torch.exp(self.A_logs.float()), self.Ds, hw)

    def forward_core_planes(self, xc, hw):
        """Same computation with torch building the walk planes (transpose / per-stage reshape / cat) around
        `msmm_scan`; kept as the parity partner of `forward_core_tokens` in the tests."""
        Bn, L, Di = xc.shape
        R, N = self.dt_rank, self.d_state
        xrow = xc.transpose(1, 2).contiguous()                                  # (B, Di, L) row-major walk
        xcol = self._stagewise_transpose(xrow, hw, to_col=True)                 # column-major walk
        Wx = self.x_proj_weight                                                 # (4, R+2N, Di)
        # W_x[k] @ xs[k] == permute_k(W_x[k] @ x): directions {0,2} on the row walk, {1,3} on the column walk (App. A.3)
        xdbl_row = torch.matmul(Wx[0::2].reshape(2 * (R + 2 * N), Di), xrow).view(Bn, 2, R + 2 * N, L)
        xdbl_col = torch.matmul(Wx[1::2].reshape(2 * (R + 2 * N), Di), xcol).view(Bn, 2, R + 2 * N, L)
        out = msmm_scan(xrow, xcol, xdbl_row, xdbl_col, self.dt_projs_weight.reshape(4 * Di, R),
                        self.dt_projs_bias.reshape(-1), -torch.exp(self.A_logs.float()), self.Ds,
                        [h * w for h, w in hw])
        assert out.dtype == torch.float32
        y = out[:, 0] + out[:, 2] + self._stagewise_transpose(out[:, 1] + out[:, 3], hw, to_col=False)
        return y.transpose(1, 2).contiguous()

    def forward_core_tokens_unfused(self, xc, hw):
        """Mamba-interface path (materialised cross-scan through index maps + selective_scan_fn); kept for parity
        tests of the fused path and for callers that want the reference's op boundary."""
        Bn, L, Di = xc.shape
        K, R, N = 4, self.dt_rank, self.d_state
        idx, inv = cross_scan_maps(hw, xc.device)
        x_dbl = F.linear(xc, self.x_proj_weight.view(K * (R + 2 * N), Di)).view(Bn, L, K, R + 2 * N)
        dts_r, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=-1)
        # dt projection has inner dimension R = dt_rank (3): as a GEMM it is degenerate (K=3), so it is written as
        # R broadcast multiply-adds; result (B, 4, Di, L) un-permuted, channels-major like the scan wants it
        dts = None
        for r_ in range(R):
            term = dts_r[..., r_].permute(0, 2, 1).unsqueeze(2) * self.dt_projs_weight[:, :, r_].to(dts_r.dtype)[None, :, :, None]
            dts = term if dts is None else dts + term
        u = xc.transpose(1, 2)                                                      # (B, Di, L) view
        Bs, Cs = Bs.permute(0, 2, 3, 1), Cs.permute(0, 2, 3, 1)                     # (B, 4, N, L) views
        gather = lambda t, k: t.index_select(-1, idx[k]).float()
        us = torch.stack([gather(u, k) for k in range(K)], dim=1).view(Bn, K * Di, L)
        dl = torch.stack([gather(dts[:, k], k) for k in range(K)], dim=1).view(Bn, K * Di, L)
        Bg = torch.stack([gather(Bs[:, k], k) for k in range(K)], dim=1)
        Cg = torch.stack([gather(Cs[:, k], k) for k in range(K)], dim=1)
        out = selective_scan_fn(us, dl, -torch.exp(self.A_logs.float()), Bg, Cg, self.Ds.float(), z=None,
                                delta_bias=self.dt_projs_bias.float().view(-1), delta_softplus=True,
                                return_last_state=False).view(Bn, K, Di, L)
        assert out.dtype == torch.float32
        y = out[:, 0]
        for k in range(1, K):
            y = y + out[:, k].index_select(-1, inv[k])
        return y.transpose(1, 2).contiguous()

    def forward(self, x, B, H, W, L_split, **kwargs):
        """x (B, L_cat, d_model), stages concatenated fine -> coarse along L."""
        hw = list(zip(H, W))
        x = linear_tokens(x, self.in_proj)
        y = self.forward_core_tokens(dwconv3x3_stages(x, hw, self.conv2d, silu=True), hw)
        assert y.dtype == torch.float32
        out = linear_tokens(layer_norm_tokens(y, self.out_norm), self.out_proj)
        return self.dropout(out) if self.dropout is not None else out


class DWConv(nn.Module):
    def __init__(self, dim=768):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, kernel_size=3, stride=1, padding=1, bias=True, groups=dim)

    def forward(self, x, H, W, silu=False):
        return dwconv3x3_tokens(x, self.dwconv.weight, self.dwconv.bias, H, W, silu=silu)   # x may be a channel slice


class ConvolutionalGLU(nn.Module):
    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = int(2 * (hidden_features or in_features) / 3)
        self.fc1 = nn.Linear(in_features, hidden_features * 2)
        self.dwconv = DWConv(hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x, H, W):
        h = linear_tokens(x, self.fc1)
        if isinstance(self.act, nn.SiLU):
            # conv + SiLU + the product with the v half in one kernel, both halves of h read in place
            av = conv_glu_core(h, self.dwconv.dwconv.weight, self.dwconv.dwconv.bias, H, W, silu=True)
        else:
            a, v = h.chunk(2, dim=-1)
            av = self.act(self.dwconv(a, H, W)) * v
        return self.drop(linear_tokens(self.drop(av), self.fc2))


class VSS_Conv_Block(nn.Module):
    def __init__(self, feature_dims, hidden_dim: int = 0, drop_path: float = 0,
                 norm_layer: Callable[..., nn.Module] = partial(nn.LayerNorm, eps=1e-6), attn_drop_rate: float = 0,
                 d_state: int = 16, ssm_ratio: int = 2., **kwargs):
        super().__init__()
        self.feature_dims, self.hidden_dim = feature_dims, hidden_dim
        self.ln_1 = norm_layer(hidden_dim)
        self.self_attention = SS2D_skip(stage_num=len(feature_dims), d_model=hidden_dim, d_state=d_state,
                                        expand=ssm_ratio, dropout=attn_drop_rate, **kwargs)
        self.drop_path = DropPath(drop_path)
        self.norm2 = norm_layer(hidden_dim)
        self.mlps = nn.ModuleList([ConvolutionalGLU(in_features=hidden_dim, hidden_features=int(hidden_dim * 4),
                                                    act_layer=nn.SiLU) for _ in feature_dims])
        self.conv_dims = [d - hidden_dim for d in feature_dims]
        self.conv_branches = nn.ModuleList([
            nn.Sequential(Conv2dCL(cd, cd, kernel_size=3, stride=1, padding=1), nn.InstanceNorm2d(cd, affine=True),
                          nn.SiLU()) for cd in self.conv_dims])

    def forward(self, inputs):
        """inputs: list of (B, C_s, H_s, W_s), fine -> coarse; returns the same shapes."""
        Bn = inputs[0].shape[0]
        H = [t.shape[2] for t in inputs]
        W = [t.shape[3] for t in inputs]
        L_split = [h * w for h, w in zip(H, W)]
        hd = self.hidden_dim
        # NHWC views of the inputs (free when they are channels_last), split into the first hd channels -> tokens-major
        # (B, L, hd) for the scan branch, and the rest -> conv branch
        halves = [SplitLast.apply(t.permute(0, 2, 3, 1), hd) for t in inputs]

        def conv_branch(s):                                                            # Conv2d, InstanceNorm2d, SiLU
            br = self.conv_branches[s]
            return _inst_norm(br[1], br[0](halves[s][1].permute(0, 3, 1, 2)), "silu").permute(0, 2, 3, 1)   # NHWC view

        # MLAGG_BRANCH_STREAM=1 (experimental): the conv branches do not depend on the scan path and can run on a second,
        # lower-priority stream next to it (the scan kernels occupy 120 of the 148 SMs)
        side = _lib.branch_stream(inputs[0].device) if (inputs[0].is_cuda and self.training) else None
        cbs = None
        if side is not None:
            main = torch.cuda.current_stream(inputs[0].device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                cbs = [conv_branch(s) for s in range(len(inputs))]
        m = CatStages.apply(*[a.flatten(1, 2) for a, _ in halves])         # (B, H, W, hd) -> (B, H W, hd): views
        n1, m = layer_norm_fork(m, self.ln_1)                # m's residual-path gradient is added inside ln_1's backward
        m = residual_drop_path(m, self.self_attention(n1, Bn, H, W, L_split), self.drop_path)
        # one packed copy per stage: with a uniform row stride the fc1 GEMM keeps its bias epilogue, the residual is the
        # one-pass kernel, and the Linear backward needs no re-pack of its saved input
        stage_tokens = SplitStages.apply(layer_norm_tokens(m, self.norm2), tuple(L_split))
        outs = []
        if side is not None:
            torch.cuda.current_stream(inputs[0].device).wait_stream(side)
        for s, t in enumerate(inputs):
            ms = stage_tokens[s]
            ms = residual_drop_path(ms, self.mlps[s](ms, H[s], W[s]), self.drop_path)
            cb = cbs[s] if cbs is not None else conv_branch(s)
            outs.append(JoinLast.apply(ms.reshape(Bn, H[s], W[s], hd), cb).permute(0, 3, 1, 2))      # channels_last
        return outs


class VSS_Conv_Layer(nn.Module):
    def __init__(self, feature_dims, hidden_dim, depth=1, attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm,
                 use_checkpoint=False, d_state=16, ssm_ratio=2., **kwargs):
        super().__init__()
        self.hidden_dim, self.use_checkpoint = hidden_dim, use_checkpoint
        self.blocks = nn.ModuleList([
            VSS_Conv_Block(feature_dims=feature_dims, hidden_dim=hidden_dim,
                           drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                           norm_layer=norm_layer, attn_drop_rate=attn_drop, d_state=d_state, ssm_ratio=ssm_ratio)
            for i in range(depth)])

    def forward(self, x):
        for blk in self.blocks:
            x = torch.utils.checkpoint.checkpoint(blk, x, use_reentrant=False) if self.use_checkpoint else blk(x)
        return x
