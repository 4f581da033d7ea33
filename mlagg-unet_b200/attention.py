"""Attention cores of the MLAgg block and of MLLA linear attention, tokens-major.

  local_diff_attention   3x3-window differential softmax attention + sub-LN  (reference
                         nnUNetTrainer_MLAgg_2D_dt_MS.py:693-717)
  pooled_diff_attention  differential softmax attention over P pooled tokens + sub-LN (:732-760), i.e. the four
                         flash_attn_func calls + cat + lambda-combine + RMSNorm of the reference in one op,
                         INCLUDING flash-attn's second head_dim**-0.5 scale
  linear_attention_core  elu+1 / RoPE / per-head K^T V state / normaliser (nnUNetTrainer_MLLA_UNet.py:234-246)

STATUS (round 1): these three are compositions of torch CUDA ops (cuBLAS batched GEMM + elementwise), written
from the math in App. A.4/A.5 rather than from the reference's op sequence.  They are the slots the fused
sm_100a kernels `local_diffattn`, `pooled_diffattn`, `linattn_state/apply` plug into;
DESIGN.md lists them as "not yet native".  Depthwise convs and the scan around them already are native.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LAMBDA_INIT = 0.8


def rmsnorm_affine(x, weight, eps):
    y = x.float()
    y = y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + eps)
    return y.type_as(x) * weight


def diff_lambda(lq1, lk1, lq2, lk2):
    return torch.exp(torch.sum(lq1 * lk1).float()) - torch.exp(torch.sum(lq2 * lk2).float()) + LAMBDA_INIT


def _shifted_neighbours(t, H, W):
    """t (B, H*W, ...) -> (B, H*W, 9, ...): the 3x3 neighbourhood, row-major over (dr, dc); zeros outside."""
    Bn = t.shape[0]
    rest = t.shape[2:]
    img = t.reshape(Bn, H, W, -1)
    pad = F.pad(img, (0, 0, 1, 1, 1, 1))
    nb = torch.stack([pad[:, 1 + dr:1 + dr + H, 1 + dc:1 + dc + W] for dr in (-1, 0, 1) for dc in (-1, 0, 1)], dim=3)
    return nb.reshape(Bn, H * W, 9, *rest)


_MASKS = {}


def _border_mask(H, W, device):
    key = (H, W, str(device))
    if key not in _MASKS:
        r = torch.arange(H, device=device).view(H, 1, 1, 1) + torch.tensor([-1, 0, 1], device=device).view(1, 1, 3, 1)
        c = torch.arange(W, device=device).view(1, W, 1, 1) + torch.tensor([-1, 0, 1], device=device).view(1, 1, 1, 3)
        ok = (r >= 0) & (r < H) & (c >= 0) & (c < W)
        _MASKS[key] = ~ok.reshape(H * W, 9)
    return _MASKS[key]


def local_diff_attention(q, k, v, lam, subln_w, H, W):
    """q (B,N,2h,hd) pre-scaled by hd**-0.5; k (B,N,2h,hd); v (B,N,h,2hd) -> (B,N,h*2hd)."""
    Bn, N, h2, hd = q.shape
    h = h2 // 2
    kn = _shifted_neighbours(k, H, W)                    # (B,N,9,2h,hd)
    vn = _shifted_neighbours(v, H, W)                    # (B,N,9,h,2hd)
    logits = torch.einsum("bnjd,bnpjd->bnjp", q, kn)
    logits = logits.masked_fill(_border_mask(H, W, q.device)[None, :, None, :], float("-inf"))
    a = logits.softmax(-1).view(Bn, N, h, 2, 9)
    a = a[:, :, :, 0] - lam.to(a.dtype) * a[:, :, :, 1]
    o = torch.einsum("bnmp,bnpmd->bnmd", a, vn)
    o = rmsnorm_affine(o, subln_w, 1e-5) * (1 - LAMBDA_INIT)
    return o.reshape(Bn, N, h * 2 * hd)


def pooled_diff_attention(q, kp, vp, lam, subln_w):
    """q (B,N,h,2,hd) pre-scaled once; kp (B,P,h,2,hd); vp (B,P,h,2hd) -> (B,N,h*2hd)."""
    Bn, N, h, _, hd = q.shape
    logits = torch.einsum("bnmjd,bpmjd->bmjnp", q, kp) * (hd ** -0.5)   # flash_attn's own scale (F4)
    a = logits.float().softmax(-1).to(q.dtype)
    o = torch.einsum("bmjnp,bpmd->bnmjd", a, vp)
    o = o[:, :, :, 0] - lam.to(o.dtype) * o[:, :, :, 1]
    o = rmsnorm_affine(o, subln_w, 1e-5) * (1 - LAMBDA_INIT)
    return o.reshape(Bn, N, h * 2 * hd)


_ROPE = {}


def rope_tables(H, W, C, device, base=10000.0):
    """cos/sin (H*W, C/2): pair i<C/4 rotates by row*theta_i, the next C/4 by col*theta_i (MLLA_UNet.py:172-188)."""
    key = (H, W, C, str(device))
    if key not in _ROPE:
        k = C // 4
        theta = 1.0 / (base ** (torch.arange(k, device=device, dtype=torch.float32) / k))
        r = torch.arange(H, device=device, dtype=torch.float32).view(H, 1, 1) * theta
        c = torch.arange(W, device=device, dtype=torch.float32).view(1, W, 1) * theta
        ang = torch.cat([r.expand(H, W, k), c.expand(H, W, k)], dim=-1).reshape(H * W, 2 * k)
        _ROPE[key] = (torch.cos(ang), torch.sin(ang))
    return _ROPE[key]


def rope_apply(x, H, W):
    """x (B, H*W, C) -> rotated fp32 tensor (the reference forces fp32)."""
    Bn, N, C = x.shape
    cs, sn = rope_tables(H, W, C, x.device)
    x = x.float().reshape(Bn, N, C // 2, 2)
    re = cs * x[..., 0] - sn * x[..., 1]
    im = sn * x[..., 0] + cs * x[..., 1]
    return torch.stack([re, im], dim=-1).reshape(Bn, N, C)


def linear_attention_core(q, k, v, H, W, num_heads):
    """q, k (B,N,C) raw projections; v (B,N,C) -> (B,N,C) fp32: phi = elu+1, RoPE, state, normaliser."""
    Bn, N, C = q.shape
    hd = C // num_heads
    q, k = F.elu(q) + 1.0, F.elu(k) + 1.0
    heads = lambda t: t.reshape(Bn, N, num_heads, hd).transpose(1, 2)
    qr, kr = heads(rope_apply(q, H, W)), heads(rope_apply(k, H, W))
    qh, kh, vh = heads(q), heads(k), heads(v)
    z = 1.0 / (torch.einsum("bhnd,bhd->bhn", qh, kh.mean(dim=2)) + 1e-6)
    state = torch.einsum("bhnd,bhne->bhde", kr * N ** -0.5, (vh * N ** -0.5).to(kr.dtype))
    o = torch.einsum("bhnd,bhde->bhne", qr, state) * z[..., None]
    return o.transpose(1, 2).reshape(Bn, N, C)
