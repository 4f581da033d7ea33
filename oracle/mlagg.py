"""oracle/mlagg.py -- TEST INFRASTRUCTURE ONLY (parity oracle / CPU baseline).

Functional CPU restatement of the as-shipped MLAgg encoder block:
  * `RMSNorm`                       nnUNetTrainer_MLAgg_2D_dt_MS.py:592-610
  * `AggregatedAttention` local     :687-717, :779-782   (neighbour gather instead of nn.Unfold)
  * `AggregatedAttention` pooled    :718-760, :779-782   (flash_attn_func replaced by its definition,
                                     *including* the second head_dim**-0.5 scale)
  * `MLLABlock.forward`             :877-911
  * `Mlp`                           :176-192
Math per SURVEY.md App. A.4.  Parameters: flat dict with the reference's state_dict names.
`flash_attn` is un-vendored (README.md:53-56 of the reference); its published definition
softmax(q k^T / sqrt(d)) v is what is restated.

Pinned against the reference module source executed in this container
(tests/golden/make_golden.py -> tests/golden/mlagg_*.pt).
Nothing under mlagg-unet_b200/ imports this module.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from .msmm import dwconv3x3_tokens

LAMBDA_INIT = 0.8


def rmsnorm(x, weight, eps):
    y = x.float()
    y = y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + eps)
    return y.type_as(x) * weight


def diff_lambda(p, prefix=""):
    l1 = torch.exp(torch.sum(p[prefix + "lambda_q1"] * p[prefix + "lambda_k1"]).float())
    l2 = torch.exp(torch.sum(p[prefix + "lambda_q2"] * p[prefix + "lambda_k2"]).float())
    return l1 - l2 + LAMBDA_INIT


def _neighbour_index(H, W, device):
    """(N, 9) flat index of the 3x3 neighbours (row-major over (dr, dc)), and (N, 9) validity."""
    r = torch.arange(H, device=device).view(H, 1, 1, 1)
    c = torch.arange(W, device=device).view(1, W, 1, 1)
    dr = torch.tensor([-1, 0, 1], device=device).view(1, 1, 3, 1)
    dc = torch.tensor([-1, 0, 1], device=device).view(1, 1, 1, 3)
    rr, cc = (r + dr).expand(H, W, 3, 3), (c + dc).expand(H, W, 3, 3)
    ok = (rr >= 0) & (rr < H) & (cc >= 0) & (cc < W)
    flat = rr.clamp(0, H - 1) * W + cc.clamp(0, W - 1)
    return flat.reshape(H * W, 9), ok.reshape(H * W, 9)


def local_diff_attention(q, k, v, lam, subln_w, H, W):
    """q (B,N,2h,hd) already scaled; k (B,N,2h,hd); v (B,N,h,2hd) -> (B,N,h*2hd)."""
    Bn, Ntok, h2, hd = q.shape
    h = h2 // 2
    nb, ok = _neighbour_index(H, W, q.device)
    kn = k[:, nb]  # (B,N,9,2h,hd)
    vn = v[:, nb]  # (B,N,9,h,2hd)
    logits = torch.einsum("bnjd,bnpjd->bnjp", q, kn)
    logits = logits.masked_fill(~ok[None, :, None, :], float("-inf"))
    a = logits.softmax(-1).view(Bn, Ntok, h, 2, 9)
    a = a[:, :, :, 0] - lam.to(a.dtype) * a[:, :, :, 1]  # (B,N,h,9)
    o = torch.einsum("bnmp,bnpmd->bnmd", a, vn)
    o = rmsnorm(o, subln_w, 1e-5) * (1 - LAMBDA_INIT)
    return o.reshape(Bn, Ntok, h * 2 * hd)


def pooled_diff_attention(q, kp, vp, lam, subln_w):
    """q (B,N,h,2,hd) already scaled once; kp (B,P,h,2,hd); vp (B,P,h,2hd) -> (B,N,h*2hd)."""
    Bn, Ntok, h, _, hd = q.shape
    s = hd ** -0.5  # flash_attn_func's own default softmax_scale (second scaling, F4)
    logits = torch.einsum("bnmjd,bpmjd->bmjnp", q, kp) * s
    a = logits.softmax(-1)  # (B,h,2,N,P)
    o = torch.einsum("bmjnp,bpmd->bnmjd", a, vp)  # (B,N,h,2,2hd)
    o = o[:, :, :, 0] - lam.to(o.dtype) * o[:, :, :, 1]
    o = rmsnorm(o, subln_w, 1e-5) * (1 - LAMBDA_INIT)
    return o.reshape(Bn, Ntok, h * 2 * hd)


def aggregated_attention_forward(p, x, H, W, num_heads, local, sr_ratio=None, prefix=""):
    """x (B,N,C) tokens-major; num_heads = h (the module's `num_heads`, i.e. block heads // 2)."""
    g = lambda k: p[prefix + k]
    Bn, Ntok, C = x.shape
    h = num_heads
    hd = C // h // 2
    q = F.linear(x, g("q.weight"), g("q.bias")) * hd ** -0.5
    kl, vl = F.linear(x, g("kv.weight"), g("kv.bias")).chunk(2, dim=-1)
    lam = diff_lambda(p, prefix)
    if local:
        o = local_diff_attention(q.view(Bn, Ntok, 2 * h, hd), kl.view(Bn, Ntok, 2 * h, hd),
                                 vl.view(Bn, Ntok, h, 2 * hd), lam, g("subln.weight"), H, W)
    else:
        ph, pw = H // sr_ratio, W // sr_ratio
        t = F.conv2d(x.transpose(1, 2).reshape(Bn, C, H, W), g("sr.weight"), g("sr.bias"))
        t = F.adaptive_avg_pool2d(F.gelu(t), (ph, pw)).flatten(2).transpose(1, 2)
        t = F.layer_norm(t, (C,), g("norm.weight"), g("norm.bias"), 1e-5)
        kp, vp = F.linear(t, g("kv.weight"), g("kv.bias")).chunk(2, dim=-1)
        o = pooled_diff_attention(q.view(Bn, Ntok, h, 2, hd), kp.view(Bn, ph * pw, h, 2, hd),
                                  vp.view(Bn, ph * pw, h, 2 * hd), lam, g("subln.weight"))
    return o + dwconv3x3_tokens(vl, g("lepe.weight"), g("lepe.bias"), H, W)


def mlp_forward(p, x, prefix=""):
    x = F.gelu(F.linear(x, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"]))
    return F.linear(x, p[prefix + "fc2.weight"], p[prefix + "fc2.bias"])


def mlla_block_forward(p, x, num_heads, sr_ratio, prefix=""):
    """As-shipped MLAgg block, NCHW -> NCHW, eval mode (DropPath identity)."""
    g = lambda k: p[prefix + k]
    Bn, C, H, W = x.shape
    t = x.flatten(2).transpose(1, 2)
    short = t
    t = F.layer_norm(t, (C,), g("norm1.weight"), g("norm1.bias"), 1e-5)
    gate = F.silu(F.linear(t, g("act_proj.weight"), g("act_proj.bias")))
    t = F.linear(t, g("in_proj.weight"), g("in_proj.bias"))
    t = F.silu(dwconv3x3_tokens(t, g("dwc.weight"), g("dwc.bias"), H, W))
    a, b = t.chunk(2, dim=-1)
    a = aggregated_attention_forward(p, a, H, W, num_heads // 2, True, prefix=prefix + "attn.0.")
    b = aggregated_attention_forward(p, b, H, W, num_heads // 2, False, sr_ratio, prefix=prefix + "attn.1.")
    t = F.linear(torch.cat([a, b], -1) * gate, g("out_proj.weight"), g("out_proj.bias"))
    t = short + t
    t = t + mlp_forward(p, F.layer_norm(t, (C,), g("norm2.weight"), g("norm2.bias"), 1e-5), prefix + "mlp.")
    return t.transpose(1, 2).reshape(Bn, C, H, W)
