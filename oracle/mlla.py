"""oracle/mlla.py -- TEST INFRASTRUCTURE ONLY (parity oracle / CPU baseline).

Functional CPU restatement of the elu+1 linear attention that BASELINE.json:north_star names by op
and that lives in the sibling trainer of the reference:
  * `RoPE`                 nnUNetTrainer_MLLA_UNet.py:169-195
  * `LinearAttention`      nnUNetTrainer_MLLA_UNet.py:198-250
  * `MLLABlock.forward`    nnUNetTrainer_MLLA_UNet.py:293-315
Math per SURVEY.md App. A.5, written with real arithmetic (cos/sin pairs) instead of
view_as_complex.  Parameters: flat dict with the reference's state_dict names.

Pinned against the reference module source executed in this container
(tests/golden/make_golden.py -> tests/golden/mlla_*.pt).
Nothing under mlagg-unet_b200/ imports this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .mlagg import mlp_forward
from .msmm import dwconv3x3_tokens


def rope_angles(H, W, C, base=10000.0, device=None):
    """(H, W, C/2) rotation angle of each complex channel pair: first C/4 pairs use the row index,
    the next C/4 the column index, theta_i = base**(-i/(C/4))."""
    k = C // 4
    theta = 1.0 / (base ** (torch.arange(k, device=device, dtype=torch.float32) / k))
    r = torch.arange(H, device=device, dtype=torch.float32).view(H, 1, 1) * theta
    c = torch.arange(W, device=device, dtype=torch.float32).view(1, W, 1) * theta
    return torch.cat([r.expand(H, W, k), c.expand(H, W, k)], dim=-1)


def rope(x, H, W):
    """x (B, H*W, C) -> rotated, fp32 (the reference forces fp32)."""
    Bn, Ntok, C = x.shape
    ang = rope_angles(H, W, C, device=x.device).reshape(1, Ntok, C // 2)
    x = x.to(torch.float32 if x.dtype != torch.float64 else x.dtype).reshape(Bn, Ntok, C // 2, 2)
    cs, sn = torch.cos(ang).to(x.dtype), torch.sin(ang).to(x.dtype)
    re = cs * x[..., 0] - sn * x[..., 1]
    im = sn * x[..., 0] + cs * x[..., 1]
    return torch.stack([re, im], dim=-1).reshape(Bn, Ntok, C)


def linear_attention_core(q, k, v, H, W, num_heads):
    """q, k raw projections, v: (B,N,C) -> (B,N,C).  nnUNetTrainer_MLLA_UNet.py:234-246: elu+1, RoPE on q and k,
    z = 1/(q . mean_n k + 1e-6), kv = (k_rope^T n^-1/2)(v n^-1/2), out = q_rope kv z.  Runs in the input dtype
    (fp64 inputs give the arbiter used by the GPU parity tests)."""
    Bn, Ntok, C = q.shape
    hd = C // num_heads
    q, k = F.elu(q) + 1.0, F.elu(k) + 1.0
    heads = lambda t: t.reshape(Bn, Ntok, num_heads, hd).transpose(1, 2)  # (B,h,N,hd)
    qr, kr = heads(rope(q, H, W)), heads(rope(k, H, W))
    qh, kh, vh = heads(q), heads(k), heads(v)
    z = 1.0 / (torch.einsum("bhnd,bhd->bhn", qh, kh.mean(dim=2)) + 1e-6)
    state = torch.einsum("bhnd,bhne->bhde", kr * Ntok ** -0.5, vh * Ntok ** -0.5)
    o = torch.einsum("bhnd,bhde->bhne", qr, state) * z[..., None]
    return o.transpose(1, 2).reshape(Bn, Ntok, C)


def linear_attention_forward(p, x, H, W, num_heads, prefix=""):
    """x (B,N,C) -> (B,N,C): q,k = W_qk x; v = x; elu+1; RoPE; per-head state; normaliser; + LePE(v)."""
    g = lambda k: p[prefix + k]
    q, k = F.linear(x, g("qk.weight"), g("qk.bias")).chunk(2, dim=-1)
    o = linear_attention_core(q, k, x, H, W, num_heads)
    return o + dwconv3x3_tokens(x, g("lepe.weight"), g("lepe.bias"), H, W)


def mlla_block_v1_forward(p, x, H, W, num_heads, prefix=""):
    """MLLA-UNet block, tokens-major (B,L,C) -> (B,L,C), eval mode."""
    g = lambda k: p[prefix + k]
    C = x.shape[-1]
    x = x + dwconv3x3_tokens(x, g("cpe1.weight"), g("cpe1.bias"), H, W)
    short = x
    t = F.layer_norm(x, (C,), g("norm1.weight"), g("norm1.bias"), 1e-5)
    gate = F.silu(F.linear(t, g("act_proj.weight"), g("act_proj.bias")))
    t = F.linear(t, g("in_proj.weight"), g("in_proj.bias"))
    t = F.silu(dwconv3x3_tokens(t, g("dwc.weight"), g("dwc.bias"), H, W))
    t = linear_attention_forward(p, t, H, W, num_heads, prefix + "attn.")
    x = short + F.linear(t.to(gate.dtype) * gate, g("out_proj.weight"), g("out_proj.bias"))
    x = x + dwconv3x3_tokens(x, g("cpe2.weight"), g("cpe2.bias"), H, W)
    return x + mlp_forward(p, F.layer_norm(x, (C,), g("norm2.weight"), g("norm2.bias"), 1e-5), prefix + "mlp.")
