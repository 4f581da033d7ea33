"""Attention cores of the MLAgg block and of MLLA linear attention, tokens-major.

  local_diff_attention   3x3-window differential softmax attention + sub-LN  (reference
                         nnUNetTrainer_MLAgg_2D_dt_MS.py:693-717)
  pooled_diff_attention  differential softmax attention over P pooled tokens + sub-LN (:732-760), i.e. the four
                         flash_attn_func calls + cat + lambda-combine + RMSNorm of the reference in one op,
                         INCLUDING flash-attn's second head_dim**-0.5 scale
  linear_attention_core  elu+1 / RoPE / per-head K^T V state / normaliser (nnUNetTrainer_MLLA_UNet.py:234-246)

All three are sm_100a kernels behind the C ABI (csrc/local_attn.cu, csrc/pooled_attn.cu, csrc/linattn.cu); the
torch ops left here are the RoPE module's standalone forward (API compatibility) and the scalar lambda.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib

LAMBDA_INIT = 0.8
_DT = {torch.float32: 0, torch.bfloat16: 1}


def rmsnorm_affine(x, weight, eps):
    y = x.float()
    y = y * torch.rsqrt(y.pow(2).mean(-1, keepdim=True) + eps)
    return y.type_as(x) * weight


class _DiffLambda(torch.autograd.Function):
    """C ABI: mlagg_diff_lambda_fwd / _bwd (csrc/ew.cu): one one-warp kernel each way instead of 8 + 14."""

    @staticmethod
    def forward(ctx, lq1, lk1, lq2, lk2):
        n = lq1.numel()
        out = torch.empty(3, device=lq1.device, dtype=torch.float32)
        with torch.cuda.device(lq1.device), _lib.timed("diff_lambda"):
            rc = _lib.lib().mlagg_diff_lambda_fwd(lq1.data_ptr(), lk1.data_ptr(), lq2.data_ptr(), lk2.data_ptr(), n,
                                                  LAMBDA_INIT, out.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "mlagg_diff_lambda_fwd")
        ctx.save_for_backward(lq1, lk1, lq2, lk2, out)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        lq1, lk1, lq2, lk2, out = ctx.saved_tensors
        n = lq1.numel()
        g = g.detach().float().reshape(1).contiguous()
        grads = torch.empty(4, n, device=lq1.device, dtype=torch.float32)
        with torch.cuda.device(lq1.device), _lib.timed("diff_lambda"):
            rc = _lib.lib().mlagg_diff_lambda_bwd(lq1.data_ptr(), lk1.data_ptr(), lq2.data_ptr(), lk2.data_ptr(),
                                                  out.data_ptr(), g.data_ptr(), n, grads.data_ptr(), _lib.stream_ptr())
        _lib.check(rc, "mlagg_diff_lambda_bwd")
        return tuple(grads[i].view_as(lq1) for i in range(4))


def diff_lambda(lq1, lk1, lq2, lk2):
    """exp(<lq1, lk1>) - exp(<lq2, lk2>) + lambda_init (reference :700-702, :745-747) as a 0-dim fp32 tensor"""
    ts = (lq1, lk1, lq2, lk2)
    if all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() for t in ts):
        return _DiffLambda.apply(*ts)
    return torch.exp(torch.sum(lq1 * lk1).float()) - torch.exp(torch.sum(lq2 * lk2).float()) + LAMBDA_INIT


def _rows(t):
    """tokens-major (B, N, C) view usable by the C ABI: unit channel stride, one uniform row stride over B*N."""
    if t.stride(2) != 1 or t.stride(0) != t.shape[1] * t.stride(1) or (t.stride(1) * t.element_size()) % 16 \
            or t.data_ptr() % 16:
        t = t.contiguous()
    return t


class _LocalDiffAttn(torch.autograd.Function):
    """C ABI: mlagg_local_diffattn_fwd / _bwd (csrc/local_attn.cu)."""

    @staticmethod
    def forward(ctx, q, kv, lam, subln_w, H, W, h, hd, scale):
        if not q.is_cuda:
            raise _lib.MlaggError("local_diff_attention: CUDA tensors required (no CPU fallback in the product path)")
        Bn, N, C = q.shape
        dt = q.dtype if q.dtype in _DT else torch.float32
        q_, kv_ = q.to(dt).contiguous(), kv.to(dt).contiguous()
        lam_ = lam.detach().float().reshape(1).contiguous()
        w_ = subln_w.detach().float().contiguous()
        out = torch.empty_like(q_)
        es = q_.element_size()
        with torch.cuda.device(q.device), _lib.timed("local_diffattn_fwd"):
            rc = _lib.lib().mlagg_local_diffattn_fwd(q_.data_ptr(), kv_.data_ptr(), kv_.data_ptr() + C * es, w_.data_ptr(),
                                                     out.data_ptr(), Bn, H, W, h, hd, C, 2 * C, C, scale,
                                                     lam_.data_ptr(), 1e-5, 1.0 - LAMBDA_INIT, _DT[dt], _lib.stream_ptr())
        _lib.check(rc, "mlagg_local_diffattn_fwd")
        ctx.save_for_backward(q_, kv_, lam_, w_)
        ctx.meta = (H, W, h, hd, scale, q.dtype, kv.dtype, lam.dtype, subln_w.dtype)
        return out.to(q.dtype)

    @staticmethod
    def backward(ctx, dout):
        q_, kv_, lam_, w_ = ctx.saved_tensors
        H, W, h, hd, scale, qdt, kvdt, lamdt, wdt = ctx.meta
        Bn, N, C = q_.shape
        dt, es = q_.dtype, q_.element_size()
        dout = _rows(dout.to(dt))          # a half of the concatenated block gradient is read in place (lddo)
        dq, dkv = torch.empty_like(q_), torch.empty_like(kv_)
        dw = _lib.zeros(w_.shape, q_.device)
        dlam = _lib.zeros(1, q_.device)
        L = _lib.lib()
        ws = torch.empty(L.mlagg_local_diffattn_ws_bytes(Bn, H, W, h, hd) // 4, device=q_.device, dtype=torch.float32)
        with torch.cuda.device(q_.device), _lib.timed("local_diffattn_bwd", 2):
            rc = L.mlagg_local_diffattn_bwd(q_.data_ptr(), kv_.data_ptr(), kv_.data_ptr() + C * es, w_.data_ptr(),
                                            dout.data_ptr(), dq.data_ptr(), dkv.data_ptr(), dkv.data_ptr() + C * es,
                                            dw.data_ptr(), dlam.data_ptr(), ws.data_ptr(), Bn, H, W, h, hd, C, 2 * C,
                                            dout.stride(1), C, 2 * C, scale, lam_.data_ptr(), 1e-5, 1.0 - LAMBDA_INIT, _DT[dt],
                                            _lib.stream_ptr())
        _lib.check(rc, "mlagg_local_diffattn_bwd")
        return dq.to(qdt), dkv.to(kvdt), dlam.reshape(()).to(lamdt), dw.to(wdt), None, None, None, None, None


def local_diff_attention(q, kv, lam, subln_w, H, W, h, hd, scale):
    """q (B,N,C) RAW projection (scale applied in-kernel); kv (B,N,2C) = [k | v]; C = 2*h*hd -> (B,N,C)."""
    return _LocalDiffAttn.apply(q, kv, lam, subln_w, H, W, h, hd, scale)


class _PooledDiffAttn(torch.autograd.Function):
    """C ABI: mlagg_pooled_diffattn_fwd / _bwd (csrc/pooled_attn.cu)."""

    @staticmethod
    def forward(ctx, q, kvp, lam, subln_w, h, hd, scale):
        if not q.is_cuda:
            raise _lib.MlaggError("pooled_diff_attention: CUDA tensors required (no CPU fallback in the product path)")
        Bn, N, C = q.shape
        P = kvp.shape[1]
        dt = q.dtype if q.dtype in _DT else torch.float32
        q_, kv_ = q.to(dt).contiguous(), kvp.to(dt).contiguous()
        lam_ = lam.detach().float().reshape(1).contiguous()
        w_ = subln_w.detach().float().contiguous()
        out = torch.empty_like(q_)
        lse = torch.empty(_lib.lib().mlagg_pooled_diffattn_saved_bytes(Bn, N, h, hd) // 4, device=q.device,
                          dtype=torch.float32)      # log-sum-exps + the per-map outputs O0 | O1 (for the backward pass)
        es = q_.element_size()
        with torch.cuda.device(q.device), _lib.timed("pooled_diffattn_fwd"):
            rc = _lib.lib().mlagg_pooled_diffattn_fwd(q_.data_ptr(), kv_.data_ptr(), kv_.data_ptr() + C * es,
                                                      w_.data_ptr(), out.data_ptr(), lse.data_ptr(), Bn, N, P, h, hd, C,
                                                      2 * C, C, scale, lam_.data_ptr(), 1e-5, 1.0 - LAMBDA_INIT,
                                                      _DT[dt], _lib.stream_ptr())
        _lib.check(rc, "mlagg_pooled_diffattn_fwd")
        ctx.save_for_backward(q_, kv_, lam_, w_, lse)
        ctx.meta = (h, hd, scale, q.dtype, kvp.dtype, lam.dtype, subln_w.dtype)
        return out.to(q.dtype)

    @staticmethod
    def backward(ctx, dout):
        q_, kv_, lam_, w_, lse = ctx.saved_tensors
        h, hd, scale, qdt, kvdt, lamdt, wdt = ctx.meta
        Bn, N, C = q_.shape
        P = kv_.shape[1]
        dt, es = q_.dtype, q_.element_size()
        dout = _rows(dout.to(dt))
        dq = torch.empty_like(q_)
        dkv = torch.zeros(Bn, P, 2 * C, device=q_.device, dtype=torch.float32)
        dw = _lib.zeros(w_.shape, q_.device)
        dlam = _lib.zeros(1, q_.device)
        L = _lib.lib()
        ws = torch.empty(L.mlagg_pooled_diffattn_ws_bytes(Bn, N, h, hd) // 4, device=q_.device, dtype=torch.float32)
        with torch.cuda.device(q_.device), _lib.timed("pooled_diffattn_bwd", 2):
            rc = L.mlagg_pooled_diffattn_bwd(q_.data_ptr(), kv_.data_ptr(), kv_.data_ptr() + C * es, w_.data_ptr(),
                                             lse.data_ptr(), dout.data_ptr(), dq.data_ptr(), dkv.data_ptr(),
                                             dkv.data_ptr() + C * 4, dw.data_ptr(), dlam.data_ptr(), ws.data_ptr(), Bn, N,
                                             P, h, hd, C, 2 * C, dout.stride(1), C, 2 * C, scale, lam_.data_ptr(), 1e-5,
                                             1.0 - LAMBDA_INIT, _DT[dt], _lib.stream_ptr())
        _lib.check(rc, "mlagg_pooled_diffattn_bwd")
        return dq.to(qdt), dkv.to(kvdt), dlam.reshape(()).to(lamdt), dw.to(wdt), None, None, None


def pooled_diff_attention(q, kvp, lam, subln_w, h, hd, scale):
    """q (B,N,C) RAW projection; kvp (B,P,2C) = [k_pool | v_pool]; C = 2*h*hd -> (B,N,C).  Both hd**-0.5 scalings
    of the shipped reference are applied in-kernel."""
    return _PooledDiffAttn.apply(q, kvp, lam, subln_w, h, hd, scale)


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


_ROPE_SEP = {}


def rope_table_separable(H, W, C, device, base=10000.0):
    """(H + W, C/4, 2) fp32 [cos, sin]: rows 0..H-1 hold row*theta_i, rows H..H+W-1 hold col*theta_i -- the reference's
    `rotations` buffer (MLLA_UNet.py:181-187) is exactly the outer arrangement of these two; built on the CPU with the
    same expressions so the values are bit-identical to that buffer."""
    key = (H, W, C, str(device), float(base))
    if key not in _ROPE_SEP:
        k = C // 4
        theta = 1 / (base ** (torch.arange(k) / k))
        ang = torch.cat([torch.arange(H).unsqueeze(-1) * theta, torch.arange(W).unsqueeze(-1) * theta], dim=0)
        _ROPE_SEP[key] = torch.stack([torch.cos(ang), torch.sin(ang)], dim=-1).float().contiguous().to(device)
    return _ROPE_SEP[key]


class _LinAttn(torch.autograd.Function):
    """C ABI: mlagg_linattn_fwd / _bwd (csrc/linattn.cu)."""

    @staticmethod
    def forward(ctx, qk, v, H, W, num_heads):
        if not qk.is_cuda:
            raise _lib.MlaggError("linear_attention: CUDA tensors required (no CPU fallback in the product path)")
        Bn, N, C2 = qk.shape
        C = C2 // 2
        hd = C // num_heads
        assert N == H * W and v.shape == (Bn, N, C) and C % num_heads == 0 and C % 4 == 0
        dt = qk.dtype if qk.dtype in _DT else torch.float32
        qk_, v_ = _rows(qk.detach().to(dt)), _rows(v.detach().to(dt))
        rope = rope_table_separable(H, W, C, qk.device)
        L = _lib.lib()
        out = torch.empty(Bn, N, C, device=qk.device, dtype=dt)
        state = torch.empty(L.mlagg_linattn_state_bytes(Bn, num_heads, hd) // 4, device=qk.device, dtype=torch.float32)
        es = qk_.element_size()
        with torch.cuda.device(qk.device), _lib.timed("linattn_fwd", 2):
            rc = L.mlagg_linattn_fwd(qk_.data_ptr(), qk_.data_ptr() + C * es, v_.data_ptr(), rope.data_ptr(),
                                     out.data_ptr(), state.data_ptr(), Bn, H, W, num_heads, hd, qk_.stride(1),
                                     qk_.stride(1), v_.stride(1), C, 1e-6, _DT[dt], _lib.stream_ptr())
        _lib.check(rc, "mlagg_linattn_fwd")
        ctx.save_for_backward(qk_, v_, rope, state)
        ctx.meta = (H, W, num_heads, hd, qk.dtype, v.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        qk_, v_, rope, state = ctx.saved_tensors
        H, W, h, hd, qkdt, vdt = ctx.meta
        Bn, N, C = v_.shape
        dt, es = qk_.dtype, qk_.element_size()
        dout = _rows(dout.to(dt))
        dqk = torch.empty(Bn, N, 2 * C, device=qk_.device, dtype=dt)
        dv = torch.empty(Bn, N, C, device=qk_.device, dtype=dt)
        L = _lib.lib()
        ws = torch.empty(state.numel(), device=qk_.device, dtype=torch.float32)
        with torch.cuda.device(qk_.device), _lib.timed("linattn_bwd", 2):
            rc = L.mlagg_linattn_bwd(qk_.data_ptr(), qk_.data_ptr() + C * es, v_.data_ptr(), rope.data_ptr(),
                                     state.data_ptr(), dout.data_ptr(), dqk.data_ptr(), dqk.data_ptr() + C * es,
                                     dv.data_ptr(), ws.data_ptr(), Bn, H, W, h, hd, qk_.stride(1), qk_.stride(1),
                                     v_.stride(1), dout.stride(1), 2 * C, 2 * C, C, 1e-6, _DT[dt], _lib.stream_ptr())
        _lib.check(rc, "mlagg_linattn_bwd")
        return dqk.to(qkdt), dv.to(vdt), None, None, None


def linear_attention_qk(qk, v, H, W, num_heads):
    """qk (B,N,2C) = the raw [q | k] projection, consumed in place; v (B,N,C) -> (B,N,C) in qk's dtype (fp32 math):
    phi = elu+1, RoPE, per-head state (1/N) sum rope(phi k) (x) v, normaliser 1/(phi q . mean phi k + 1e-6)."""
    return _LinAttn.apply(qk, v, H, W, num_heads)


def linear_attention_core(q, k, v, H, W, num_heads):
    """Same op with q and k given separately (B,N,C each)."""
    return _LinAttn.apply(torch.cat([q, k], dim=-1), v, H, W, num_heads)
