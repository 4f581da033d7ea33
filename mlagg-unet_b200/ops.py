"""torch.autograd.Function wrappers over the C ABI (include/mlagg_b200.h) for the stencil ops of the hot path.

All functions take TOKENS-MAJOR activations (B, N, C) with N = H*W row-major -- the layout the reference's
blocks already use between their Linear layers -- so none of the reference's `permute(...).contiguous()`
round trips around nn.Conv2d are needed.  CUDA only; no CPU path.
"""
from __future__ import annotations

import os

import torch

from . import _lib, gemm

_DT = {torch.float32: 0, torch.bfloat16: 1}


def _io(x):
    """kernels take fp32 or bf16 activations; fp16 (reference autocast dtype) is widened to fp32"""
    return x if x.dtype in _DT else x.float()


def _tok_strides(t, C):
    """(pixel stride, image stride) of a tokens-major (B, N, C) view the stencil kernels can address in place, or None:
    unit channel stride, strides that keep the 4-channel vectors aligned."""
    if t.dim() != 3 or t.stride(2) != 1 or t.stride(1) < C:
        return None
    ld, bs = t.stride(1), t.stride(0)
    if C % 4 == 0 and (ld % 4 or bs % 4 or (t.data_ptr() % (4 * t.element_size()))):
        return None
    return ld, bs


def _tok_operand(t, C):
    """t as the kernels take it (fp32 / bf16) plus its (ld, bs); copies only when the view cannot be addressed in place"""
    t = _io(t)
    st = _tok_strides(t, C)
    if st is None:
        t = t.contiguous()
        st = (t.stride(1), t.stride(0))
    return t, st


class _DWConv3x3(torch.autograd.Function):
    """C ABI: mlagg_dwconv3x3_fwd_strided / _bwd_strided.  y = act(conv(x) + b) [+ residual]"""

    @staticmethod
    def forward(ctx, x, weight, bias, H, W, silu, residual):
        if not x.is_cuda:
            raise _lib.MlaggError("dwconv3x3_tokens: CUDA tensor required (no CPU fallback in the product path)")
        Bn, N, C = x.shape
        assert N == H * W and weight.shape == (C, 1, 3, 3)
        xin, (ldx, bsx) = _tok_operand(x, C)
        w32 = weight.detach().float().contiguous()
        b32 = None if bias is None else bias.detach().float().contiguous()
        res, (ldr, bsr) = (None, (C, N * C))
        if residual is not None:
            assert residual.shape == x.shape
            res, (ldr, bsr) = _tok_operand(residual.to(xin.dtype), C)
        y = torch.empty(Bn, N, C, device=x.device, dtype=xin.dtype)
        with torch.cuda.device(x.device), _lib.timed("dwconv3x3_fwd", 1, (2 + (res is not None)) * y.numel() * y.element_size()):
            rc = _lib.lib().mlagg_dwconv3x3_fwd_strided(_lib.ptr(xin), _lib.ptr(w32), _lib.ptr(b32), _lib.ptr(res),
                                                        _lib.ptr(y), Bn, H, W, C, ldx, bsx, ldr, bsr, C, N * C, int(silu),
                                                        0, _DT[xin.dtype], _lib.stream_ptr())
        _lib.check(rc, "mlagg_dwconv3x3_fwd_strided")
        ctx.save_for_backward(xin, w32, b32)
        ctx.meta = (H, W, bool(silu), x.dtype, weight.dtype, None if bias is None else bias.dtype, (ldx, bsx),
                    None if residual is None else residual.dtype)
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        xin, w32, b32 = ctx.saved_tensors
        H, W, silu, xdt, wdt, bdt, (ldx, bsx), rdt = ctx.meta
        Bn, N, C = xin.shape
        dy, (ldg, bsg) = _tok_operand(dy.to(xin.dtype), C)
        dz = torch.empty(Bn, N, C, device=xin.device, dtype=xin.dtype)
        dx = torch.empty_like(dz)
        dw = _lib.zeros((C, 9), xin.device)
        db = _lib.zeros(C, xin.device) if b32 is not None else None
        with torch.cuda.device(xin.device), _lib.timed("dwconv3x3_bwd", 2, 6 * dz.numel() * dz.element_size()):
            rc = _lib.lib().mlagg_dwconv3x3_bwd_strided(_lib.ptr(xin), _lib.ptr(w32), _lib.ptr(b32), _lib.ptr(dy),
                                                        _lib.ptr(dz), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), Bn, H, W,
                                                        C, ldx, bsx, ldg, bsg, C, N * C, None, None, 0, 0, int(silu),
                                                        _DT[xin.dtype], _lib.stream_ptr())
        _lib.check(rc, "mlagg_dwconv3x3_bwd_strided")
        return (dx.to(xdt), dw.view(C, 1, 3, 3).to(wdt), None if db is None else db.to(bdt), None, None, None,
                None if rdt is None else dy.to(rdt))


def dwconv3x3_tokens(x, weight, bias, H, W, silu=False, residual=None):
    """x (B, H*W, C) -> depthwise 3x3 (pad 1) [+ SiLU] [+ residual]; weight is the nn.Conv2d(C, C, 3, groups=C)
    parameter.  x / residual may be channel slices of wider activations (read in place)."""
    return _DWConv3x3.apply(x, weight, bias, H, W, silu, residual)


class _ConvGLUCore(torch.autograd.Function):
    """ConvolutionalGLU between fc1 and fc2 (reference MambaSkip.py:572-574): h = fc1(x) (B, N, 2 hid) ->
    act(dwconv(h[..., :hid])) * h[..., hid:] in ONE kernel on the two halves of h in place; the backward writes the
    gradients of both halves into one (B, N, 2 hid) tensor (C ABI: mlagg_dwconv3x3_fwd_strided with residual_mul = 1,
    mlagg_dwconv3x3_bwd_strided with mul / dmul)."""

    @staticmethod
    def forward(ctx, h, weight, bias, H, W, silu):
        if not h.is_cuda:
            raise _lib.MlaggError("conv_glu_core: CUDA tensor required (no CPU fallback in the product path)")
        Bn, N, C2 = h.shape
        C = C2 // 2
        assert N == H * W and C2 == 2 * C and weight.shape == (C, 1, 3, 3)
        hin = _io(h).contiguous()
        w32 = weight.detach().float().contiguous()
        b32 = None if bias is None else bias.detach().float().contiguous()
        y = torch.empty(Bn, N, C, device=h.device, dtype=hin.dtype)
        es = hin.element_size()
        with torch.cuda.device(h.device), _lib.timed("dwconv3x3_fwd", 1, 3 * y.numel() * es):
            rc = _lib.lib().mlagg_dwconv3x3_fwd_strided(hin.data_ptr(), _lib.ptr(w32), _lib.ptr(b32), hin.data_ptr() + C * es,
                                                        _lib.ptr(y), Bn, H, W, C, C2, N * C2, C2, N * C2, C, N * C,
                                                        int(silu), 1, _DT[hin.dtype], _lib.stream_ptr())
        _lib.check(rc, "mlagg_dwconv3x3_fwd_strided")
        ctx.save_for_backward(hin, w32, b32)
        ctx.meta = (H, W, bool(silu), h.dtype, weight.dtype, None if bias is None else bias.dtype)
        return y.to(h.dtype)

    @staticmethod
    def backward(ctx, dy):
        hin, w32, b32 = ctx.saved_tensors
        H, W, silu, hdt, wdt, bdt = ctx.meta
        Bn, N, C2 = hin.shape
        C = C2 // 2
        dy, (ldg, bsg) = _tok_operand(dy.to(hin.dtype), C)
        dz = torch.empty(Bn, N, C, device=hin.device, dtype=hin.dtype)
        dh = torch.empty_like(hin)
        dw = _lib.zeros((C, 9), hin.device)
        db = _lib.zeros(C, hin.device) if b32 is not None else None
        es = hin.element_size()
        with torch.cuda.device(hin.device), _lib.timed("dwconv3x3_bwd", 2, 8 * dz.numel() * es):
            rc = _lib.lib().mlagg_dwconv3x3_bwd_strided(hin.data_ptr(), _lib.ptr(w32), _lib.ptr(b32), _lib.ptr(dy),
                                                        _lib.ptr(dz), dh.data_ptr(), _lib.ptr(dw), _lib.ptr(db), Bn, H, W,
                                                        C, C2, N * C2, ldg, bsg, C2, N * C2, hin.data_ptr() + C * es,
                                                        dh.data_ptr() + C * es, C2, N * C2, int(silu), _DT[hin.dtype],
                                                        _lib.stream_ptr())
        _lib.check(rc, "mlagg_dwconv3x3_bwd_strided")
        return dh.to(hdt), dw.view(C, 1, 3, 3).to(wdt), None if db is None else db.to(bdt), None, None, None


def conv_glu_core(h, weight, bias, H, W, silu=True):
    """h (B, H*W, 2 hid) = fc1 output -> act(dwconv3x3(h[..., :hid])) * h[..., hid:]  (B, H*W, hid)"""
    return _ConvGLUCore.apply(h, weight, bias, H, W, silu)


class _DWConv3x3Stages(torch.autograd.Function):
    """Per-stage depthwise 3x3 + SiLU on the stage-concatenated sequence x (B, L, C), L = sum H_s W_s, each stage with its
    own Conv2d parameters (reference MambaSkip.py:521-523: split, permute, conv2d[i], act[i], flatten, cat): the stage
    segments are read and written in place through the image stride L*C -- no split / contiguous / cat copies, and the
    backward writes one dx instead of autograd's zero-fill + copy + add per stage."""

    @staticmethod
    def forward(ctx, x, hw, silu, *params):
        if not x.is_cuda:
            raise _lib.MlaggError("dwconv3x3_stages: CUDA tensor required (no CPU fallback in the product path)")
        Bn, L, C = x.shape
        ns = len(hw)
        ws, bs_ = params[:ns], params[ns:]
        xin = _io(x).contiguous()
        y = torch.empty_like(xin)
        w32 = [w.detach().float().contiguous() for w in ws]
        b32 = [None if b is None else b.detach().float().contiguous() for b in bs_]
        es, off = xin.element_size(), 0
        with torch.cuda.device(x.device), _lib.timed("dwconv3x3_fwd", ns, 2 * y.numel() * es):
            for s, (h, w) in enumerate(hw):
                rc = _lib.lib().mlagg_dwconv3x3_fwd_strided(xin.data_ptr() + off * C * es, _lib.ptr(w32[s]),
                                                            _lib.ptr(b32[s]), None, y.data_ptr() + off * C * es, Bn, h, w,
                                                            C, C, L * C, C, L * C, C, L * C, int(silu), 0, _DT[xin.dtype],
                                                            _lib.stream_ptr())
                _lib.check(rc, "mlagg_dwconv3x3_fwd_strided")
                off += h * w
        assert off == L
        ctx.save_for_backward(xin, *w32, *[b for b in b32 if b is not None])
        ctx.meta = (tuple(hw), bool(silu), x.dtype, [w.dtype for w in ws], [None if b is None else b.dtype for b in bs_])
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        hw, silu, xdt, wdts, bdts = ctx.meta
        ns = len(hw)
        saved = ctx.saved_tensors
        xin, w32 = saved[0], saved[1:1 + ns]
        it = iter(saved[1 + ns:])
        b32 = [None if d is None else next(it) for d in bdts]
        Bn, L, C = xin.shape
        dy = dy.to(xin.dtype).contiguous()
        dx = torch.empty_like(xin)
        es, off = xin.element_size(), 0
        dws, dbs = [], []
        with torch.cuda.device(xin.device), _lib.timed("dwconv3x3_bwd", 2 * ns, 6 * dx.numel() * es):
            for s, (h, w) in enumerate(hw):
                dz = torch.empty(Bn, h * w, C, device=xin.device, dtype=xin.dtype)
                dw = _lib.zeros((C, 9), xin.device)
                db = _lib.zeros(C, xin.device) if b32[s] is not None else None
                o = off * C * es
                rc = _lib.lib().mlagg_dwconv3x3_bwd_strided(xin.data_ptr() + o, _lib.ptr(w32[s]), _lib.ptr(b32[s]),
                                                            dy.data_ptr() + o, _lib.ptr(dz), dx.data_ptr() + o,
                                                            _lib.ptr(dw), _lib.ptr(db), Bn, h, w, C, C, L * C, C, L * C, C,
                                                            L * C, None, None, 0, 0, int(silu), _DT[xin.dtype],
                                                            _lib.stream_ptr())
                _lib.check(rc, "mlagg_dwconv3x3_bwd_strided")
                dws.append(dw.view(C, 1, 3, 3).to(wdts[s]))
                dbs.append(None if db is None else db.to(bdts[s]))
                off += h * w
        return (dx.to(xdt), None, None, *dws, *dbs)


def dwconv3x3_stages(x, hw, convs, silu=True):
    """x (B, sum H_s W_s, C); convs[s] = nn.Conv2d(C, C, 3, padding=1, groups=C) of stage s."""
    hw = tuple((int(h), int(w)) for h, w in hw)
    return _DWConv3x3Stages.apply(x, hw, silu, *[c.weight for c in convs], *[c.bias for c in convs])


class _CausalConv1d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, silu):
        if not x.is_cuda:
            raise _lib.MlaggError("causal_conv1d_fn: CUDA tensor required (no CPU fallback in the product path)")
        Bn, C, L = x.shape
        K = weight.shape[1]
        x32, w32 = x.float().contiguous(), weight.detach().float().contiguous()
        b32 = None if bias is None else bias.detach().float().contiguous()
        y = torch.empty_like(x32)
        with torch.cuda.device(x.device), _lib.timed("causal_conv1d_fwd"):
            rc = _lib.lib().mlagg_causal_conv1d_fwd(_lib.ptr(x32), _lib.ptr(w32), _lib.ptr(b32), _lib.ptr(y), Bn, C, L,
                                                    K, int(silu), _lib.stream_ptr())
        _lib.check(rc, "mlagg_causal_conv1d_fwd")
        ctx.save_for_backward(x32, w32, b32)
        ctx.meta = (bool(silu), x.dtype, weight.dtype, None if bias is None else bias.dtype)
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        x32, w32, b32 = ctx.saved_tensors
        silu, xdt, wdt, bdt = ctx.meta
        Bn, C, L = x32.shape
        K = w32.shape[1]
        dy = dy.float().contiguous()
        dx = torch.empty_like(x32)
        dw = torch.zeros_like(w32)
        db = torch.zeros(C, device=x32.device, dtype=torch.float32) if b32 is not None else None
        with torch.cuda.device(x32.device), _lib.timed("causal_conv1d_bwd"):
            rc = _lib.lib().mlagg_causal_conv1d_bwd(_lib.ptr(x32), _lib.ptr(w32), _lib.ptr(b32), _lib.ptr(dy),
                                                    _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), Bn, C, L, K, int(silu),
                                                    _lib.stream_ptr())
        _lib.check(rc, "mlagg_causal_conv1d_bwd")
        return dx.to(xdt), dw.to(wdt), None if db is None else db.to(bdt), None


def causal_conv1d_fn(x, weight, bias=None, activation=None):
    """Same call shape as causal_conv1d.causal_conv1d_fn: x (B, C, L), weight (C, K<=4), activation in
    {None, 'silu', 'swish'}."""
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu, or swish")
    return _CausalConv1d.apply(x, weight, bias, activation is not None)


def _ln_forward(ctx, x, weight, bias, eps, out_dtype):
    if not x.is_cuda:
        raise _lib.MlaggError("layer_norm_tokens: CUDA tensor required (no CPU fallback in the product path)")
    C = x.shape[-1]
    xin = _io(x).contiguous()
    odt = out_dtype if out_dtype in _DT else torch.float32
    w32 = weight.detach().float().contiguous()
    b32 = None if bias is None else bias.detach().float().contiguous()
    M = xin.numel() // C
    y = torch.empty(xin.shape, device=x.device, dtype=odt)
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty(M, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device), _lib.timed("layernorm_fwd", 1, xin.numel() * xin.element_size() + y.numel() * y.element_size() + 8 * M):
        rc = _lib.lib().mlagg_layernorm_fwd(_lib.ptr(xin), _lib.ptr(w32), _lib.ptr(b32), _lib.ptr(y), _lib.ptr(mean),
                                            _lib.ptr(rstd), M, C, float(eps), _DT[xin.dtype], _DT[odt],
                                            _lib.stream_ptr())
    _lib.check(rc, "mlagg_layernorm_fwd")
    ctx.save_for_backward(xin, w32, mean, rstd)
    ctx.meta = (x.dtype, weight.dtype, None if bias is None else bias.dtype, odt)
    return y


def _ln_backward(ctx, dy, dres):
    xin, w32, mean, rstd = ctx.saved_tensors
    xdt, wdt, bdt, odt = ctx.meta
    C = xin.shape[-1]
    M = xin.numel() // C
    dy = dy.to(odt).contiguous()
    if dres is not None:
        dres = dres.to(xin.dtype).contiguous()
    dx = torch.empty_like(xin)
    dw = _lib.zeros(C, xin.device)
    db = _lib.zeros(C, xin.device) if bdt is not None else None
    with torch.cuda.device(xin.device), _lib.timed("layernorm_bwd", 1, (2 + (dres is not None)) * xin.numel() * xin.element_size()
                                                       + dy.numel() * dy.element_size() + 8 * M):
        rc = _lib.lib().mlagg_layernorm_bwd_res(_lib.ptr(xin), _lib.ptr(w32), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(dy),
                                                _lib.ptr(dres), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db), M, C,
                                                _DT[xin.dtype], _DT[odt], _lib.stream_ptr())
    _lib.check(rc, "mlagg_layernorm_bwd_res")
    return dx.to(xdt), dw.to(wdt), None if db is None else db.to(bdt), None, None


class _LayerNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        return _ln_forward(ctx, x, weight, bias, eps, out_dtype)

    @staticmethod
    def backward(ctx, dy):
        return _ln_backward(ctx, dy, None)


class _LayerNormFork(torch.autograd.Function):
    """(LN(x), x) for a pre-norm residual branch `x + f(LN(x))`: the second output is x itself, to be used as the
    shortcut.  Both gradients then arrive at THIS node, and dx = LN'(d_ln) + d_shortcut is one kernel
    (mlagg_layernorm_bwd_res) instead of the LayerNorm backward plus autograd's accumulation add over the full tensor."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        return _ln_forward(ctx, x, weight, bias, eps, out_dtype), x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dshort):
        if dy is None:                       # the normalised output was not used
            return dshort, None, None, None, None
        return _ln_backward(ctx, dy, dshort)


def layer_norm_tokens(x, norm: torch.nn.LayerNorm, out_dtype=None):
    """LayerNorm over the last dim with `norm`'s parameters.  out_dtype None -> the dtype the consumer wants: the
    autocast dtype when autocast is on (the reference's fp32 result is down-cast by every consuming Linear anyway),
    else x.dtype."""
    if out_dtype is None:
        out_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
    C = x.shape[-1]
    if C % 4 != 0 or C > 1024 or norm.weight is None:
        return norm(x)
    return _LayerNorm.apply(x, norm.weight, norm.bias, norm.eps, out_dtype)


def layer_norm_fork(x, norm: torch.nn.LayerNorm, out_dtype=None):
    """(LayerNorm(x), shortcut) with shortcut == x: use the shortcut in the residual add that closes the branch."""
    if out_dtype is None:
        out_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
    C = x.shape[-1]
    if C % 4 != 0 or C > 1024 or norm.weight is None or not x.requires_grad:
        return layer_norm_tokens(x, norm, out_dtype), x
    return _LayerNormFork.apply(x, norm.weight, norm.bias, norm.eps, out_dtype)


def colsum(x2d):
    """fp32 column sums of a (M, C) fp32 / bf16 matrix with unit column stride (C ABI: mlagg_colsum)."""
    if not x2d.is_cuda:
        raise _lib.MlaggError("colsum: CUDA tensor required (no CPU fallback in the product path)")
    if x2d.dtype not in _DT or x2d.stride(1) != 1:
        x2d = _io(x2d).contiguous()
    M, C = x2d.shape
    out = _lib.zeros(C, x2d.device)
    with torch.cuda.device(x2d.device), _lib.timed("colsum"):
        rc = _lib.lib().mlagg_colsum(_lib.ptr(x2d), _lib.ptr(out), M, C, x2d.stride(0), _DT[x2d.dtype], _lib.stream_ptr())
    _lib.check(rc, "mlagg_colsum")
    return out


# ---- per-step cache of parameters already cast to the autocast dtype.  Autocast (and a plain `.to(bf16)` here) launches
# one cast kernel per weight and bias per forward -- ~330 tiny launches per train step.  The trainer instead refreshes
# ONE list of bf16 copies with a multi-tensor copy at the start of its step (`refresh_cast_cache`), and `_Linear` picks its
# operands from it while the step is running.  Outside a trainer step the cache is inactive and the cast happens inline.
_CAST = {"active": False, "dtype": None, "src": [], "dst": [], "map": {}}


def build_cast_cache(params, dtype):
    _CAST["src"] = [p for p in params if p.is_cuda and p.dtype == torch.float32]
    _CAST["dst"] = [torch.empty_like(p, dtype=dtype) for p in _CAST["src"]]
    _CAST["map"] = {id(p): d for p, d in zip(_CAST["src"], _CAST["dst"])}
    _CAST["dtype"] = dtype


def refresh_cast_cache():
    if _CAST["src"]:
        with torch.no_grad():
            torch._foreach_copy_(_CAST["dst"], _CAST["src"])
        _CAST["active"] = True


def release_cast_cache():
    _CAST["active"] = False


def _cast_param(t, dtype):
    if t is None:
        return None
    if _CAST["active"] and dtype == _CAST["dtype"]:
        base = t._base if t._base is not None else t           # views of a parameter (conv weight .view(C, C), .t())
        hit = _CAST["map"].get(id(base))
        if hit is not None:
            return hit if base is t else hit.as_strided(t.size(), t.stride(), t.storage_offset())
    return t.to(dtype)


def _rows2d(t):
    """(..., C) -> (M, C) VIEW with unit column stride and one uniform row stride, or None when that needs a copy."""
    if t.dim() < 2 or t.stride(-1) != 1:
        return None
    ld = t.stride(-2)
    if ld < t.shape[-1]:
        return None
    n = t.shape[-2]
    for d in range(t.dim() - 3, -1, -1):
        if t.shape[d] != 1 and t.stride(d) != n * ld:
            return None
        n *= t.shape[d]
    return t.as_strided((n, t.shape[-1]), (ld, 1), t.storage_offset())


class GradContiguous(torch.autograd.Function):
    """Identity whose backward hands on a CONTIGUOUS gradient.  The gradient arriving at a stage output comes from conv
    backward passes in whatever memory format they chose; left alone, every residual add / mask multiply of the stage
    inherits those strides (un-vectorised strided kernels, 5x slower) and every Linear backward re-packs its own copy."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g.contiguous()


class _Linear(torch.autograd.Function):
    """y = x W^T + b.  Under bf16 autocast all three GEMMs (forward, data gradient, weight gradient) run on the sm_100a
    tensor-core kernels of csrc/gemm_tc.cu (TMA + tcgen05.mma + TMEM) behind mlagg_linear_*: bias added from the fp32
    parameter in the epilogue, W consumed as stored for the data gradient (MN-major operand), fp32 weight gradient
    reduced across the token split.  fp32 (the 1e-4 parity path) stays on cuBLAS -- a library GEMM in full fp32; the
    bias gradient is ONE HBM-bound column-sum pass (csrc/reduce.cu) either way."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        cdt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
        xc, wc = x.to(cdt), _cast_param(weight, cdt)
        x2 = _rows2d(xc)
        ctx.tc = bool(cdt == torch.bfloat16 and x2 is not None and gemm.supported(x2, wc))
        if ctx.tc:
            y2, _ = gemm.linear_fwd(x2, wc, bias)
            y = y2.view(*xc.shape[:-1], wc.shape[0])
        else:
            bc = _cast_param(bias, cdt)
            with torch.autocast("cuda", enabled=False):
                if x2 is None:
                    y = torch.nn.functional.linear(xc, wc, bc)
                else:
                    # 2-D GEMM on the (tokens, Cin) view, row stride = the parent's width: channel slices of a wider
                    # activation (the halves of MLLABlock's `chunk`) are consumed in place and the bias stays in the GEMM
                    # epilogue (F.linear on a strided 3-D input runs matmul + a separate un-vectorised bias add)
                    y = (torch.mm(x2, wc.t()) if bc is None else torch.addmm(bc, x2, wc.t())).view(*xc.shape[:-1], wc.shape[0])
        ctx.save_for_backward(xc, wc)
        ctx.meta = (x.dtype, weight.dtype, None if bias is None else bias.dtype)
        # parameters whose gradients may be produced on the side stream and handed to the trainer directly (_lib.stash_grad)
        ctx.leaves = (_lib.leaf_param(weight), _lib.leaf_param(bias))
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, wc = ctx.saved_tensors
        xdt, wdt, bdt = ctx.meta
        Cout, Cin = wc.shape
        dy2 = dy.to(wc.dtype)
        dy2 = _rows2d(dy2) if _rows2d(dy2) is not None else dy2.reshape(-1, Cout)
        x2 = _rows2d(xc) if _rows2d(xc) is not None else xc.reshape(-1, Cin)
        if ctx.tc and gemm.rows_ok(dy2) and gemm.rows_ok(x2):
            wleaf, bleaf = ctx.leaves
            if (_lib.side_active() and ctx.needs_input_grad[1] and wleaf is not None and wleaf.dtype == torch.float32
                    and (bdt is None or (bleaf is not None and ctx.needs_input_grad[2] and bleaf.dtype == torch.float32))):
                # weight (and bias) gradient on the side stream, handed to the trainer; the data gradient -- which the rest
                # of the backward pass waits for -- on this stream, next to it
                if bdt is None:
                    _lib.stash_grad(wleaf, gemm.linear_bwd_weight(dy2, x2, side=True))
                else:
                    dw, db = gemm.linear_bwd_weight(dy2, x2, want_db=True, side=True)
                    _lib.stash_grad(wleaf, dw)
                    _lib.stash_grad(bleaf, db)
                dx = gemm.linear_bwd_data(dy2, wc).view(xc.shape).to(xdt) if ctx.needs_input_grad[0] else None
                return dx, None, None
            dx = gemm.linear_bwd_data(dy2, wc).view(xc.shape).to(xdt) if ctx.needs_input_grad[0] else None
            if ctx.needs_input_grad[1] and bdt is not None and ctx.needs_input_grad[2]:
                dw, db = gemm.linear_bwd_weight(dy2, x2, want_db=True)      # bias gradient from the same operand tiles
                return dx, dw.to(wdt), db.to(bdt)
            dw = gemm.linear_bwd_weight(dy2, x2).to(wdt) if ctx.needs_input_grad[1] else None
        else:
            with torch.autocast("cuda", enabled=False):
                dx = torch.mm(dy2, wc).view(xc.shape).to(xdt) if ctx.needs_input_grad[0] else None
                dw = torch.mm(dy2.t(), x2).to(wdt) if ctx.needs_input_grad[1] else None
        db = colsum(dy2).to(bdt) if (bdt is not None and ctx.needs_input_grad[2]) else None
        return dx, dw, db


class _MlpFused(torch.autograd.Function):
    """fc1 -> exact GELU -> fc2 (reference Mlp, nnUNetTrainer_MLAgg_2D_dt_MS.py:176-192; the 1x1-conv pair of MedNeXtBlock
    :230-324 has the same shape) as ONE autograd node on the tensor-core kernels: GELU is the epilogue of the fc1 GEMM
    (which also stores the pre-activation), and its gradient is the epilogue of fc2's data-gradient GEMM -- no
    element-wise kernel on the (tokens, hidden) tensor in either direction."""

    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2):
        x2 = _rows2d(x.to(torch.bfloat16))
        w1c, w2c = _cast_param(w1, torch.bfloat16), _cast_param(w2, torch.bfloat16)
        h, pre = gemm.linear_fwd(x2, w1c, b1, act="gelu", want_pre=True)
        y, _ = gemm.linear_fwd(h, w2c, b2)
        ctx.save_for_backward(x2, pre, h, w1c, w2c)
        ctx.meta = (x.dtype, x.shape, w1.dtype, w2.dtype, None if b1 is None else b1.dtype, None if b2 is None else b2.dtype)
        ctx.leaves = tuple(_lib.leaf_param(t) for t in (w1, b1, w2, b2))
        return y.view(*x.shape[:-1], w2c.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, pre, h, w1c, w2c = ctx.saved_tensors
        xdt, xshape, w1dt, w2dt, b1dt, b2dt = ctx.meta
        dy2 = dy.to(torch.bfloat16)
        dy2 = _rows2d(dy2) if _rows2d(dy2) is not None else dy2.reshape(-1, w2c.shape[0])
        if not gemm.rows_ok(dy2):
            dy2 = dy2.contiguous()
        need = ctx.needs_input_grad
        l1, lb1, l2, lb2 = ctx.leaves
        side = (_lib.side_active() and all(t is not None and t.dtype == torch.float32 for t in ctx.leaves) and all(need[1:]))
        dpre = gemm.linear_bwd_data(dy2, w2c, aux=pre, act="gelu")            # (dy W2) * gelu'(pre)
        dw2, db2 = gemm.linear_bwd_weight(dy2, h, want_db=True, side=side)
        dx = gemm.linear_bwd_data(dpre, w1c).view(xshape).to(xdt) if need[0] else None
        dw1, db1 = gemm.linear_bwd_weight(dpre, x2, want_db=True, side=side)
        if side:    # weight / bias gradients were produced on the side stream: handed to the trainer, not to autograd
            for leaf, gten in ((l1, dw1), (lb1, db1), (l2, dw2), (lb2, db2)):
                _lib.stash_grad(leaf, gten)
            return dx, None, None, None, None
        return (dx, dw1.to(w1dt), None if b1dt is None else db1.to(b1dt), dw2.to(w2dt), None if b2dt is None else db2.to(b2dt))


def mlp_gelu_tokens(x, w1, b1, w2, b2):
    """fc2(gelu(fc1(x))) for tokens-major x (..., C); weights (hidden, C) / (C_out, hidden) as nn.Linear stores them.
    The fused tensor-core node under bf16 autocast, else the two Linear nodes with torch's GELU between them."""
    if (x.is_cuda and torch.is_autocast_enabled("cuda") and torch.get_autocast_dtype("cuda") == torch.bfloat16
            and x.dim() >= 2 and w1.shape[0] % 8 == 0 and w1.shape[1] % 8 == 0 and w2.shape[0] % 8 == 0
            and w1.stride(-1) == 1 and w2.stride(-1) == 1):
        x16 = x.to(torch.bfloat16)
        x2 = _rows2d(x16)
        if x2 is not None and gemm.rows_ok(x2):
            return _MlpFused.apply(x, w1, b1, w2, b2)
    return _Linear.apply(torch.nn.functional.gelu(_Linear.apply(x, w1, b1)), w2, b2)


def linear_tokens(x, lin: torch.nn.Linear):
    """`lin(x)` for tokens-major activations; same numbers as nn.Linear under the surrounding autocast state."""
    if not x.is_cuda:
        raise _lib.MlaggError("linear_tokens: CUDA tensor required (no CPU fallback in the product path)")
    return _Linear.apply(x, lin.weight, lin.bias)


_ACT = {None: 0, "none": 0, "leaky_relu": 1, "silu": 2}


class _InstNorm(torch.autograd.Function):
    """C ABI: mlagg_instnorm_fwd / _bwd (csrc/instnorm.cu) on a channels_last (B, C, H, W) map."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, act, slope):
        Bn, C, H, W = x.shape
        xt = x.permute(0, 2, 3, 1)                                 # (B, H, W, C): contiguous when x is channels_last
        if not xt.is_contiguous():
            xt = xt.contiguous()
        w32 = None if weight is None else weight.detach().float().contiguous()
        b32 = None if bias is None else bias.detach().float().contiguous()
        y = torch.empty_like(xt)
        stats = torch.empty(Bn, C, 2, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device), _lib.timed("instnorm_fwd", 3, 3 * y.numel() * y.element_size()):
            rc = _lib.lib().mlagg_instnorm_fwd(_lib.ptr(xt), _lib.ptr(w32), _lib.ptr(b32), _lib.ptr(y), _lib.ptr(stats),
                                               Bn, H * W, C, float(eps), act, float(slope), _DT[xt.dtype], _lib.stream_ptr())
        _lib.check(rc, "mlagg_instnorm_fwd")
        ctx.save_for_backward(xt, w32, b32, stats)
        ctx.meta = (act, float(slope), None if weight is None else weight.dtype, None if bias is None else bias.dtype)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xt, w32, b32, stats = ctx.saved_tensors
        act, slope, wdt, bdt = ctx.meta
        Bn, H, W, C = xt.shape
        dyt = dy.to(xt.dtype).permute(0, 2, 3, 1)
        if not dyt.is_contiguous():
            dyt = dyt.contiguous()
        dx = torch.empty_like(xt)
        sums = torch.empty(Bn, C, 2, device=xt.device, dtype=torch.float32)
        dw = _lib.zeros(C, xt.device) if w32 is not None else None
        db = _lib.zeros(C, xt.device) if b32 is not None else None
        with torch.cuda.device(xt.device), _lib.timed("instnorm_bwd", 3, 5 * dx.numel() * dx.element_size()):
            rc = _lib.lib().mlagg_instnorm_bwd(_lib.ptr(xt), _lib.ptr(w32), _lib.ptr(b32), _lib.ptr(stats), _lib.ptr(dyt),
                                               _lib.ptr(dx), _lib.ptr(sums), _lib.ptr(dw), _lib.ptr(db), Bn, H * W, C, act,
                                               slope, _DT[xt.dtype], _lib.stream_ptr())
        _lib.check(rc, "mlagg_instnorm_bwd")
        return (dx.permute(0, 3, 1, 2), None if dw is None else dw.to(wdt), None if db is None else db.to(bdt),
                None, None, None)


class _InstNormRes(torch.autograd.Function):
    """C ABI: mlagg_instnorm_res_fwd / _bwd.  y = act(instance_norm(x) + residual) on channels_last (B, C, H, W) maps,
    act None | 'leaky_relu': the add and the activation of monai's UnetResBlock tail inside the normalisation's apply pass;
    the backward reads the activation's slope off the sign of the saved output."""

    @staticmethod
    def forward(ctx, x, residual, weight, bias, eps, act, slope):
        Bn, C, H, W = x.shape
        xt, rt = x.permute(0, 2, 3, 1), residual.permute(0, 2, 3, 1)
        w32 = None if weight is None else weight.detach().float().contiguous()
        b32 = None if bias is None else bias.detach().float().contiguous()
        y = torch.empty_like(xt)
        stats = torch.empty(Bn, C, 2, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device), _lib.timed("instnorm_fwd", 3, 4 * y.numel() * y.element_size()):
            rc = _lib.lib().mlagg_instnorm_res_fwd(_lib.ptr(xt), _lib.ptr(w32), _lib.ptr(b32), _lib.ptr(rt), _lib.ptr(y),
                                                   _lib.ptr(stats), Bn, H * W, C, float(eps), act, float(slope),
                                                   _DT[xt.dtype], _lib.stream_ptr())
        _lib.check(rc, "mlagg_instnorm_res_fwd")
        ctx.save_for_backward(xt, y, w32, b32, stats)
        ctx.meta = (act, float(slope), None if weight is None else weight.dtype, None if bias is None else bias.dtype)
        return y.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        xt, y, w32, b32, stats = ctx.saved_tensors
        act, slope, wdt, bdt = ctx.meta
        Bn, H, W, C = xt.shape
        dyt = dy.to(xt.dtype).permute(0, 2, 3, 1)
        if not dyt.is_contiguous():
            dyt = dyt.contiguous()
        dx, dres = torch.empty_like(xt), torch.empty_like(xt)
        sums = torch.empty(Bn, C, 2, device=xt.device, dtype=torch.float32)
        dw = _lib.zeros(C, xt.device) if w32 is not None else None
        db = _lib.zeros(C, xt.device) if b32 is not None else None
        with torch.cuda.device(xt.device), _lib.timed("instnorm_bwd", 3, 8 * dx.numel() * dx.element_size()):
            rc = _lib.lib().mlagg_instnorm_res_bwd(_lib.ptr(xt), _lib.ptr(w32), _lib.ptr(b32), _lib.ptr(stats), _lib.ptr(y),
                                                   _lib.ptr(dyt), _lib.ptr(dx), _lib.ptr(dres), _lib.ptr(sums), _lib.ptr(dw),
                                                   _lib.ptr(db), Bn, H * W, C, act, slope, _DT[xt.dtype], _lib.stream_ptr())
        _lib.check(rc, "mlagg_instnorm_res_bwd")
        return (dx.permute(0, 3, 1, 2), dres.permute(0, 3, 1, 2), None if dw is None else dw.to(wdt),
                None if db is None else db.to(bdt), None, None, None)


def instance_norm_res_cl(x, residual, weight=None, bias=None, eps=1e-5, act="leaky_relu", slope=0.01):
    """act(instance_norm(x) + residual) as one node when the row-streaming kernels take the shape (channels_last CUDA maps
    of one dtype, C a multiple of the 16-byte vector); None otherwise -- the caller then adds and activates itself."""
    if not (supports_instance_norm_cl(x) and residual.shape == x.shape and residual.dtype == x.dtype and act in (None, "leaky_relu")):
        return None
    C = x.shape[1]
    v = 16 // x.element_size()
    if C % v or C // v > 256 or not x.permute(0, 2, 3, 1).is_contiguous() or not residual.permute(0, 2, 3, 1).is_contiguous():
        return None
    return _InstNormRes.apply(x, residual, weight, bias, eps, _ACT[act], slope)


def instance_norm_cl(x, weight=None, bias=None, eps=1e-5, act=None, slope=0.01):
    """Per-(image, channel) normalisation of a (B, C, H, W) map kept in channels_last memory, optional affine and a
    fused activation (None | 'leaky_relu' | 'silu').  Equals nn.InstanceNorm2d (training statistics) and
    nn.GroupNorm(num_groups=C).  Shapes / dtypes the kernel does not take (C % 4 != 0, fp16, CPU) are the caller's
    business -- see `supports_instance_norm_cl`."""
    return _InstNorm.apply(x, weight, bias, eps, _ACT[act], slope)


def supports_instance_norm_cl(x):
    return x.is_cuda and x.dim() == 4 and x.shape[1] % 4 == 0 and x.dtype in _DT


class _AvgPoolTokens(torch.autograd.Function):
    """C ABI: mlagg_avgpool_tokens_fwd / _bwd (csrc/pool.cu)."""

    @staticmethod
    def forward(ctx, x, H, W, pH, pW, gelu):
        if not x.is_cuda:
            raise _lib.MlaggError("avgpool_tokens: CUDA tensor required (no CPU fallback in the product path)")
        Bn, N, C = x.shape
        assert N == H * W
        xin = _io(x).contiguous()
        y = torch.empty(Bn, pH * pW, C, device=x.device, dtype=xin.dtype)
        with torch.cuda.device(x.device), _lib.timed("avgpool_fwd"):
            rc = _lib.lib().mlagg_avgpool_tokens_fwd(_lib.ptr(xin), _lib.ptr(y), Bn, H, W, C, pH, pW, int(gelu),
                                                     _DT[xin.dtype], _lib.stream_ptr())
        _lib.check(rc, "mlagg_avgpool_tokens_fwd")
        ctx.save_for_backward(xin)
        ctx.meta = (H, W, pH, pW, bool(gelu), x.dtype)
        return y.to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        (xin,) = ctx.saved_tensors
        H, W, pH, pW, gelu, xdt = ctx.meta
        Bn, N, C = xin.shape
        dy = dy.to(xin.dtype).contiguous()
        dx = torch.empty_like(xin)
        with torch.cuda.device(xin.device), _lib.timed("avgpool_bwd"):
            rc = _lib.lib().mlagg_avgpool_tokens_bwd(_lib.ptr(xin), _lib.ptr(dy), _lib.ptr(dx), Bn, H, W, C, pH, pW,
                                                     int(gelu), _DT[xin.dtype], _lib.stream_ptr())
        _lib.check(rc, "mlagg_avgpool_tokens_bwd")
        return dx.to(xdt), None, None, None, None, None


def avgpool_tokens(x, H, W, pH, pW, gelu=False):
    """x (B, H*W, C) -> adaptive average pool to (B, pH*pW, C) of gelu(x) (gelu=True) or x, torch bin semantics."""
    return _AvgPoolTokens.apply(x, H, W, pH, pW, gelu)


def _ew_ok(*ts):
    """operands the element-wise kernels take directly: CUDA, same fp32 / bf16 dtype and shape, contiguous, numel % 4 == 0"""
    t0 = ts[0]
    return (t0.is_cuda and t0.dtype in _DT and t0.numel() % 4 == 0 and t0.numel() > 0
            and all(t.dtype == t0.dtype and t.shape == t0.shape and t.is_contiguous() for t in ts))


class _ResidualScale(torch.autograd.Function):
    """C ABI: mlagg_residual_scale.  out = x + scale[b] * y."""

    @staticmethod
    def forward(ctx, x, y, scale):
        out = torch.empty_like(x)
        n = x.numel()
        with torch.cuda.device(x.device), _lib.timed("residual_scale"):
            rc = _lib.lib().mlagg_residual_scale(_lib.ptr(x), _lib.ptr(y), _lib.ptr(scale), _lib.ptr(out), n,
                                                 n // x.shape[0], _DT[x.dtype], _lib.stream_ptr())
        _lib.check(rc, "mlagg_residual_scale")
        ctx.save_for_backward(scale)
        return out

    @staticmethod
    def backward(ctx, g):
        (scale,) = ctx.saved_tensors
        if not _ew_ok(g):
            g = g.contiguous()
        n = g.numel()
        dy = torch.empty_like(g)
        with torch.cuda.device(g.device), _lib.timed("residual_scale"):
            rc = _lib.lib().mlagg_residual_scale(None, _lib.ptr(g), _lib.ptr(scale), _lib.ptr(dy), n, n // g.shape[0],
                                                 _DT[g.dtype], _lib.stream_ptr())
        _lib.check(rc, "mlagg_residual_scale")
        return g, dy, None


def residual_drop_path(x, y, drop_path):
    """`x + drop_path(y)` (reference nnUNetTrainer_MLAgg_2D_dt_MS.py:907-908, MambaSkip.py:733,745): a plain add when the
    module is the identity, else ONE pass with the per-sample keep mask / keep as a (B) fp32 vector."""
    p = float(getattr(drop_path, "drop_prob", 0.0) or 0.0)
    if p == 0.0 or not drop_path.training:
        return x + drop_path(y)
    if not (_ew_ok(x, y) and (x.numel() // x.shape[0]) % 4 == 0):
        return x + drop_path(y)
    keep = 1.0 - p
    scale = torch.empty(x.shape[0], device=x.device, dtype=torch.float32).bernoulli_(keep)
    if keep > 0.0 and getattr(drop_path, "scale_by_keep", True):
        scale.div_(keep)
    return _ResidualScale.apply(x, y, scale)


class _SiluGate(torch.autograd.Function):
    """C ABI: mlagg_silu_gate_fwd / _bwd.  out = t * silu(z)."""

    @staticmethod
    def forward(ctx, t, z):
        out = torch.empty_like(t)
        with torch.cuda.device(t.device), _lib.timed("silu_gate_fwd"):
            rc = _lib.lib().mlagg_silu_gate_fwd(_lib.ptr(t), _lib.ptr(z), _lib.ptr(out), t.numel(), _DT[t.dtype],
                                                _lib.stream_ptr())
        _lib.check(rc, "mlagg_silu_gate_fwd")
        ctx.save_for_backward(t, z)
        return out

    @staticmethod
    def backward(ctx, g):
        t, z = ctx.saved_tensors
        g = g.to(t.dtype).contiguous()
        dt, dz = torch.empty_like(t), torch.empty_like(z)
        with torch.cuda.device(t.device), _lib.timed("silu_gate_bwd"):
            rc = _lib.lib().mlagg_silu_gate_bwd(_lib.ptr(t), _lib.ptr(z), _lib.ptr(g), _lib.ptr(dt), _lib.ptr(dz),
                                                t.numel(), _DT[t.dtype], _lib.stream_ptr())
        _lib.check(rc, "mlagg_silu_gate_bwd")
        return dt, dz


def silu_gate(t, z):
    """t * silu(z) in one pass (the block's output gate, reference :881 / :907)."""
    if z.dtype != t.dtype:
        z = z.to(t.dtype)
    if not _ew_ok(t, z):
        return t * torch.nn.functional.silu(z)
    return _SiluGate.apply(t, z)


class _BiasAddCL(torch.autograd.Function):
    """C ABI: mlagg_bias_add_cl (in place on the fresh conv output) / mlagg_colsum for the gradient."""

    @staticmethod
    def forward(ctx, y, bias):
        C = y.shape[1]
        b32 = bias.detach().float().contiguous()
        with torch.cuda.device(y.device), _lib.timed("bias_add_cl"):
            rc = _lib.lib().mlagg_bias_add_cl(_lib.ptr(y), _lib.ptr(b32), y.numel(), C, _DT[y.dtype], _lib.stream_ptr())
        _lib.check(rc, "mlagg_bias_add_cl")
        ctx.mark_dirty(y)
        ctx.bdt = bias.dtype
        return y

    @staticmethod
    def backward(ctx, g):
        C = g.shape[1]
        gt = g.permute(0, 2, 3, 1)
        if gt.is_contiguous() and g.dtype in _DT:
            db = colsum(gt.reshape(-1, C))
        else:
            db = g.float().sum((0, 2, 3))
        return g, db.to(ctx.bdt)


def _conv_bias_cl(y, bias):
    """y (B, C, H, W) = bias-free conv output; adds `bias` per channel.  Channels_last fp32 / bf16 CUDA maps with C % 4 == 0
    take the one-pass kernel (and the column-sum bias gradient); anything else the plain broadcast add."""
    if (bias is not None and y.is_cuda and y.dim() == 4 and y.dtype in _DT and y.shape[1] % 4 == 0 and y.numel() > 0
            and y.permute(0, 2, 3, 1).is_contiguous() and y._base is None):
        return _BiasAddCL.apply(y, bias)
    return y if bias is None else y + bias.to(y.dtype).view(1, -1, 1, 1)


class _Conv2dSplit(torch.autograd.Function):
    """Bias-free cuDNN conv2d (the conv stages, off the named path) whose backward is issued as TWO library calls: the
    input gradient on the current stream -- the rest of the backward pass waits for it -- and the weight gradient, which
    only the optimizer reads, on the side stream of `_lib.side_launch` next to it."""

    @staticmethod
    def forward(ctx, x, weight, stride, padding, dilation, groups):
        cdt = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else x.dtype
        xc, wc = x.to(cdt), _cast_param(weight, cdt)
        with torch.autocast("cuda", enabled=False):
            y = torch.nn.functional.conv2d(xc, wc, None, stride, padding, dilation, groups)
        ctx.save_for_backward(xc, wc)
        ctx.conf = (tuple(stride), tuple(padding), tuple(dilation), groups, x.dtype, weight.dtype)
        ctx.wleaf = _lib.leaf_param(weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, wc = ctx.saved_tensors
        stride, padding, dilation, groups, xdt, wdt = ctx.conf
        dy = dy.to(wc.dtype)
        bwd = torch.ops.aten.convolution_backward
        dx = dw = None
        if ctx.needs_input_grad[1]:
            if _lib.side_active() and ctx.wleaf is not None and ctx.wleaf.dtype == wdt:
                with _lib.side_launch(dy, xc, wc):   # handed to the trainer directly: never through autograd (see _lib)
                    _lib.stash_grad(ctx.wleaf, bwd(dy, xc, wc, None, stride, padding, dilation, False, [0, 0], groups,
                                                   [False, True, False])[1].to(wdt))
            else:
                dw = bwd(dy, xc, wc, None, stride, padding, dilation, False, [0, 0], groups, [False, True, False])[1].to(wdt)
        if ctx.needs_input_grad[0]:
            dx = bwd(dy, xc, wc, None, stride, padding, dilation, False, [0, 0], groups, [True, False, False])[0].to(xdt)
        return dx, dw, None, None, None, None


def conv2d_split(conv, x):
    """conv(x) without its bias for an nn.Conv2d with zero padding; plain `_conv_forward` when gradients are off / on CPU"""
    if x.dim() == 4 and x.shape[1] == 1 and x.is_cuda and x.is_contiguous():
        # a one-channel map has ambiguous strides and torch reads them as NCHW: cuDNN then returns an NCHW result that every
        # channels_last consumer (and the gradient coming back) has to re-lay out.  Spell the NHWC strides out (a view).
        x = x.as_strided(x.size(), (x.stride(0), 1, x.stride(2), x.stride(3)))
    if (x.is_cuda and torch.is_grad_enabled() and conv.padding_mode == "zeros" and not isinstance(conv.padding, str)
            and (x.requires_grad or conv.weight.requires_grad)):
        return _Conv2dSplit.apply(x, conv.weight, conv.stride, conv.padding, conv.dilation, conv.groups)
    return conv._conv_forward(x, conv.weight, None)


class Conv2dCL(torch.nn.Conv2d):
    """nn.Conv2d (same parameters / state_dict) whose bias is applied by `_conv_bias_cl` after a bias-free cuDNN call."""

    def forward(self, x):
        if self.bias is None or not x.is_cuda:
            return super().forward(x)
        return _conv_bias_cl(conv2d_split(self, x), self.bias)


class ConvTranspose2dCL(torch.nn.ConvTranspose2d):
    def forward(self, x, output_size=None):
        if self.bias is None or not x.is_cuda or output_size is not None:
            return super().forward(x, output_size)
        y = torch.nn.functional.conv_transpose2d(x, self.weight, None, self.stride, self.padding, self.output_padding,
                                                 self.groups, self.dilation)
        return _conv_bias_cl(y, self.bias)


# ---- strided row copies (C ABI: mlagg_copy_rows) for the channel / stage splits and joins around the MSMM -------------------
def _rows3(t):
    """t (B, R, C) or (B, H, W, C) view with unit channel stride whose pixel dims collapse to one row stride ->
    (ptr-carrying tensor, B, R, C, ld, bs) or None"""
    if t.dim() == 4:
        if t.stride(3) != 1 or t.stride(1) != t.shape[2] * t.stride(2):
            return None
        return t, t.shape[0], t.shape[1] * t.shape[2], t.shape[3], t.stride(2), t.stride(0)
    if t.dim() == 3 and t.stride(2) == 1:
        return t, t.shape[0], t.shape[1], t.shape[2], t.stride(1), t.stride(0)
    return None


def copy_rows_(dst, src):
    """dst[...] = src[...] for two equally shaped row-strided views (see `_rows3`); falls back to Tensor.copy_ for
    anything the kernel does not address (other dtypes, CPU, non-collapsible views)."""
    a, b = _rows3(dst), _rows3(src)
    if (a is None or b is None or not dst.is_cuda or dst.dtype not in _DT or src.dtype != dst.dtype
            or a[1:4] != b[1:4] or dst.numel() == 0 or min(a[4], b[4]) < a[3]):
        dst.copy_(src)
        return dst
    with torch.cuda.device(dst.device), _lib.timed("copy_rows"):
        rc = _lib.lib().mlagg_copy_rows(src.data_ptr(), b[4], b[5], dst.data_ptr(), a[4], a[5], a[1], a[2], a[3],
                                        _DT[dst.dtype], _lib.stream_ptr())
    _lib.check(rc, "mlagg_copy_rows")
    return dst


def add_rows_(dst, src):
    """dst[...] += src[...] for two equally shaped row-strided views (C ABI: mlagg_add_rows); Tensor.add_ for anything the
    kernel does not address."""
    a, b = _rows3(dst), _rows3(src)
    if (a is not None and b is not None and dst.is_cuda and dst.dtype in _DT and src.dtype == dst.dtype and a[1:4] == b[1:4]
            and dst.numel() > 0 and min(a[4], b[4]) >= a[3]):
        with torch.cuda.device(dst.device), _lib.timed("add_rows"):
            rc = _lib.lib().mlagg_add_rows(src.data_ptr(), b[4], b[5], dst.data_ptr(), a[4], a[5], a[1], a[2], a[3],
                                           _DT[dst.dtype], _lib.stream_ptr())
        if rc == 0:
            return dst
    dst.add_(src)
    return dst


class SplitLast(torch.autograd.Function):
    """x[..., :k], x[..., k:] as views; the backward assembles ONE gradient with two strided row copies.  Autograd's own
    slice gradients are two zero-filled full-size tensors (contiguous in the logical NCHW order, which then drags every
    accumulation into the stage output's channels_last gradient onto strided kernels) plus two copies and an add."""

    @staticmethod
    def forward(ctx, x, k):
        ctx.k = k
        return x[..., :k], x[..., k:]

    @staticmethod
    def backward(ctx, ga, gb):
        out = torch.empty(ga.shape[:-1] + (ga.shape[-1] + gb.shape[-1],), device=ga.device, dtype=ga.dtype)
        copy_rows_(out[..., :ctx.k], ga)
        copy_rows_(out[..., ctx.k:], gb.to(ga.dtype))
        return out, None


class JoinLast(torch.autograd.Function):
    """torch.cat([a, b], dim=-1) of two (B, H, W, .) maps with strided row copies; the gradients are views."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.k = a.shape[-1]
        out = torch.empty(a.shape[:-1] + (a.shape[-1] + b.shape[-1],), device=a.device, dtype=a.dtype)
        copy_rows_(out[..., :ctx.k], a)
        copy_rows_(out[..., ctx.k:], b.to(a.dtype))
        return out

    @staticmethod
    def backward(ctx, g):
        return g[..., :ctx.k], g[..., ctx.k:]


class JoinLastDense(JoinLast):
    """JoinLast whose gradients are DENSE copies (two strided row copies): for consumers that need packed tensors anyway
    (cuDNN convolutions re-pack a channel-slice view with torch's generic strided-copy kernel, ~4x slower)."""

    @staticmethod
    def backward(ctx, g):
        ga = torch.empty(g.shape[:-1] + (ctx.k,), device=g.device, dtype=g.dtype)
        gb = torch.empty(g.shape[:-1] + (g.shape[-1] - ctx.k,), device=g.device, dtype=g.dtype)
        copy_rows_(ga, g[..., :ctx.k])
        copy_rows_(gb, g[..., ctx.k:])
        return ga, gb


def _copy_rows_raw(src, src_off, ld_s, bs_s, dst, dst_off, ld_d, bs_d, batch, rows, cols):
    """mlagg_copy_rows on explicit (element offset, row stride, batch stride) addressing of two same-dtype tensors"""
    es = src.element_size()
    with torch.cuda.device(src.device), _lib.timed("copy_rows"):
        rc = _lib.lib().mlagg_copy_rows(src.data_ptr() + src_off * es, ld_s, bs_s, dst.data_ptr() + dst_off * es, ld_d, bs_d,
                                        batch, rows, cols, _DT[src.dtype], _lib.stream_ptr())
    _lib.check(rc, "mlagg_copy_rows")


class UpShuffleJoin(torch.autograd.Function):
    """Pixel-shuffle of a 2x2 / stride-2 transposed convolution computed as a per-token GEMM, joined with the skip map:
         y    (B, H W, 4 Co)   columns ordered (di, dj, co) -- the GEMM output of `x . W[ci, co, di, dj]`
         skip (B, 2H, 2W, Cs)  channels_last view
         out  (B, 2H, 2W, Co + Cs):  out[b, 2i+di, 2j+dj, :Co] = y[b, i W + j, (2 di + dj) Co : ...],  out[..., Co:] = skip
    Five strided row copies forward, five backward (dense gradients).  Replaces UnetrUpBlock's cuDNN transposed
    convolution (an sm_75-era kernel with an NCHW result), the layout conversion of that result and torch.cat -- the
    single largest piece of torch glue in the step (191 us forward, 3 x ~140 us backward at 10 x 96 x 320 x 320)."""

    @staticmethod
    def forward(ctx, y, skip, H, W):
        Bn, Co, Cs = y.shape[0], y.shape[2] // 4, skip.shape[-1]
        ld = Co + Cs
        out = torch.empty(Bn, 2 * H, 2 * W, ld, device=y.device, dtype=y.dtype)
        y = y.contiguous()
        for di in range(2):
            for dj in range(2):
                _copy_rows_raw(y, (2 * di + dj) * Co, 4 * Co, W * 4 * Co, out, (di * 2 * W + dj) * ld, 2 * ld, 4 * W * ld,
                               Bn * H, W, Co)
        copy_rows_(out[..., Co:], skip.to(y.dtype))
        ctx.meta = (H, W, Co, Cs, skip.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        H, W, Co, Cs, sdt = ctx.meta
        g = g.contiguous()
        Bn, ld = g.shape[0], Co + Cs
        gy = torch.empty(Bn, H * W, 4 * Co, device=g.device, dtype=g.dtype)
        for di in range(2):
            for dj in range(2):
                _copy_rows_raw(g, (di * 2 * W + dj) * ld, 2 * ld, 4 * W * ld, gy, (2 * di + dj) * Co, 4 * Co, W * 4 * Co,
                               Bn * H, W, Co)
        gs = torch.empty(Bn, 2 * H, 2 * W, Cs, device=g.device, dtype=g.dtype)
        copy_rows_(gs, g[..., Co:])
        return gy, gs.to(sdt), None, None


class PadTopLeftAdd(torch.autograd.Function):
    """F.pad(a, (1, 0, 1, 0)) + F.pad(b, (1, 0, 1, 0)) for two channels_last maps (PatchExpand, reference
    nnUNetTrainer_MLAgg_2D_dt_MS.py:520-546): one zero fill and one add into the interior instead of two fills, two
    strided copies and an add; the backward hands ONE dense copy of the interior gradient to both branches."""

    @staticmethod
    def forward(ctx, a, b):
        Bn, C, H, W = a.shape
        out = torch.empty(Bn, C, H + 1, W + 1, device=a.device, dtype=a.dtype, memory_format=torch.channels_last).zero_()
        if b is None:
            out[:, :, 1:, 1:].copy_(a)
        else:
            torch.add(a, b, out=out[:, :, 1:, 1:])
        ctx.two = b is not None
        return out

    @staticmethod
    def backward(ctx, g):
        gi = g[:, :, 1:, 1:].contiguous(memory_format=torch.channels_last)
        return gi, (gi if ctx.two else None)


class SplitKV(torch.autograd.Function):
    """kv -> (kv, v) with v = kv[..., C:] for the LePE branch (reference :690-691, :716, :759).  Both gradients arrive at
    this node: the LePE gradient is added IN PLACE into the v half of the attention core's kv gradient (one half-size
    add) instead of autograd's zero-filled full-size tensor + slice copy + full-size add."""

    @staticmethod
    def forward(ctx, kv, C):
        ctx.C = C
        return kv.view_as(kv), kv[..., C:]

    @staticmethod
    def backward(ctx, g_kv, g_v):
        if g_kv is None:
            g_kv = torch.zeros(g_v.shape[:-1] + (g_v.shape[-1] + ctx.C,), device=g_v.device, dtype=g_v.dtype)
        if g_v is not None:
            if not g_kv.is_contiguous():
                g_kv = g_kv.contiguous()
            add_rows_(g_kv[..., ctx.C:], g_v.to(g_kv.dtype))
        return g_kv, None


class CatStages(torch.autograd.Function):
    """per-stage token maps (B, L_s, C) (row-strided views allowed) -> the stage-concatenated sequence (B, sum L_s, C);
    the gradients are views of the sequence gradient."""

    @staticmethod
    def forward(ctx, *parts):
        Bn, C = parts[0].shape[0], parts[0].shape[2]
        ctx.lens = [p.shape[1] for p in parts]
        out = torch.empty(Bn, sum(ctx.lens), C, device=parts[0].device, dtype=parts[0].dtype)
        off = 0
        for p, n in zip(parts, ctx.lens):
            copy_rows_(out[:, off:off + n], p.to(out.dtype))
            off += n
        return out

    @staticmethod
    def backward(ctx, g):
        outs, off = [], 0
        for n in ctx.lens:
            outs.append(g[:, off:off + n])
            off += n
        return tuple(outs)


class SplitStages(torch.autograd.Function):
    """the inverse: (B, sum L_s, C) -> packed per-stage (B, L_s, C) tensors (uniform row stride: the per-token GEMMs keep
    their bias epilogue); the backward writes the stage gradients into ONE sequence gradient (no zero-filled
    full-size tensor per stage, no adds)."""

    @staticmethod
    def forward(ctx, x, lens):
        ctx.lens = tuple(lens)
        outs, off = [], 0
        for n in ctx.lens:
            o = torch.empty(x.shape[0], n, x.shape[2], device=x.device, dtype=x.dtype)
            copy_rows_(o, x[:, off:off + n])
            outs.append(o)
            off += n
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gs):
        Bn, C = gs[0].shape[0], gs[0].shape[2]
        out = torch.empty(Bn, sum(ctx.lens), C, device=gs[0].device, dtype=gs[0].dtype)
        off = 0
        for g, n in zip(gs, ctx.lens):
            copy_rows_(out[:, off:off + n], g.to(out.dtype))
            off += n
        return out, None
