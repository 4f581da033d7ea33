"""Drop-in for `mamba_ssm.ops.selective_scan_interface.selective_scan_fn` backed by the sm_100a kernels.

Reference call site: mlagg/nnunetv2/training/nnUNetTrainer/variants/mamba/MambaSkip.py:445-451
(`selective_scan_fn(xs, dts, As, Bs, Cs, Ds, z=None, delta_bias=..., delta_softplus=True,
return_last_state=False)`); FFI being replaced: selective_scan_cuda.fwd / .bwd (vmamba/csms6s.py:224, :235).

Same positional/keyword signature and the same conventions: u, delta (B, D, L); A (D, N) real; B, C
(B, N, L) or (B, G, N, L); D, delta_bias (D,) fp32; the scan runs in fp32 and returns u.dtype.
CUDA only -- there is no CPU path here (the CPU restatement lives in oracle/ and is test-only).
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn.functional as F

from . import _lib


def _f32c(t):
    return None if t is None else t.detach().float().contiguous()


def _wants_grad(*ts):
    """Decided OUTSIDE Function.apply: inside `forward` grad mode is always off and Parameters passed in directly keep
    requires_grad = True under torch.no_grad(), which would make inference write (and allocate) the state checkpoints."""
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


class SelectiveScanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, return_last_state=False,
                need_grad=True):
        if not u.is_cuda:
            raise _lib.MlaggError("selective_scan_fn: CUDA tensors required (no CPU fallback in the product path)")
        if A.is_complex() or B.dim() not in (3, 4) or C.dim() != B.dim():
            raise _lib.MlaggError("selective_scan_fn: real A and time-varying B, C of rank 3 or 4 are supported")
        ctx.in_dtypes = (u.dtype, delta.dtype, A.dtype, B.dtype, C.dtype,
                         None if D is None else D.dtype, None if delta_bias is None else delta_bias.dtype)
        ctx.squeeze = B.dim() == 3
        u32, dl32, A32, B32, C32, D32, b32 = map(_f32c, (u, delta, A, B, C, D, delta_bias))
        if ctx.squeeze:
            B32, C32 = B32.unsqueeze(1), C32.unsqueeze(1)
        Bn, Dm, L = u32.shape
        N, G = A32.shape[1], B32.shape[1]
        L_ = _lib.lib()
        out = torch.empty_like(u32)
        ckpt = None
        if need_grad:
            ckpt = torch.empty(L_.mlagg_scan_ckpt_bytes(Bn, Dm, L, N) // 4, device=u.device, dtype=torch.float32)
        last = torch.empty(Bn, Dm, N, device=u.device, dtype=torch.float32) if return_last_state else None
        with torch.cuda.device(u.device), _lib.timed("scan_fwd"):
            rc = L_.mlagg_selective_scan_fwd(_lib.ptr(u32), _lib.ptr(dl32), _lib.ptr(A32), _lib.ptr(B32),
                                             _lib.ptr(C32), _lib.ptr(D32), _lib.ptr(b32), _lib.ptr(out),
                                             _lib.ptr(ckpt), _lib.ptr(last), Bn, Dm, L, N, G,
                                             int(bool(delta_softplus)), _lib.stream_ptr())
        _lib.check(rc, "mlagg_selective_scan_fwd")
        ctx.delta_softplus = bool(delta_softplus)
        ctx.has = (D is not None, delta_bias is not None)
        ctx.save_for_backward(u32, dl32, A32, B32, C32, D32, b32, ckpt)
        out = out.to(u.dtype)
        if return_last_state:
            ctx.mark_non_differentiable(last)
            return out, last
        return out

    @staticmethod
    def backward(ctx, dout, *unused):
        u, delta, A, B, C, D, bias, ckpt = ctx.saved_tensors
        Bn, Dm, L = u.shape
        N, G = A.shape[1], B.shape[1]
        dout = dout.float().contiguous()
        du, dd = torch.empty_like(u), torch.empty_like(u)
        dA, dB, dC = torch.zeros_like(A), torch.zeros_like(B), torch.zeros_like(C)
        dD = torch.zeros_like(D) if D is not None else None
        db = torch.zeros_like(bias) if bias is not None else None
        L_ = _lib.lib()
        with torch.cuda.device(u.device), _lib.timed("scan_bwd"):
            rc = L_.mlagg_selective_scan_bwd(_lib.ptr(u), _lib.ptr(delta), _lib.ptr(A), _lib.ptr(B), _lib.ptr(C),
                                             _lib.ptr(D), _lib.ptr(bias), _lib.ptr(dout), _lib.ptr(ckpt),
                                             _lib.ptr(du), _lib.ptr(dd), _lib.ptr(dA), _lib.ptr(dB), _lib.ptr(dC),
                                             _lib.ptr(dD), _lib.ptr(db), Bn, Dm, L, N, G,
                                             int(ctx.delta_softplus), _lib.stream_ptr())
        _lib.check(rc, "mlagg_selective_scan_bwd")
        if ctx.squeeze:
            dB, dC = dB.squeeze(1), dC.squeeze(1)
        dt = ctx.in_dtypes
        cast = lambda g, d: None if g is None else g.to(d)
        return (cast(du, dt[0]), cast(dd, dt[1]), cast(dA, dt[2]), cast(dB, dt[3]), cast(dC, dt[4]),
                cast(dD, dt[5]), cast(db, dt[6]), None, None, None)


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False):
    """out = S6 scan of u (plus D*u), optionally gated by silu(z); see module docstring."""
    res = SelectiveScanFn.apply(u, delta, A, B, C, D, delta_bias, delta_softplus, return_last_state,
                                _wants_grad(u, delta, A, B, C, D, delta_bias))
    out, last = res if return_last_state else (res, None)
    if z is not None:
        out = out * F.silu(z)
    return (out, last) if return_last_state else out


# --------------------------------------------------------------------------------------------------------------------
# Fused MSMM scan (C ABI: mlagg_msmm_scan_fwd / _bwd): SS2D_skip.forward_corev0 without the materialised cross-scan
# --------------------------------------------------------------------------------------------------------------------
class MSMMScanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xrow, xcol, xdbl_row, xdbl_col, Wdt, dt_bias, A, Ds, stage_lens, need_grad=True):
        if not xrow.is_cuda:
            raise _lib.MlaggError("msmm_scan: CUDA tensors required (no CPU fallback in the product path)")
        ctx.in_dtypes = tuple(t.dtype for t in (xrow, xcol, xdbl_row, xdbl_col, Wdt, dt_bias, A, Ds))
        xrow_, xcol_, xr_, xc_, W_, b_, A_, D_ = map(_f32c, (xrow, xcol, xdbl_row, xdbl_col, Wdt, dt_bias, A, Ds))
        Bn, Di, L = xrow_.shape
        N, R = A_.shape[1], W_.shape[1]
        assert sum(stage_lens) == L and xr_.shape == (Bn, 2, R + 2 * N, L) and xc_.shape == xr_.shape
        lens = (ctypes.c_int * len(stage_lens))(*[int(v) for v in stage_lens])
        L_ = _lib.lib()
        out = torch.empty(Bn, 4, Di, L, device=xrow.device, dtype=torch.float32)
        ckpt = None
        if need_grad:
            ckpt = torch.empty(L_.mlagg_scan_ckpt_bytes(Bn, 4 * Di, L, N) // 4, device=xrow.device, dtype=torch.float32)
        with torch.cuda.device(xrow.device), _lib.timed("scan_fwd"):
            rc = L_.mlagg_msmm_scan_fwd(_lib.ptr(xrow_), _lib.ptr(xcol_), _lib.ptr(xr_), _lib.ptr(xc_), _lib.ptr(W_),
                                        _lib.ptr(b_), _lib.ptr(A_), _lib.ptr(D_), _lib.ptr(out), _lib.ptr(ckpt),
                                        Bn, Di, N, R, len(stage_lens), lens, _lib.stream_ptr())
        _lib.check(rc, "mlagg_msmm_scan_fwd")
        ctx.stage_lens = tuple(int(v) for v in stage_lens)
        ctx.save_for_backward(xrow_, xcol_, xr_, xc_, W_, b_, A_, D_, ckpt)
        return out

    @staticmethod
    def backward(ctx, dout):
        xrow, xcol, xr, xc, W, b, A, D, ckpt = ctx.saved_tensors
        Bn, Di, L = xrow.shape
        N, R = A.shape[1], W.shape[1]
        dout = dout.float().contiguous()
        du = torch.empty(Bn, 4, Di, L, device=xrow.device, dtype=torch.float32)
        dxr, dxc = torch.zeros_like(xr), torch.zeros_like(xc)
        dW, db, dA, dD = (_lib.zeros(t.shape, t.device) for t in (W, b, A, D))
        lens = (ctypes.c_int * len(ctx.stage_lens))(*ctx.stage_lens)
        L_ = _lib.lib()
        with torch.cuda.device(xrow.device), _lib.timed("scan_bwd"):
            rc = L_.mlagg_msmm_scan_bwd(_lib.ptr(xrow), _lib.ptr(xcol), _lib.ptr(xr), _lib.ptr(xc), _lib.ptr(W),
                                        _lib.ptr(b), _lib.ptr(A), _lib.ptr(D), _lib.ptr(dout), _lib.ptr(ckpt),
                                        _lib.ptr(du), _lib.ptr(dxr), _lib.ptr(dxc), _lib.ptr(dW), _lib.ptr(db),
                                        _lib.ptr(dA), _lib.ptr(dD), Bn, Di, N, R, len(ctx.stage_lens), lens, 0,
                                        _lib.stream_ptr())
        _lib.check(rc, "mlagg_msmm_scan_bwd")
        dt = ctx.in_dtypes
        return ((du[:, 0] + du[:, 2]).to(dt[0]), (du[:, 1] + du[:, 3]).to(dt[1]), dxr.to(dt[2]), dxc.to(dt[3]),
                dW.to(dt[4]), db.to(dt[5]), dA.to(dt[6]), dD.to(dt[7]), None, None)


def msmm_scan(xrow, xcol, xdbl_row, xdbl_col, Wdt, dt_bias, A, Ds, stage_lens):
    """4-direction multi-scale selective scan on un-permuted operands; see include/mlagg_b200.h (mlagg_msmm_scan_fwd).
    Returns out (B, 4, Di, L): direction k in row-major (k even) / column-major (k odd) order, mirroring undone."""
    return MSMMScanFn.apply(xrow, xcol, xdbl_row, xdbl_col, Wdt, dt_bias, A, Ds, tuple(stage_lens),
                            _wants_grad(xrow, xcol, xdbl_row, xdbl_col, Wdt, dt_bias, A, Ds))


# --------------------------------------------------------------------------------------------------------------------
# Tokens-major MSMM core: walk packing + fused scan + cross-merge behind ONE autograd node (C ABI: mlagg_walk_pack,
# mlagg_msmm_scan_fwd / _bwd, mlagg_walk_unpack).  Replaces SS2D_skip.forward_corev0 (MambaSkip.py:405-473) between the
# x_proj GEMM and out_norm, including everything its autograd graph does around the scan.
# --------------------------------------------------------------------------------------------------------------------
_DT = {torch.float32: 0, torch.bfloat16: 1}


def _tok(t):
    """tokens-major (B, L, C) operand for the walk kernels: fp32 / bf16, unit channel stride"""
    if t.dtype not in _DT:
        t = t.float()
    return t if t.stride(2) == 1 else t.contiguous()


def xdbl_pad(c35):
    """columns per walk of the tokens-major x_proj output: 2 directions x (dt_rank + 2 N), padded to a multiple of 4"""
    return (2 * c35 + 3) // 4 * 4


class MSMMTokensFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xc, xdbl, Wdt, dt_bias, A, Ds, hw, need_grad=True):
        if not xc.is_cuda:
            raise _lib.MlaggError("msmm_scan_tokens: CUDA tensors required (no CPU fallback in the product path)")
        ctx.in_dtypes = tuple(t.dtype for t in (xc, xdbl, Wdt, dt_bias, A, Ds))
        xc_, xd_ = _tok(xc.detach()), _tok(xdbl.detach())
        W_, b_, A_, D_ = map(_f32c, (Wdt, dt_bias, A, Ds))
        Bn, L, Di = xc_.shape
        N, R = A_.shape[1], W_.shape[1]
        C35 = R + 2 * N
        P = xdbl_pad(C35)
        hw = tuple((int(h), int(w)) for h, w in hw)
        assert sum(h * w for h, w in hw) == L and xd_.shape == (Bn, L, 2 * P)
        ns = len(hw)
        Hs, Ws = (ctypes.c_int * ns)(*[h for h, _ in hw]), (ctypes.c_int * ns)(*[w for _, w in hw])
        lens = (ctypes.c_int * ns)(*[h * w for h, w in hw])
        dev = xc.device
        L_ = _lib.lib()
        st = _lib.stream_ptr()
        xrow = torch.empty(Bn, Di, L, device=dev, dtype=torch.float32)
        xcol = torch.empty_like(xrow)
        xr = torch.empty(Bn, 2, C35, L, device=dev, dtype=torch.float32)
        xcl = torch.empty_like(xr)
        out = torch.empty(Bn, 4, Di, L, device=dev, dtype=torch.float32)
        ckpt = None
        if need_grad:
            ckpt = torch.empty(L_.mlagg_scan_ckpt_bytes(Bn, 4 * Di, L, N) // 4, device=dev, dtype=torch.float32)
        y = torch.empty(Bn, L, Di, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            with _lib.timed("walk_pack", 4):
                for src, c0, nc, dst, col in ((xc_, 0, Di, xrow, 0), (xc_, 0, Di, xcol, 1),
                                              (xd_, 0, 2 * C35, xr, 0), (xd_, P, 2 * C35, xcl, 1)):
                    rc = L_.mlagg_walk_pack(src.data_ptr(), _DT[src.dtype], src.stride(1), src.stride(0), c0, nc,
                                            dst.data_ptr(), nc * L, Bn, ns, Hs, Ws, col, st)
                    _lib.check(rc, "mlagg_walk_pack")
            with _lib.timed("scan_fwd"):
                rc = L_.mlagg_msmm_scan_fwd(_lib.ptr(xrow), _lib.ptr(xcol), _lib.ptr(xr), _lib.ptr(xcl), _lib.ptr(W_),
                                            _lib.ptr(b_), _lib.ptr(A_), _lib.ptr(D_), _lib.ptr(out), _lib.ptr(ckpt),
                                            Bn, Di, N, R, ns, lens, st)
            _lib.check(rc, "mlagg_msmm_scan_fwd")
            with _lib.timed("walk_unpack", 2):
                pl = Di * L * 4                                           # bytes per direction plane
                for k, col, acc in ((0, 0, 0), (1, 1, 1)):                 # y = out0 + out2 (+)= cols(out1 + out3)
                    rc = L_.mlagg_walk_unpack(out.data_ptr() + k * pl, out.data_ptr() + (k + 2) * pl, 4 * Di * L, Di, Di,
                                              y.data_ptr(), 0, Di, L * Di, 0, Bn, ns, Hs, Ws, col, acc, st)
                    _lib.check(rc, "mlagg_walk_unpack")
        ctx.hw = hw
        ctx.save_for_backward(xrow, xcol, xr, xcl, W_, b_, A_, D_, ckpt)
        return y

    @staticmethod
    def backward(ctx, dy):
        xrow, xcol, xr, xcl, W, b, A, D, ckpt = ctx.saved_tensors
        Bn, Di, L = xrow.shape
        N, R = A.shape[1], W.shape[1]
        C35 = R + 2 * N
        P = xdbl_pad(C35)
        hw = ctx.hw
        ns = len(hw)
        Hs, Ws = (ctypes.c_int * ns)(*[h for h, _ in hw]), (ctypes.c_int * ns)(*[w for _, w in hw])
        lens = (ctypes.c_int * ns)(*[h * w for h, w in hw])
        dev = xrow.device
        dt = ctx.in_dtypes
        dy = _tok(dy)
        dyw = torch.empty(Bn, 2, Di, L, device=dev, dtype=torch.float32)
        du = torch.empty(Bn, 4, Di, L, device=dev, dtype=torch.float32)
        dxd = torch.zeros(2, Bn, 2, C35, L, device=dev, dtype=torch.float32)    # [row walk | column walk], one memset
        dW, db, dA, dD = (_lib.zeros(t.shape, t.device) for t in (W, b, A, D))
        dxc = torch.empty(Bn, L, Di, device=dev, dtype=dt[0] if dt[0] in _DT else torch.float32)
        dxdbl = torch.empty(Bn, L, 2 * P, device=dev, dtype=dt[1] if dt[1] in _DT else torch.float32)
        L_ = _lib.lib()
        st = _lib.stream_ptr()
        pl = Di * L * 4
        with torch.cuda.device(dev):
            with _lib.timed("walk_pack", 2):
                for col in (0, 1):
                    rc = L_.mlagg_walk_pack(dy.data_ptr(), _DT[dy.dtype], dy.stride(1), dy.stride(0), 0, Di,
                                            dyw.data_ptr() + col * pl, 2 * Di * L, Bn, ns, Hs, Ws, col, st)
                    _lib.check(rc, "mlagg_walk_pack")
            with _lib.timed("scan_bwd"):
                rc = L_.mlagg_msmm_scan_bwd(_lib.ptr(xrow), _lib.ptr(xcol), _lib.ptr(xr), _lib.ptr(xcl), _lib.ptr(W),
                                            _lib.ptr(b), _lib.ptr(A), _lib.ptr(D), _lib.ptr(dyw), _lib.ptr(ckpt),
                                            _lib.ptr(du), _lib.ptr(dxd[0]), _lib.ptr(dxd[1]), _lib.ptr(dW), _lib.ptr(db),
                                            _lib.ptr(dA), _lib.ptr(dD), Bn, Di, N, R, ns, lens, 1, st)
            _lib.check(rc, "mlagg_msmm_scan_bwd")
            with _lib.timed("walk_unpack", 4):
                for k, col, acc in ((0, 0, 0), (1, 1, 1)):                 # dx = du0 + du2 (+)= cols(du1 + du3)
                    rc = L_.mlagg_walk_unpack(du.data_ptr() + k * pl, du.data_ptr() + (k + 2) * pl, 4 * Di * L, Di, Di,
                                              dxc.data_ptr(), _DT[dxc.dtype], Di, L * Di, 0, Bn, ns, Hs, Ws, col, acc, st)
                    _lib.check(rc, "mlagg_walk_unpack")
                for col in (0, 1):                                         # the two walks fill disjoint column blocks
                    rc = L_.mlagg_walk_unpack(dxd[col].data_ptr(), None, 2 * C35 * L, 2 * C35, P, dxdbl.data_ptr(),
                                              _DT[dxdbl.dtype], 2 * P, L * 2 * P, col * P, Bn, ns, Hs, Ws, col, 0, st)
                    _lib.check(rc, "mlagg_walk_unpack")
        return (dxc.to(dt[0]), dxdbl.to(dt[1]), dW.to(dt[2]), db.to(dt[3]), dA.to(dt[4]), dD.to(dt[5]), None, None)


def msmm_scan_tokens(xc, xdbl, Wdt, dt_bias, A, Ds, hw):
    """xc (B, L, Di): conv + SiLU output, tokens-major, stages concatenated; xdbl (B, L, 2 P), P = xdbl_pad(R + 2 N): the
    x_proj output with columns [direction 0 | direction 2 | pad | direction 1 | direction 3 | pad]; hw = [(H_s, W_s)].
    Returns the merged scan output y (B, L, Di) fp32 (reference MambaSkip.py:405-473)."""
    return MSMMTokensFn.apply(xc, xdbl, Wdt, dt_bias, A, Ds, tuple(hw), _wants_grad(xc, xdbl, Wdt, dt_bias, A, Ds))
