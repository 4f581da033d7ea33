"""Projection GEMMs of the hot path on the sm_100a tensor cores (C ABI: mlagg_linear_fwd / _bwd_data / _bwd_weight,
csrc/gemm_tc.cu: TMA + tcgen05.mma + TMEM).  Thin tensor-level wrappers; the autograd Functions that use them live in
ops.py.  Reference sites: nnUNetTrainer_MLAgg_2D_dt_MS.py:673-674, :849-850, :867, :902, :176-192; MambaSkip.py:301, :345,
:431, :559-577.

bf16 only: the fp32 parity path keeps cuBLAS (a library GEMM in full fp32 -- the tensor cores have no fp32 mode)."""
from __future__ import annotations

import torch

from . import _lib

ACT = {None: 0, "none": 0, "gelu": 1, "silu": 2}
_DT = {torch.float32: 0, torch.bfloat16: 1}


def rows_ok(t):
    """2-D bf16 CUDA operand the kernels take in place: unit column stride, row stride % 8 == 0, 16-byte aligned base"""
    return (t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 8 == 0
            and t.stride(0) >= t.shape[1] and t.data_ptr() % 16 == 0)


def supported(x2, w):
    """x2 (M, K), w (N, K)"""
    return rows_ok(x2) and rows_ok(w) and w.shape[0] % 8 == 0 and w.shape[1] % 8 == 0 and x2.shape[0] > 0


def linear_fwd(x2, w, bias=None, act=None, out_dtype=torch.bfloat16, want_pre=False, out=None):
    """y = act(x2 @ w.T + bias); returns (y, pre) with pre = the bf16 pre-activation when want_pre"""
    M, K = x2.shape
    N = w.shape[0]
    y = torch.empty(M, N, device=x2.device, dtype=out_dtype) if out is None else out
    pre = torch.empty(M, N, device=x2.device, dtype=torch.bfloat16) if want_pre else None
    b = None if bias is None else bias.detach().float().contiguous()
    with torch.cuda.device(x2.device), _lib.timed("linear_fwd", 1, 2 * (M * K + N * K + M * N * (2 if pre is not None else 1))):
        rc = _lib.lib().mlagg_linear_fwd(x2.data_ptr(), x2.stride(0), w.data_ptr(), w.stride(0), _lib.ptr(b), y.data_ptr(),
                                         y.stride(0), _lib.ptr(pre), 0 if pre is None else pre.stride(0), M, N, K, ACT[act],
                                         _DT[y.dtype], _lib.stream_ptr())
    _lib.check(rc, "mlagg_linear_fwd")
    return y, pre


def linear_bwd_data(dy2, w, aux=None, act=None, out_dtype=torch.bfloat16):
    """dx = (dy2 @ w) * act'(aux); dy2 (M, N), w (N, K), aux (M, K) or None"""
    M, N = dy2.shape
    K = w.shape[1]
    dx = torch.empty(M, K, device=dy2.device, dtype=out_dtype)
    with torch.cuda.device(dy2.device), _lib.timed("linear_bwd_data", 1, 2 * (M * N + N * K + M * K * (2 if aux is not None else 1))):
        rc = _lib.lib().mlagg_linear_bwd_data(dy2.data_ptr(), dy2.stride(0), w.data_ptr(), w.stride(0), _lib.ptr(aux),
                                              0 if aux is None else aux.stride(0), ACT[act], dx.data_ptr(), dx.stride(0),
                                              M, N, K, _DT[out_dtype], _lib.stream_ptr())
    _lib.check(rc, "mlagg_linear_bwd_data")
    return dx


class _Here:
    def __enter__(self):
        return False

    def __exit__(self, *exc):
        return False


def linear_bwd_weight(dy2, x2, want_db=False, side=False):
    """dw (N, K) fp32 = dy2.T @ x2 [, db (N,) fp32 = dy2.sum(0) from the same pass]; dy2 (M, N), x2 (M, K).
    side=True (only with _lib.side_active(), and only when the results go to _lib.stash_grad, never through autograd):
    the launch goes to the side stream, next to the main backward chain."""
    M, N = dy2.shape
    K = x2.shape[1]
    with (_lib.side_launch(dy2, x2) if side else _Here()):
        dw = _lib.zeros((N, K), dy2.device)
        db = _lib.zeros((N,), dy2.device) if want_db else None
        with torch.cuda.device(dy2.device), _lib.timed("linear_bwd_weight", 1, 2 * M * (N + K) + 4 * N * K):
            rc = _lib.lib().mlagg_linear_bwd_weight(dy2.data_ptr(), dy2.stride(0), x2.data_ptr(), x2.stride(0), dw.data_ptr(),
                                                    dw.stride(0), _lib.ptr(db), M, N, K, _lib.stream_ptr())
    _lib.check(rc, "mlagg_linear_bwd_weight")
    return (dw, db) if want_db else dw
