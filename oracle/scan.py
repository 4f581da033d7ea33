"""oracle/scan.py -- TEST INFRASTRUCTURE ONLY (parity oracle / CPU baseline).

CPU restatement of the selective scan that `SS2D_skip.forward_corev0` calls
(reference: mlagg/nnunetv2/training/nnUNetTrainer/variants/mamba/MambaSkip.py:445-451).
The arithmetic lives in the un-vendored, un-pinned `mamba-ssm` package
(`selective_scan_fn` / `selective_scan_ref`; reference README.md:49-50); what is
restated here is that package's published S6 recurrence and
its analytic gradient (App. A.2).  PARITY UNPINNED against mamba-ssm itself (the package is absent everywhere).
What it IS pinned to: (i) Hugging Face transformers' `MambaMixer.slow_forward` -- an independent torch
restatement of the same definition that ships in the image -- forward and every gradient, 2e-5
(tests/test_oracle_pin_hf.py); (ii) through the CUDA path, vLLM's `selective_scan_fwd`, mamba-ssm's forward
kernel carried into vLLM (tests/test_scan_gpu.py::test_forward_matches_mamba_ssm_derived_cuda_kernel).

Three layers, each checking the next:
  * `selective_scan_loop`   -- plain torch, one python iteration per time step (small L only);
                               differentiable by autograd, so it also pins the analytic backward.
  * `scan_fwd_c/scan_bwd_c` -- oracle/scan_ref.c through ctypes (fp32 or fp64 arithmetic, OpenMP).
  * `selective_scan_oracle` -- autograd.Function over the C code, same call signature as
                               mamba_ssm.ops.selective_scan_interface.selective_scan_fn.

Nothing under mlagg-unet_b200/ imports this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build_c_oracle(force: bool = False) -> str:
    """Compile oracle/scan_ref.c (gcc, OpenMP) -> oracle/libscan_ref.so."""
    so = os.path.join(_HERE, "libscan_ref.so")
    src = os.path.join(_HERE, "scan_ref.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "clean"], check=True)
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build_c_oracle())
    return _LIB


def c_threads() -> int:
    return int(_lib().scan_ref_f32_threads())


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _prep(t):
    return None if t is None else t.detach().to("cpu", torch.float32).contiguous()


def scan_fwd_c(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, fp64=False,
               return_last_state=False):
    """(Bn,D,L) x2, (D,N), (Bn,G,N,L) x2 -> out (Bn,D,L) fp32 [, last_state (Bn,D,N)]."""
    u, delta, A, B, C, D, delta_bias = map(_prep, (u, delta, A, B, C, D, delta_bias))
    if B.dim() == 3:
        B = B.unsqueeze(1)
    if C.dim() == 3:
        C = C.unsqueeze(1)
    Bn, Dm, L = u.shape
    N, G = A.shape[1], B.shape[1]
    out = torch.empty_like(u)
    last = torch.empty(Bn, Dm, N) if return_last_state else None
    fn = _lib().scan_ref_f64_fwd if fp64 else _lib().scan_ref_f32_fwd
    fn(_p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(delta_bias), int(bool(delta_softplus)),
       Bn, Dm, L, N, G, _p(out), _p(last))
    return (out, last) if return_last_state else out


def scan_bwd_c(u, delta, A, B, C, D, delta_bias, delta_softplus, dout, fp64=False):
    """Analytic backward (App. A.2) -> (du, ddelta, dA, dB, dC, dD, ddelta_bias)."""
    u, delta, A, B, C, D, delta_bias, dout = map(_prep, (u, delta, A, B, C, D, delta_bias, dout))
    squeeze_b = B.dim() == 3
    squeeze_c = C.dim() == 3
    if squeeze_b:
        B = B.unsqueeze(1)
    if squeeze_c:
        C = C.unsqueeze(1)
    Bn, Dm, L = u.shape
    N, G = A.shape[1], B.shape[1]
    du, dd = torch.empty_like(u), torch.empty_like(u)
    dA, dB, dC = torch.empty_like(A), torch.empty_like(B), torch.empty_like(C)
    dD = torch.empty(Dm) if D is not None else None
    db = torch.empty(Dm) if delta_bias is not None else None
    fn = _lib().scan_ref_f64_bwd if fp64 else _lib().scan_ref_f32_bwd
    fn(_p(u), _p(delta), _p(A), _p(B), _p(C), _p(D), _p(delta_bias), int(bool(delta_softplus)), _p(dout),
       Bn, Dm, L, N, G, _p(du), _p(dd), _p(dA), _p(dB), _p(dC), _p(dD), _p(db))
    if squeeze_b:
        dB = dB.squeeze(1)
    if squeeze_c:
        dC = dC.squeeze(1)
    return du, dd, dA, dB, dC, dD, db


def selective_scan_loop(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                        return_last_state=False):
    """Per-time-step torch restatement of App. A.1 (any dtype; differentiable).

    u, delta (Bn,D,L); A (D,N); B, C (Bn,N,L) or (Bn,G,N,L); D, delta_bias (D,); z (Bn,D,L) or None.
    """
    dtype_in = u.dtype
    u, delta = u.to(A.dtype if A.dtype == torch.float64 else torch.float32), delta.to(
        A.dtype if A.dtype == torch.float64 else torch.float32)
    if delta_bias is not None:
        delta = delta + delta_bias.to(delta.dtype)[..., None]
    if delta_softplus:
        delta = F.softplus(delta)
    Bn, Dm, L = u.shape
    N = A.shape[1]
    if B.dim() == 3:
        B = B.unsqueeze(1)
    if C.dim() == 3:
        C = C.unsqueeze(1)
    G = B.shape[1]
    rep = Dm // G
    Bx = B.to(u.dtype).repeat_interleave(rep, dim=1)  # (Bn, D, N, L)
    Cx = C.to(u.dtype).repeat_interleave(rep, dim=1)
    h = u.new_zeros(Bn, Dm, N)
    ys = []
    for t in range(L):
        a = torch.exp(delta[:, :, t, None] * A)
        h = a * h + delta[:, :, t, None] * Bx[:, :, :, t] * u[:, :, t, None]
        ys.append((h * Cx[:, :, :, t]).sum(-1))
    y = torch.stack(ys, dim=2)
    out = y if D is None else y + u * D.to(u.dtype)[..., None]
    if z is not None:
        out = out * F.silu(z.to(out.dtype))
    out = out.to(dtype_in)
    return (out, h) if return_last_state else out


class _OracleScan(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, delta, A, B, C, D, delta_bias, delta_softplus, fp64):
        ctx.save_for_backward(u, delta, A, B, C, D, delta_bias)
        ctx.flags = (delta_softplus, fp64)
        return scan_fwd_c(u, delta, A, B, C, D, delta_bias, delta_softplus, fp64).to(u.device, u.dtype)

    @staticmethod
    def backward(ctx, dout):
        u, delta, A, B, C, D, delta_bias = ctx.saved_tensors
        sp, fp64 = ctx.flags
        du, dd, dA, dB, dC, dD, db = scan_bwd_c(u, delta, A, B, C, D, delta_bias, sp, dout, fp64)
        cast = lambda g, ref: None if g is None else g.to(ref.dtype)
        return (cast(du, u), cast(dd, delta), cast(dA, A), cast(dB, B), cast(dC, C),
                cast(dD, D) if D is not None else None,
                cast(db, delta_bias) if delta_bias is not None else None, None, None)


def selective_scan_oracle(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                          return_last_state=False, fp64=False):
    """Same signature as mamba_ssm's selective_scan_fn; CPU, C code, analytic backward."""
    out = _OracleScan.apply(u, delta, A, B, C, D, delta_bias, delta_softplus, fp64)
    if z is not None:
        out = out * F.silu(z.to(out.dtype))
    if return_last_state:
        _, last = scan_fwd_c(u, delta, A, B, C, D, delta_bias, delta_softplus, fp64, True)
        return out, last
    return out


def selective_scan_chunked(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=False, chunks=4):
    """Chunk-parallel form of the same recurrence ("next" item 1) -- a MATH PROTOTYPE for the round-2 kernels,
    test infrastructure like everything else here.  L is cut into `chunks` pieces:
      pass A (parallel over chunks)   local scan of every chunk from a zero state -> y_local, h_end_local, and the running
                                      sums S_t of delta inside the chunk;
      pass B (sequential, K steps)    h_start[k+1] = exp(A * S_total[k]) * h_start[k] + h_end_local[k];
      pass C (parallel over all t)    y_t += sum_n C_t,n * exp(A_n * S_t) * h_start[k],n      (no recurrence).
    Returns exactly what `selective_scan_loop` returns (any dtype; differentiable through autograd)."""
    dt = torch.float64 if A.dtype == torch.float64 else torch.float32
    u_, dl = u.to(dt), delta.to(dt)
    if delta_bias is not None:
        dl = dl + delta_bias.to(dt)[..., None]
    if delta_softplus:
        dl = F.softplus(dl)
    Bn, Dm, L = u_.shape
    if B.dim() == 3:
        B = B.unsqueeze(1)
    if C.dim() == 3:
        C = C.unsqueeze(1)
    rep = Dm // B.shape[1]
    Bx = B.to(dt).repeat_interleave(rep, dim=1)                    # (Bn, D, N, L)
    Cx = C.to(dt).repeat_interleave(rep, dim=1)
    bounds = [round(i * L / chunks) for i in range(chunks + 1)]
    h_start = u_.new_zeros(Bn, Dm, A.shape[1])
    ys = []
    for k in range(chunks):
        t0, t1 = bounds[k], bounds[k + 1]
        if t1 == t0:
            continue
        # pass A: local scan from zero (delta already activated: no bias / softplus again)
        y_loc, h_end = selective_scan_loop(u_[:, :, t0:t1], dl[:, :, t0:t1], A.to(dt), Bx[..., t0:t1], Cx[..., t0:t1],
                                           None, None, None, False, return_last_state=True)
        S = torch.cumsum(dl[:, :, t0:t1], dim=-1)                  # inclusive running sum of delta inside the chunk
        # pass C: carried-state term, one exponential per (channel, state, step), no dependence between steps
        P = torch.exp(A.to(dt)[None, :, :, None] * S[:, :, None, :])            # (Bn, D, N, Lc)
        ys.append(y_loc + (Cx[..., t0:t1] * P * h_start[..., None]).sum(2))
        # pass B: chain the chunk states
        h_start = P[..., -1] * h_start + h_end
    y = torch.cat(ys, dim=-1)
    if D is not None:
        y = y + u_ * D.to(dt)[..., None]
    return y.to(u.dtype)
