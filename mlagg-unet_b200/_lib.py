"""ctypes binding of libmlagg_b200.so (C ABI: include/mlagg_b200.h).

The product path has NO fallback: if the shared library is missing or a CUDA tensor is not given,
the call raises.  `build()` compiles the library in-tree with nvcc for sm_100a."""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmlagg_b200.so")
_lib = None

c_p, c_i, c_f, c_sz, c_ll = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_longlong

# name -> (restype, argtypes); must list every symbol include/mlagg_b200.h declares
SIGNATURES = {
    "mlagg_version": (c_i, []),
    "mlagg_error_string": (ctypes.c_char_p, [c_i]),
    "mlagg_last_cuda_error": (ctypes.c_char_p, []),
    "mlagg_scan_ckpt_bytes": (c_sz, [c_i] * 4),
    "mlagg_selective_scan_fwd": (c_i, [c_p] * 10 + [c_i] * 6 + [c_p]),
    "mlagg_selective_scan_bwd": (c_i, [c_p] * 16 + [c_i] * 6 + [c_p]),
    "mlagg_msmm_scan_fwd": (c_i, [c_p] * 10 + [c_i] * 5 + [c_p, c_p]),
    "mlagg_msmm_scan_bwd": (c_i, [c_p] * 17 + [c_i] * 5 + [c_p, c_i, c_p]),
    "mlagg_residual_scale": (c_i, [c_p] * 4 + [c_ll, c_ll, c_i, c_p]),
    "mlagg_copy_rows": (c_i, [c_p, c_ll, c_ll, c_p, c_ll, c_ll, c_i, c_ll, c_i, c_i, c_p]),
    "mlagg_dice_ce_stats_fwd": (c_i, [c_p] * 4 + [c_i, c_ll, c_i, c_ll, c_ll, c_ll, c_i, c_i, c_p]),
    "mlagg_dice_ce_stats_bwd": (c_i, [c_p] * 5 + [c_i, c_ll, c_i, c_ll, c_ll, c_ll, c_i, c_i, c_p]),
    "mlagg_add_rows": (c_i, [c_p, c_ll, c_ll, c_p, c_ll, c_ll, c_i, c_ll, c_i, c_i, c_p]),
    "mlagg_bias_add_cl": (c_i, [c_p, c_p, c_ll, c_i, c_i, c_p]),
    "mlagg_silu_gate_fwd": (c_i, [c_p] * 3 + [c_ll, c_i, c_p]),
    "mlagg_silu_gate_bwd": (c_i, [c_p] * 5 + [c_ll, c_i, c_p]),
    "mlagg_diff_lambda_fwd": (c_i, [c_p] * 4 + [c_i, c_f, c_p, c_p]),
    "mlagg_diff_lambda_bwd": (c_i, [c_p] * 6 + [c_i, c_p, c_p]),
    "mlagg_walk_pack": (c_i, [c_p, c_i, c_ll, c_ll, c_i, c_i, c_p, c_ll, c_i, c_i, c_p, c_p, c_i, c_p]),
    "mlagg_walk_unpack": (c_i, [c_p, c_p, c_ll, c_i, c_i, c_p, c_i, c_ll, c_ll, c_i, c_i, c_i, c_p, c_p, c_i, c_i, c_p]),
    "mlagg_layernorm_fwd": (c_i, [c_p] * 6 + [c_ll, c_i, c_f, c_i, c_i, c_p]),
    "mlagg_layernorm_bwd": (c_i, [c_p] * 8 + [c_ll, c_i, c_i, c_i, c_p]),
    "mlagg_layernorm_bwd_res": (c_i, [c_p] * 9 + [c_ll, c_i, c_i, c_i, c_p]),
    "mlagg_linattn_state_bytes": (c_sz, [c_i] * 3),
    "mlagg_linattn_fwd": (c_i, [c_p] * 6 + [c_i] * 5 + [c_ll] * 4 + [c_f, c_i, c_p]),
    "mlagg_linattn_bwd": (c_i, [c_p] * 10 + [c_i] * 5 + [c_ll] * 7 + [c_f, c_i, c_p]),
    "mlagg_instnorm_fwd": (c_i, [c_p] * 5 + [c_i] * 3 + [c_f, c_i, c_f, c_i, c_p]),
    "mlagg_instnorm_bwd": (c_i, [c_p] * 9 + [c_i] * 3 + [c_i, c_f, c_i, c_p]),
    "mlagg_instnorm_res_fwd": (c_i, [c_p] * 6 + [c_i] * 3 + [c_f, c_i, c_f, c_i, c_p]),
    "mlagg_instnorm_res_bwd": (c_i, [c_p] * 11 + [c_i] * 3 + [c_i, c_f, c_i, c_p]),
    "mlagg_avgpool_tokens_fwd": (c_i, [c_p] * 2 + [c_i] * 8 + [c_p]),
    "mlagg_avgpool_tokens_bwd": (c_i, [c_p] * 3 + [c_i] * 8 + [c_p]),
    "mlagg_colsum": (c_i, [c_p, c_p, c_ll, c_i, c_ll, c_i, c_p]),
    "mlagg_linear_fwd": (c_i, [c_p, c_ll, c_p, c_ll, c_p, c_p, c_ll, c_p, c_ll, c_ll, c_i, c_i, c_i, c_i, c_p]),
    "mlagg_linear_bwd_data": (c_i, [c_p, c_ll, c_p, c_ll, c_p, c_ll, c_i, c_p, c_ll, c_ll, c_i, c_i, c_i, c_p]),
    "mlagg_linear_bwd_weight": (c_i, [c_p, c_ll, c_p, c_ll, c_p, c_ll, c_p, c_ll, c_i, c_i, c_p]),
    "mlagg_dwconv3x3_fwd": (c_i, [c_p] * 4 + [c_i] * 6 + [c_p]),
    "mlagg_dwconv3x3_bwd": (c_i, [c_p] * 8 + [c_i] * 6 + [c_p]),
    "mlagg_dwconv3x3_fwd_strided": (c_i, [c_p] * 5 + [c_i] * 4 + [c_ll] * 6 + [c_i, c_i, c_i, c_p]),
    "mlagg_dwconv3x3_bwd_strided": (c_i, [c_p] * 8 + [c_i] * 4 + [c_ll] * 6 + [c_p, c_p, c_ll, c_ll] + [c_i, c_i, c_p]),
    "mlagg_causal_conv1d_fwd": (c_i, [c_p] * 4 + [c_i] * 5 + [c_p]),
    "mlagg_causal_conv1d_bwd": (c_i, [c_p] * 7 + [c_i] * 5 + [c_p]),
    "mlagg_pooled_diffattn_ws_bytes": (c_sz, [c_i] * 4),
    "mlagg_pooled_diffattn_saved_bytes": (c_sz, [c_i] * 4),
    "mlagg_pooled_diffattn_fwd": (c_i, [c_p] * 6 + [c_i] * 5 + [c_ll] * 3 + [c_f, c_p, c_f, c_f] + [c_i, c_p]),
    "mlagg_pooled_diffattn_bwd": (c_i, [c_p] * 12 + [c_i] * 5 + [c_ll] * 5 + [c_f, c_p, c_f, c_f] + [c_i, c_p]),
    "mlagg_local_diffattn_ws_bytes": (c_sz, [c_i] * 5),
    "mlagg_local_diffattn_fwd": (c_i, [c_p] * 5 + [c_i] * 5 + [c_ll] * 3 + [c_f, c_p, c_f, c_f] + [c_i, c_p]),
    "mlagg_local_diffattn_bwd": (c_i, [c_p] * 11 + [c_i] * 5 + [c_ll] * 5 + [c_f, c_p, c_f, c_f] + [c_i, c_p]),
}


class MlaggError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> mlagg-unet_b200/libmlagg_b200.so"""
    r = subprocess.run(["make", f"-j{os.cpu_count() or 4}", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise MlaggError("building libmlagg_b200.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MlaggError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU / PyTorch fallback for the hot path)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        L = lib()
        raise MlaggError(f"{what}: {L.mlagg_error_string(rc).decode()} (code {rc}) "
                         f"{L.mlagg_last_cuda_error().decode()}")


# ---- bookkeeping for bench.py: how many of OUR kernels were launched, and (optionally) per-kernel CUDA-event timing
STATS = {"launches": 0, "events": None, "bytes": {}}


class timed:
    """with timed("scan_bwd", n_kernels): ...  -- counts launches; records CUDA events on the current stream when
    STATS["events"] is a dict (bench.py turns that on for the timed region)."""

    def __init__(self, name, n_kernels=1, nbytes=0):
        self.name, self.n, self.nbytes = name, n_kernels, nbytes      # nbytes: ALGORITHMIC HBM bytes of the call (bench roofline)

    def __enter__(self):
        STATS["launches"] += self.n
        if STATS["events"] is not None:
            import torch
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if STATS["events"] is not None:
            self.e1.record()
            STATS["events"].setdefault(self.name, []).append((self.e0, self.e1))
            STATS["bytes"][self.name] = STATS["bytes"].get(self.name, 0) + self.nbytes
        return False


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


# ---- side stream for work nothing downstream waits for.  The weight-gradient GEMMs feed only the optimizer, are mostly
# latency-bound (small grids, split-K atomics, 10 - 20 us each, 120 per step) and read the same dy as the data-gradient GEMM
# that follows them: they run on a second stream next to the main backward chain and are joined before the gradients
# are gathered.  Operands are kept referenced until the join so that the caching allocator cannot hand their memory to a
# later main-stream kernel while the side kernel still reads it.  Only active inside a trainer step (side_begin()).
# A tensor produced on the side stream must not travel through autograd (AccumulateGrad may clone it, a view's backward
# may copy it -- kernels on the main stream that the engine would not order behind the side stream): the autograd
# wrappers hand such gradients over with stash_grad(parameter, tensor), return None to autograd, and the trainer attaches
# them to .grad after side_join().
_SIDE = {"stream": None, "on": False, "keep": [], "pending": 0, "grads": []}


def side_begin(device):
    import os
    import torch
    if os.environ.get("MLAGG_SIDE_STREAM", "1") == "0":
        _SIDE["on"] = False
        return
    if _SIDE["stream"] is None or _SIDE["stream"].device != torch.device(device):
        _SIDE["stream"] = torch.cuda.Stream(device=device)
    _SIDE["on"] = True


class side_launch:
    """with side_launch(t1, t2, ...) as on_side: ...   -- the body runs on the side stream behind everything already queued
    on the current stream; the tensors stay referenced until side_join().  A no-op outside a trainer step."""

    def __init__(self, *tensors):
        self.tensors = tensors

    def __enter__(self):
        import torch
        self.active = _SIDE["on"]
        if not self.active:
            return False
        if _SIDE["pending"] >= 24:                 # bound the memory held for pending launches
            side_join(final=False)
        st = _SIDE["stream"]
        ev = torch.cuda.Event()
        ev.record()
        st.wait_event(ev)
        _SIDE["keep"].append(self.tensors)
        _SIDE["pending"] += 1
        self.ctx = torch.cuda.stream(st)
        self.ctx.__enter__()
        return True

    def __exit__(self, *exc):
        if self.active:
            self.ctx.__exit__(*exc)
        return False


def side_active():
    return _SIDE["on"]


def leaf_param(t):
    """the leaf Parameter a gradient for `t` ends up in, if `t` is one or a same-size dense view of one; else None"""
    import torch
    if t is None:
        return None
    base = t if t._base is None else t._base
    if not (isinstance(base, torch.nn.Parameter) and base.is_leaf and base.requires_grad and base.numel() == t.numel()):
        return None
    if t is not base and not (t.is_contiguous() and (base.is_contiguous() or (base.dim() == 4 and base.shape[2] == base.shape[3] == 1))):
        return None
    return base


def stash_grad(param, g):
    _SIDE["grads"].append((param, g.view(param.shape) if g.shape != param.shape else g))


def take_stashed_grads():
    out, _SIDE["grads"] = _SIDE["grads"], []
    return out


_BRANCH = {"stream": None}


def branch_stream(device):
    """a (default = lowest priority) stream for an independent sub-graph of the forward pass; autograd replays the
    sub-graph's backward nodes on the same stream.  Only with MLAGG_BRANCH_STREAM=1 (experimental)."""
    import os
    import torch
    if os.environ.get("MLAGG_SIDE_STREAM", "1") == "0" or os.environ.get("MLAGG_BRANCH_STREAM", "0") != "1":
        return None
    if _BRANCH["stream"] is None or _BRANCH["stream"].device != torch.device(device):
        _BRANCH["stream"] = torch.cuda.Stream(device=device)
    return _BRANCH["stream"]


def side_join(final=True):
    """the current stream waits for everything launched through side_launch"""
    import torch
    if _SIDE["pending"]:
        torch.cuda.current_stream().wait_stream(_SIDE["stream"])
        _SIDE["keep"].clear()
        _SIDE["pending"] = 0
    if final:
        _SIDE["on"] = False


# ---- per-step arena of zero-initialised fp32 scratch.  Every backward wrapper needs small zero-filled accumulators (bias /
# weight gradients, column sums, scalar sums): ~350 separate fill kernels per train step.  While a trainer step is running
# they are carved out of one buffer that is cleared by a single memset at the start of the step; outside a step (tests,
# direct module use) `zeros` is plain torch.zeros.
_ARENA = {"buf": None, "off": 0, "active": False}
_ARENA_FLOATS, _ARENA_MAX_REQ = 4 << 20, 1 << 16


def arena_begin(device):
    import torch
    if _ARENA["buf"] is None or _ARENA["buf"].device != torch.device(device):
        _ARENA["buf"] = torch.empty(_ARENA_FLOATS, device=device, dtype=torch.float32)
    _ARENA["buf"].zero_()
    _ARENA["off"], _ARENA["active"] = 0, True


def arena_end():
    _ARENA["active"] = False


def zeros(shape, device):
    """fp32 zeros of `shape`: a 256-byte aligned slice of the step arena when one is active, else torch.zeros."""
    import math
    import torch
    shape = (shape,) if isinstance(shape, int) else tuple(shape)
    n = math.prod(shape)
    if _ARENA["active"] and n <= _ARENA_MAX_REQ and _ARENA["buf"].device == torch.device(device):
        off = _ARENA["off"]
        if off + n <= _ARENA_FLOATS:
            _ARENA["off"] = off + (n + 63) // 64 * 64
            return _ARENA["buf"][off:off + n].view(shape)
    return torch.zeros(shape, device=device, dtype=torch.float32)
