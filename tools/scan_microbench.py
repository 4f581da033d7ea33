"""Selective-scan microbench (BASELINE config 2 / the config-3 scan shape): CUDA-event timing with an L2 flush
between iterations; prints achieved algorithmic GB/s (5120 B/(b,l) fwd, 8704 B/(b,l) bwd at
D=384, G=4, N=16)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlagg_unet_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=10)
    ap.add_argument("--D", type=int, default=384)
    ap.add_argument("--G", type=int, default=4)
    ap.add_argument("--L", type=int, default=34000)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--bwd", type=int, default=1)
    a = ap.parse_args()
    dev = "cuda"
    N = 16
    torch.manual_seed(0)
    u = torch.randn(a.B, a.D, a.L, device=dev)
    dl = torch.randn(a.B, a.D, a.L, device=dev)
    A = -torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(a.D, 1).contiguous()
    Bm = torch.randn(a.B, a.G, N, a.L, device=dev)
    Cm = torch.randn(a.B, a.G, N, a.L, device=dev)
    Dk = torch.ones(a.D, device=dev)
    dt = torch.exp(torch.rand(a.D, device=dev) * 4.6 - 6.9)
    bias = dt + torch.log(-torch.expm1(-dt))
    out = torch.empty_like(u)
    L_ = _lib.lib()
    ckpt = torch.empty(L_.mlagg_scan_ckpt_bytes(a.B, a.D, a.L, N) // 4, device=dev)
    dout = torch.randn_like(u)
    du, dd = torch.empty_like(u), torch.empty_like(u)
    dA, dB, dC = torch.zeros_like(A), torch.zeros_like(Bm), torch.zeros_like(Cm)
    dD, db = torch.zeros_like(Dk), torch.zeros_like(bias)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    p = _lib.ptr

    def fwd(with_ckpt):
        _lib.check(L_.mlagg_selective_scan_fwd(p(u), p(dl), p(A), p(Bm), p(Cm), p(Dk), p(bias), p(out),
                                               p(ckpt) if with_ckpt else None, None, a.B, a.D, a.L, N, a.G, 1, st), "fwd")

    def bwd():
        _lib.check(L_.mlagg_selective_scan_bwd(p(u), p(dl), p(A), p(Bm), p(Cm), p(Dk), p(bias), p(dout), p(ckpt),
                                               p(du), p(dd), p(dA), p(dB), p(dC), p(dD), p(db),
                                               a.B, a.D, a.L, N, a.G, 1, st), "bwd")

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2], ts[0]

    fb = 4 * (3 * a.B * a.D * a.L + 2 * a.B * a.G * N * a.L)
    bb = 4 * (5 * a.B * a.D * a.L + 4 * a.B * a.G * N * a.L)
    res = {"shape": [a.B, a.D, a.L, a.G]}
    for name, fn, nbytes in (("fwd_infer", lambda: fwd(False), fb), ("fwd_train", lambda: fwd(True), fb),
                             ("bwd", bwd, bb)):
        if name == "bwd" and not a.bwd:
            continue
        med, best = timeit(fn)
        res[name] = {"ms_median": round(med, 4), "ms_best": round(best, 4), "alg_GBps": round(nbytes / med / 1e6, 1)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
