"""Projection GEMMs of one train step at config 3 (B = 10, 320 x 320): the tcgen05 / TMEM / TMA kernels (mlagg_linear_*)
beside cuBLAS (torch.mm / addmm, what round 1 shipped) on the same bf16 operands.  L2 is flushed between launches by
READING a 256 MB buffer (a write flush leaves ~126 MB of dirty lines whose write-back is then charged to the timed kernel).
Per shape: forward (bias epilogue), data gradient, weight gradient; microseconds and algorithmic GB/s
(operands + result once) against the measured HBM peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlagg_unet_b200 import gemm  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t_us(fn, n=20):
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(n):
        flush.max()          # read flush: L2 ends up full of clean lines
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n * 1e3


# (name, tokens, N_out, K_in, calls per step)
B = 10
SHAPES = []
for s, (n, C) in enumerate([(25600, 96), (6400, 192), (1600, 384), (400, 768)]):
    M = B * n
    SHAPES += [(f"s{s} in/act/out_proj", M, C, C, 6), (f"s{s} fc1", M, 2 * C, C, 2), (f"s{s} fc2", M, C, 2 * C, 2),
               (f"s{s} q", M, C // 2, C // 2, 4), (f"s{s} kv", M, C, C // 2, 4)]
SHAPES += [("msmm in_proj", B * 34000, 96, 48, 1), ("msmm x_proj", B * 34000, 144, 96, 1), ("msmm out_proj", B * 34000, 48, 96, 1),
           ("msmm glu fc1 s0", B * 25600, 256, 48, 1), ("msmm glu fc2 s0", B * 25600, 48, 128, 1)]

print(f"{'shape':22s} {'M':>7s} {'N':>5s} {'K':>5s} | {'fwd tc':>8s} {'cublas':>8s} {'GB/s':>6s} | {'dX tc':>8s} {'cublas':>8s} | {'dW tc':>8s} {'cublas':>8s}")
tot = {"tc": 0.0, "cb": 0.0}
for name, M, N, K, calls in SHAPES:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda")
    b16 = b.bfloat16()
    dy = torch.randn(M, N, device="cuda").bfloat16()
    f_tc = t_us(lambda: gemm.linear_fwd(x, w, b))
    f_cb = t_us(lambda: torch.addmm(b16, x, w.t()))
    d_tc = t_us(lambda: gemm.linear_bwd_data(dy, w))
    d_cb = t_us(lambda: torch.mm(dy, w))
    w_tc = t_us(lambda: gemm.linear_bwd_weight(dy, x))
    w_cb = t_us(lambda: torch.mm(dy.t(), x).float())
    gbs = 2 * (M * K + M * N + N * K) / (f_tc * 1e-6) / 1e9
    print(f"{name:22s} {M:7d} {N:5d} {K:5d} | {f_tc:8.1f} {f_cb:8.1f} {gbs:6.0f} | {d_tc:8.1f} {d_cb:8.1f} | {w_tc:8.1f} {w_cb:8.1f}")
    tot["tc"] += calls * (f_tc + d_tc + w_tc)
    tot["cb"] += calls * (f_cb + d_cb + w_cb)
print(f"per train step (calls weighted): tcgen05 {tot['tc'] / 1e3:.2f} ms   cuBLAS {tot['cb'] / 1e3:.2f} ms   (HBM peak {PEAK:.0f} GB/s)")
