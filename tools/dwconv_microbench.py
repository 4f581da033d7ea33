"""Depthwise 3x3 convolution family at the config-3 stage shapes (tokens-major bf16, B = 10): forward and backward, ring
kernels (shared-memory row rings fed by cp.async) against the strip kernels they replace (MLAGG_DWCONV_STRIP=1), L2
flushed between iterations.  GB/s = algorithmic bytes (forward: x + y [+ residual]; backward: x, dy read, dz written and
read, dx written [+ v, dv]) / time; the HBM peak measured on this pool is 6.4 - 6.5 TB/s.

    python tools/dwconv_microbench.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mlagg_unet_b200.ops import conv_glu_core, dwconv3x3_tokens  # noqa: E402

flush = None


def t_us(fn, n=10, warm=3):
    global flush
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        flush.max()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def one():
    """python tools/dwconv_microbench.py one  -- a few launches of each ring kernel at 10 x 160 x 160 x 96, for ncu"""
    torch.manual_seed(0)
    H, W, C = 160, 160, 96
    x = torch.randn(10, H * W, C, device="cuda").bfloat16().requires_grad_()
    w = (0.3 * torch.randn(C, 1, 3, 3, device="cuda")).requires_grad_()
    b = (0.1 * torch.randn(C, device="cuda")).requires_grad_()
    g = torch.randn(10, H * W, C, device="cuda").bfloat16()
    for _ in range(3):
        y = dwconv3x3_tokens(x, w, b, H, W, silu=True)
        torch.autograd.grad(y, [x, w, b], g)
    torch.cuda.synchronize()


def main():
    if sys.argv[1:] == ["one"]:
        return one()
    torch.manual_seed(0)
    Bn = int(os.environ.get("BATCH", 10))
    shapes = [(160, 160, 48), (160, 160, 96), (80, 80, 96), (80, 80, 192), (40, 40, 384), (20, 20, 768)]
    print(f"{'shape':>22s} {'op':>10s} {'ring us':>9s} {'GB/s':>7s} {'strip us':>9s} {'GB/s':>7s}")
    for H, W, C in shapes:
        x = torch.randn(Bn, H * W, C, device="cuda").bfloat16().requires_grad_()
        h2 = torch.randn(Bn, H * W, 2 * C, device="cuda").bfloat16().requires_grad_()
        w = (0.3 * torch.randn(C, 1, 3, 3, device="cuda")).requires_grad_()
        b = (0.1 * torch.randn(C, device="cuda")).requires_grad_()
        g = torch.randn(Bn, H * W, C, device="cuda").bfloat16()
        e = Bn * H * W * C * 2
        y = dwconv3x3_tokens(x, w, b, H, W, silu=True)
        yg = conv_glu_core(h2, w, b, H, W, silu=True)
        cases = [("fwd silu", lambda: dwconv3x3_tokens(x, w, b, H, W, silu=True), 2 * e),
                 ("bwd silu", lambda: torch.autograd.grad(y, [x, w, b], g, retain_graph=True), 6 * e),
                 ("glu fwd", lambda: conv_glu_core(h2, w, b, H, W, silu=True), 3 * e),
                 ("glu bwd", lambda: torch.autograd.grad(yg, [h2, w, b], g, retain_graph=True), 8 * e)]
        for name, fn, nbytes in cases:
            os.environ.pop("MLAGG_DWCONV_STRIP", None)
            t1 = t_us(fn)
            os.environ["MLAGG_DWCONV_STRIP"] = "1"
            t2 = t_us(fn)
            os.environ.pop("MLAGG_DWCONV_STRIP", None)
            print(f"{str((Bn, H, W, C)):>22s} {name:>10s} {t1:9.1f} {nbytes / t1 / 1e3:7.0f} {t2:9.1f} {nbytes / t2 / 1e3:7.0f}", flush=True)


if __name__ == "__main__":
    main()
