"""Row-streaming normalisation / pooling / loss kernels at the config-3 shapes: time, algorithmic bytes, GB/s (L2 flushed
between iterations; the HBM peak measured on this pool is 6.4 - 6.5 TB/s).  `one`: a few launches of each for ncu.

    python tools/norm_microbench.py [one]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mlagg_unet_b200 import ops  # noqa: E402
from mlagg_unet_b200.trainer import DeepSupervisionDiceCE  # noqa: E402

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


def cases():
    torch.manual_seed(0)
    out = []
    for Bn, H, W, C in ((10, 320, 320, 48), (10, 160, 160, 96), (10, 80, 80, 192)):
        x = torch.randn(Bn, C, H, W, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_()
        g = torch.randn_like(x)
        e = x.numel() * 2
        y = ops.instance_norm_cl(x, None, None, 1e-5, "leaky_relu", 0.01)
        out.append((f"instnorm fwd {Bn}x{H}x{W}x{C}", lambda x=x: ops.instance_norm_cl(x, None, None, 1e-5, "leaky_relu", 0.01), 3 * e))
        out.append((f"instnorm bwd {Bn}x{H}x{W}x{C}", lambda y=y, x=x, g=g: torch.autograd.grad(y, x, g, retain_graph=True), 5 * e))
    for M, C in ((256000, 96), (256000, 48), (64000, 192), (16000, 384)):
        x = torch.randn(10, M // 10, C, device="cuda").bfloat16().requires_grad_()
        ln = torch.nn.LayerNorm(C).cuda()
        g = torch.randn_like(x)
        e = x.numel() * 2
        y, short = ops.layer_norm_fork(x, ln, out_dtype=torch.bfloat16)
        out.append((f"layernorm fwd {M}x{C}", lambda x=x, ln=ln: ops.layer_norm_tokens(x, ln, out_dtype=torch.bfloat16), 2 * e))
        out.append((f"layernorm bwd+res {M}x{C}", lambda y=y, short=short, x=x, g=g: torch.autograd.grad([y, short], x, [g, g], retain_graph=True), 4 * e))
    x = torch.randn(10, 25600, 48, device="cuda").bfloat16().requires_grad_()
    y = ops.avgpool_tokens(x, 160, 160, 10, 10, gelu=True)
    g = torch.randn_like(y)
    out.append(("avgpool+gelu fwd 10x160x160x48", lambda: ops.avgpool_tokens(x, 160, 160, 10, 10, gelu=True), x.numel() * 2))
    out.append(("avgpool+gelu bwd 10x160x160x48", lambda: torch.autograd.grad(y, x, g, retain_graph=True), 2 * x.numel() * 2))
    logits = torch.randn(10, 320 * 320, 16, device="cuda").bfloat16()[..., :14].reshape(10, 320, 320, 14).permute(0, 3, 1, 2).requires_grad_()
    target = torch.randint(0, 14, (10, 1, 320, 320), device="cuda").float()
    lf = DeepSupervisionDiceCE(1)
    loss = lf.one(logits, target)
    e = 10 * 320 * 320 * 14 * 2
    out.append(("dice+CE stats fwd 10x14x320x320", lambda: lf.one(logits, target), e + target.numel() * 4))
    out.append(("dice+CE stats bwd 10x14x320x320", lambda: torch.autograd.grad(loss, logits, retain_graph=True), 2 * e + target.numel() * 4))
    return out


def main():
    cs = cases()
    if sys.argv[1:] == ["one"]:
        for _, fn, _ in cs:
            for _ in range(2):
                fn()
        torch.cuda.synchronize()
        return
    print(f"{'op':40s} {'us':>8s} {'GB/s':>7s}   (autograd wrapper and its small launches included)")
    for name, fn, nbytes in cs:
        t = t_us(fn)
        print(f"{name:40s} {t:8.1f} {nbytes / t / 1e3:7.0f}", flush=True)


if __name__ == "__main__":
    main()
