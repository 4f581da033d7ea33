"""How accurate is the gradient of the four lambda vectors (one scalar d lambda per attention module: a sum over every
token with heavy cancellation) at the SHIPPED stage-1 shape, fp32?  Product kernels and the oracle formulation in fp32,
both against the oracle formulation in fp64 on the same device.  TEST INFRASTRUCTURE (imports oracle/).

    python tools/lambda_grad_probe.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mlagg_unet_b200.mlagg import AggregatedAttention  # noqa: E402
from oracle import mlagg as o_mlagg  # noqa: E402


def rel(a, b):
    a, b = a.detach(), b.detach()
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def main():
    torch.manual_seed(0)
    Bn, H, W, dim, heads, sr = 10, 160, 160, 48, 1, 16
    for local in (True, False):
        m = AggregatedAttention(dim, (H, W), num_heads=heads, local=local, sr_ratio=sr).cuda().eval()
        with torch.no_grad():
            for n, p in m.named_parameters():
                if "lambda_" in n:
                    p.normal_(0, 0.1)
        x = torch.randn(Bn, H * W, dim, device="cuda")
        g = torch.randn(Bn, H * W, dim, device="cuda") * (1 + torch.linspace(0, 1, H * W, device="cuda")[None, :, None])
        names = [n for n, _ in m.named_parameters()]

        def run(fn, dt):
            ps = {n: p.detach().to(dt).requires_grad_() for n, p in m.named_parameters()}
            xx = x.to(dt).requires_grad_()
            y = fn(ps, xx)
            grads = torch.autograd.grad((y * g.to(dt)).sum(), [xx] + [ps[n] for n in names], allow_unused=True)
            return y.detach(), dict(zip(["x"] + names, grads))

        orc = lambda ps, xx: o_mlagg.aggregated_attention_forward(ps, xx, H, W, heads, local, sr)
        y64, g64 = run(orc, torch.float64)
        y32, g32 = run(orc, torch.float32)
        m.zero_grad(set_to_none=True)
        xx = x.detach().clone().requires_grad_()
        yo = m(xx, H, W)
        (yo * g).sum().backward()
        go = {"x": xx.grad, **{n: p.grad for n, p in m.named_parameters()}}
        print(f"== local={local}: forward ours {rel(yo, y64):.2e}  oracle-fp32 {rel(y32, y64):.2e}")
        for n in ["x"] + names:
            if g64[n] is None or go[n] is None:
                continue
            print(f"   d {n:14s} ours {rel(go[n], g64[n]):.2e}   oracle-fp32 {rel(g32[n], g64[n]):.2e}   |g| {float(g64[n].abs().max()):.3e}")


if __name__ == "__main__":
    main()
