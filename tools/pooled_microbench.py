"""Pooled differential attention microbench at the stage shapes of config 3 (B=10, hd=24, P=100), bf16."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlagg_unet_b200 import attention as att

def t_ms(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

stages = [(25600, 48), (6400, 96), (1600, 192), (400, 384)] if len(sys.argv) < 2 else [(25600, 48)]
for N, C in stages:
    h, hd, P, Bn = C // 48, 24, 100, 10
    q = torch.randn(Bn, N, C, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    kv = torch.randn(Bn, P, 2 * C, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    lam = torch.tensor(0.8, device="cuda", requires_grad=True)
    w = torch.ones(2 * hd, device="cuda", requires_grad=True)
    do = torch.randn(Bn, N, C, device="cuda", dtype=torch.bfloat16)
    f = t_ms(lambda: att.pooled_diff_attention(q, kv, lam, w, h, hd, hd ** -0.5))
    def fb():
        o = att.pooled_diff_attention(q, kv, lam, w, h, hd, hd ** -0.5)
        torch.autograd.grad(o, [q, kv, lam, w], do)
    t = t_ms(fb)
    print(json.dumps({"N": N, "C": C, "fwd_ms": round(f, 4), "bwd_ms": round(t - f, 4),
                      "fwd_GFMA_per_s": round(Bn * N * h * P * 144 / f / 1e6, 1)}))
