"""Local 3x3 differential attention at the four stage shapes of config 3 (B = 10, hd = 24), bf16: the thread-per-token
kernels (MLAGG_LOCAL_UNTILED=1) next to the shared-memory tiled ones; forward and the two backward passes, L2 flushed."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlagg_unet_b200 import attention as att  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t_us(fn, n=10):
    for _ in range(3):
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


SHAPES = ((160, 48), (80, 96), (40, 192), (20, 384))
if len(sys.argv) > 1:
    SHAPES = tuple(sc for sc in SHAPES if sc[0] == int(sys.argv[1]))
for side, C in SHAPES:
    h, hd, Bn, N = C // 48, 24, 10, side * side
    q = torch.randn(Bn, N, C, device="cuda").bfloat16().requires_grad_()
    kv = torch.randn(Bn, N, 2 * C, device="cuda").bfloat16().requires_grad_()
    lam = torch.tensor(0.8, device="cuda", requires_grad=True)
    w = torch.ones(2 * hd, device="cuda", requires_grad=True)
    do = torch.randn(Bn, N, C, device="cuda").bfloat16()
    row = {"side": side, "h": h}
    for name, env in (("untiled", "1"), ("tiled", "")):
        if env:
            os.environ["MLAGG_LOCAL_UNTILED"] = env
        else:
            os.environ.pop("MLAGG_LOCAL_UNTILED", None)
        f = t_us(lambda: att.local_diff_attention(q, kv, lam, w, side, side, h, hd, hd ** -0.5))
        fb = t_us(lambda: torch.autograd.grad(att.local_diff_attention(q, kv, lam, w, side, side, h, hd, hd ** -0.5),
                                              [q, kv, lam, w], do))
        row[name + "_fwd_us"] = round(f, 1)
        row[name + "_fwd_bwd_us"] = round(fb, 1)
    print(json.dumps(row), flush=True)
