"""Fused MSMM scan microbench at the config-3 shape (B=10, Di=96, 4 stages of a 320x320 input)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlagg_unet_b200.selective_scan_interface import msmm_scan

B, Di, N, R = 10, 96, 16, 3
hw = [(160, 160), (80, 80), (40, 40), (20, 20)]
lens = [h * w for h, w in hw]
L = sum(lens)
dev = "cuda"
torch.manual_seed(0)
xrow = torch.randn(B, Di, L, device=dev, requires_grad=True)
xcol = torch.randn(B, Di, L, device=dev, requires_grad=True)
xr = torch.randn(B, 2, R + 2 * N, L, device=dev, requires_grad=True)
xc = torch.randn(B, 2, R + 2 * N, L, device=dev, requires_grad=True)
Wdt = (torch.rand(4 * Di, R, device=dev) - 0.5).requires_grad_()
bias = torch.randn(4 * Di, device=dev).requires_grad_()
A = (-torch.arange(1, N + 1, device=dev, dtype=torch.float32).repeat(4 * Di, 1)).requires_grad_()
Ds = torch.ones(4 * Di, device=dev, requires_grad=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


with torch.no_grad():
    t_inf = timeit(lambda: msmm_scan(xrow, xcol, xr, xc, Wdt, bias, A, Ds, lens))
out = msmm_scan(xrow, xcol, xr, xc, Wdt, bias, A, Ds, lens)
g = torch.randn_like(out)
t_fwd = timeit(lambda: msmm_scan(xrow, xcol, xr, xc, Wdt, bias, A, Ds, lens))
def fb():
    o = msmm_scan(xrow, xcol, xr, xc, Wdt, bias, A, Ds, lens)
    o.backward(g)
t_fb = timeit(fb)
print(json.dumps({"fused_fwd_infer_ms": t_inf, "fused_fwd_train_ms": t_fwd, "fused_fwd_bwd_ms(incl. torch zero-fills)": t_fb}))
