"""Linear-attention core microbench (BASELINE config 2: B=10, dim 256, heads 8, L = 1k..64k), fwd and fwd+bwd,
CUDA-event timing, algorithmic bytes 4*B*N*C*e fwd (q, k, v in, out) per SURVEY.md 8d."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlagg_unet_b200 import attention as att

def t_ms(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
Bn, C, heads = 10, 256, 8
for dt in (torch.float32, torch.bfloat16):
    for side in (32, 64, 128, 256):
        N = side * side
        qk = torch.randn(Bn, N, 2 * C, device="cuda", dtype=dt, requires_grad=True)
        v = torch.randn(Bn, N, C, device="cuda", dtype=dt, requires_grad=True)
        do = torch.randn(Bn, N, C, device="cuda", dtype=dt)
        e = qk.element_size()
        f = t_ms(lambda: att.linear_attention_qk(qk, v, side, side, heads))
        def fb():
            o = att.linear_attention_qk(qk, v, side, side, heads)
            torch.autograd.grad(o, [qk, v], do)
        fbm = t_ms(fb)
        fwd_bytes, bwd_bytes = 4 * Bn * N * C * e, 7 * Bn * N * C * e   # bwd: q,k,v,dO in; dq,dk,dv out
        print(json.dumps({"dtype": str(dt)[6:], "L": N, "fwd_ms": round(f, 4), "fwd_GBps": round(fwd_bytes / f / 1e6, 1),
                          "fwd_frac": round(fwd_bytes / f / 1e6 / peak, 3), "bwd_ms": round(fbm - f, 4),
                          "bwd_GBps": round(bwd_bytes / (fbm - f) / 1e6, 1)}))
