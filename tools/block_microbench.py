"""BASELINE config 2, block part: MLLABlock(dim=256, heads=4, mlp_ratio=2, pooled 8x8) forward and forward+backward on
(10, 256, sqrt(L), sqrt(L)) inputs, L = 1k .. 64k, fp32 and bf16 autocast, CUDA-event timing; the reference's CPU path
(oracle.mlagg.mlla_block_forward, fp32, all host threads) is timed beside it at the smallest sizes."""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlagg_unet_b200.mlagg import MLLABlock

def t_ms(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

torch.manual_seed(0)
for side in (32, 64, 128, 256):
    L = side * side
    blk = MLLABlock(dim=256, input_resolution=(side, side), num_heads=4, mlp_ratio=2, sr_ratio=side // 8).cuda()
    x = torch.randn(10, 256, side, side, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_()
    row = {"L": L}
    for name, ctx in (("fp32", torch.autocast("cuda", enabled=False)), ("bf16", torch.autocast("cuda", dtype=torch.bfloat16))):
        def fwd():
            with ctx:
                return blk(x)
        def fb():
            with ctx:
                y = blk(x)
            y.float().sum().backward()
        row[name + "_fwd_ms"] = round(t_ms(fwd), 3)
        row[name + "_fwd_bwd_ms"] = round(t_ms(fb), 3)
    if side <= 64:   # CPU reference path (oracle), forward only
        from oracle.mlagg import mlla_block_forward
        p = {k: v.detach().cpu() for k, v in blk.named_parameters()}
        xc = x.detach().cpu().contiguous()
        torch.set_num_threads(os.cpu_count() or 1)
        with torch.no_grad():
            mlla_block_forward(p, xc, 4, side // 8)
            t = time.perf_counter(); mlla_block_forward(p, xc, 4, side // 8); row["cpu_oracle_fwd_ms"] = round((time.perf_counter() - t) * 1e3, 1)
        row["cpu_threads"] = torch.get_num_threads()
    print(json.dumps(row))
