"""BASELINE config 2, block part: MLLABlock(dim=256, heads=4, mlp_ratio=2, pooled 8x8) forward and forward+backward on
(10, 256, sqrt(L), sqrt(L)) inputs, L = 1k .. 64k, fp32 and bf16 autocast, CUDA-event timing.  (The CPU reference arm
lives in bench.py -- oracle/ is only imported by tests/, smoke() and bench.py.)"""
import json, os, sys
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
    print(json.dumps(row))
