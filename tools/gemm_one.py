"""One projection GEMM shape, a few launches (for `ncu -k regex:gemm_tc` captures): python tools/gemm_one.py M N K [fwd|dx|dw]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mlagg_unet_b200 import gemm  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4])
which = sys.argv[4] if len(sys.argv) > 4 else "fwd"
x = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
b = torch.randn(N, device="cuda")
dy = torch.randn(M, N, device="cuda").bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5):
    flush.max()
    if which == "fwd":
        gemm.linear_fwd(x, w, b)
    elif which == "dx":
        gemm.linear_bwd_data(dy, w)
    else:
        gemm.linear_bwd_weight(dy, x, want_db=True)
torch.cuda.synchronize()
print("ok")
