"""Walk pack / unpack kernels (csrc/walk.cu) at the config-3 shape: B = 10, stages 160^2 + 80^2 + 40^2 + 20^2 (L = 34 000),
96 channels, bf16 tokens <-> fp32 planes.  CUDA events, L2 flushed between iterations; GB/s = (source + destination bytes) / time.

    python tools/walk_microbench.py          # table
    python tools/walk_microbench.py one      # one launch of each column-walk kernel (for `ncu -k regex:col2d`)
"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mlagg_unet_b200 import _lib

hw = [(160, 160), (80, 80), (40, 40), (20, 20)]
Bn, C = 10, 96
L = sum(h * w for h, w in hw)
ns = len(hw)
Hs, Ws = (ctypes.c_int * ns)(*[h for h, _ in hw]), (ctypes.c_int * ns)(*[w for _, w in hw])
Lb = _lib.lib()
tok = torch.randn(Bn, L, C, device="cuda").bfloat16()
plane = torch.empty(Bn, C, L, device="cuda")
p2 = torch.randn(Bn, 2, C, L, device="cuda")
out = torch.empty(Bn, L, C, device="cuda").bfloat16()


def pack(col):
    _lib.check(Lb.mlagg_walk_pack(tok.data_ptr(), 1, C, L * C, 0, C, plane.data_ptr(), C * L, Bn, ns, Hs, Ws, col,
                                  _lib.stream_ptr()), "pack")


def unpack(col):
    _lib.check(Lb.mlagg_walk_unpack(p2.data_ptr(), p2.data_ptr() + C * L * 4, 2 * C * L, C, C, out.data_ptr(), 1, C, L * C, 0,
                                    Bn, ns, Hs, Ws, col, 0, _lib.stream_ptr()), "unpack")


def timeit(fn, iters=10):
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
    fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


if len(sys.argv) > 1 and sys.argv[1] == "one":
    pack(1); unpack(1); torch.cuda.synchronize()
else:
    n = Bn * L * C
    for name, fn, byts in (("pack   row", lambda: pack(0), n * 2 + n * 4), ("pack   col", lambda: pack(1), n * 2 + n * 4),
                           ("unpack row (2 planes)", lambda: unpack(0), 2 * n * 4 + n * 2),
                           ("unpack col (2 planes)", lambda: unpack(1), 2 * n * 4 + n * 2)):
        ms = timeit(fn)
        print(f"{name:24s} {ms * 1e3:8.1f} us  {byts / ms / 1e6:8.0f} GB/s")
