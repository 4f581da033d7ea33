"""Selective-scan forward: this repo's sm_100a kernel next to a mamba-ssm-derived CUDA kernel on the same B200.

mamba-ssm itself (the reference's dependency, README.md:49-50) is not in the image; vLLM 0.22 is, and its
`selective_scan_fwd` op (csrc/mamba/mamba_ssm/selective_scan_fwd.cu) is mamba-ssm's forward kernel carried over with
varlen / state-cache additions -- the closest thing to "the reference's CUDA build" that can run here.  Forward only
(vLLM ships no backward).  Prints parity (max |a-b| / max |b|) and CUDA-event times with an L2 flush between iterations.

    python tools/mamba_kernel_compare.py [L ...]
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mlagg_unet_b200.selective_scan_interface import selective_scan_fn


def timeit(fn, iters=10):
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def check(vllm_scan):
    """small-shape parity of output and final state; prints `PARITY <rel out> <rel state>`"""
    g = torch.Generator(device="cuda").manual_seed(21)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)
    Bn, D, G, N, L = 2, 64, 4, 16, 3000
    u, dl = r(Bn, D, L), 0.5 * r(Bn, D, L)
    A = -torch.rand(D, N, device="cuda", generator=g) * 4 - 0.1
    Bm, Cm, Dk, bias = r(Bn, G, N, L), r(Bn, G, N, L), r(D), 0.3 * r(D)
    state = torch.zeros(Bn, D, N, device="cuda")
    ref = vllm_scan(u.clone(), state, dl.clone(), A, Bm, Cm, Dk, None, bias, True).clone()
    torch.cuda.synchronize()
    with torch.no_grad():
        out, last = selective_scan_fn(u, dl, A, Bm, Cm, Dk, None, bias, True, return_last_state=True)
    print("PARITY %.3e %.3e" % (float((out - ref).abs().max() / ref.abs().max()),
                                float((last - state).abs().max() / state.abs().max())))


def main():
    try:
        from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn as vllm_scan
    except Exception as e:
        print("vllm selective scan unavailable:", repr(e))
        return
    if "--check" in sys.argv:
        check(vllm_scan)
        return
    Ls = [int(v) for v in sys.argv[1:]] or [1024, 4096, 16384, 34000, 65536]
    Bn, D, G, N = 10, 384, 4, 16
    print(f"B={Bn} D={D} G={G} N={N} fp32, delta_softplus + delta_bias + D skip (the MSMM call, MambaSkip.py:445-451)")
    print(f"{'L':>7} {'ours fwd ms':>12} {'ours fwd+ckpt ms':>17} {'vllm(mamba-ssm) fwd ms':>23} {'speed-up':>9} {'rel diff':>10}")
    for L in Ls:
        g = torch.Generator(device="cuda").manual_seed(L)
        r = lambda *s: torch.randn(*s, device="cuda", generator=g)
        u, dl = r(Bn, D, L), 0.5 * r(Bn, D, L)
        A = -torch.exp(torch.log(torch.arange(1, N + 1, device="cuda", dtype=torch.float32)).repeat(D, 1))
        Bm, Cm = r(Bn, G, N, L), r(Bn, G, N, L)
        Dk, bias = torch.ones(D, device="cuda"), 0.1 * r(D)
        with torch.no_grad():
            ours = selective_scan_fn(u, dl, A, Bm, Cm, Dk, None, bias, True)
            t_inf = timeit(lambda: selective_scan_fn(u, dl, A, Bm, Cm, Dk, None, bias, True))
        ug = u.clone().requires_grad_()
        t_ck = timeit(lambda: selective_scan_fn(ug, dl, A, Bm, Cm, Dk, None, bias, True))
        try:
            state = torch.zeros(Bn, D, N, device="cuda")
            ref = vllm_scan(u.clone(), state, dl.clone(), A, Bm, Cm, Dk, None, bias, True).clone()
            dscr = dl.clone()

            def run():
                vllm_scan(u, state, dscr, A, Bm, Cm, Dk, None, bias, True)   # writes its output into dscr (in place)
            t_v = timeit(run)
            rel = float((ours - ref).abs().max() / ref.abs().max())
            print(f"{L:7d} {t_inf:12.3f} {t_ck:17.3f} {t_v:23.3f} {t_v / t_inf:8.2f}x {rel:10.2e}")
        except Exception as e:
            print(f"{L:7d} {t_inf:12.3f} {t_ck:17.3f}   vllm call failed: {repr(e)[:200]}")


if __name__ == "__main__":
    main()
