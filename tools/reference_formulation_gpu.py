"""GPU-vs-GPU comparators for north_star's ">= 1.5x on MLAgg + MSMM fwd+bwd" (the reference's CUDA build of
mamba-ssm cannot be installed here, so its stand-ins are the reference's OWN torch-op formulation run on the same B200).

TEST / BENCH INFRASTRUCTURE: this tool imports `oracle/` (the restatement of the reference's op sequence) and runs it on
CUDA tensors; nothing under mlagg-unet_b200/ does.

  pooled   the pooled differential-attention core (reference nnUNetTrainer_MLAgg_2D_dt_MS.py:745-760): 4 x
           flash_attn_func (flash-attn 2.8.3, sm_100 cubins) + cat + lambda-combine + RMSNorm + scale, against
           mlagg_pooled_diffattn_* -- at the four stage shapes of config 3 and of config 5 (SURVEY K8's bar: >= 1.5x)
  block    MLLABlock forward and forward+backward at BASELINE config 2 (B = 10, C = 256, L = 1k .. 64k): the reference
           formulation (oracle.mlagg.mlla_block_forward: permute / gather-unfold / softmax / einsum torch ops, the pooled
           branch through flash_attn_func) against the product block, bf16 autocast
  msmm     VSS_Conv_Block forward at the config-3 stage shapes: oracle.msmm.vss_conv_block_forward with the materialised
           4-direction cross-scan and vLLM's mamba-ssm-derived `selective_scan_fwd` as the scan, against the product
           module (fused operand addressing).  Forward only: no backward kernel of mamba-ssm exists in the image.

    python tools/reference_formulation_gpu.py [pooled] [block] [msmm]
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mlagg_unet_b200 import attention as att  # noqa: E402
from oracle import mlagg as o_mlagg  # noqa: E402
from oracle import msmm as o_msmm  # noqa: E402

flush = None


def t_ms(fn, n=10, warm=3):
    global flush
    if flush is None:
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        flush.max()                       # read flush: clean lines only
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def fa2_pooled(q, kp, vp, lam, w, h, hd):
    """the reference's pooled branch with flash_attn_func (:745-760); q (B,N,C) already scaled ONCE, kvp split"""
    from flash_attn import flash_attn_func
    Bn, N, C = q.shape
    P = kp.shape[1]
    q4 = q.view(Bn, N, h, 2, hd)
    k4 = kp.view(Bn, P, h, 2, hd)
    v4 = vp.view(Bn, P, h, 2, hd)
    q1, q2 = q4[:, :, :, 0], q4[:, :, :, 1]
    k1, k2 = k4[:, :, :, 0], k4[:, :, :, 1]
    v1, v2 = v4[:, :, :, 0], v4[:, :, :, 1]
    a11, a12 = flash_attn_func(q1, k1, v1), flash_attn_func(q1, k1, v2)
    a21, a22 = flash_attn_func(q2, k2, v1), flash_attn_func(q2, k2, v2)
    o = torch.cat([a11, a12], -1) - lam.to(q.dtype) * torch.cat([a21, a22], -1)        # (B, N, h, 2hd)
    o = o_mlagg.rmsnorm(o, w, 1e-5) * (1 - o_mlagg.LAMBDA_INIT)
    return o.reshape(Bn, N, C)


def run_pooled():
    rows = []
    for cfg, stages, Bn in (("config 3 (10 x 320^2)", [(25600, 48, 100), (6400, 96, 100), (1600, 192, 100), (400, 384, 100)], 10),
                            ("config 5 (4 x 512^2)", [(65536, 48, 256), (16384, 96, 256), (4096, 192, 256), (1024, 384, 256)], 4)):
        for N, C, P in stages:
            h, hd = C // 48, 24
            g = torch.Generator(device="cuda").manual_seed(N)
            q = torch.randn(Bn, N, C, device="cuda", generator=g).bfloat16().requires_grad_()
            kv = torch.randn(Bn, P, 2 * C, device="cuda", generator=g).bfloat16().requires_grad_()
            lam = torch.tensor(0.8, device="cuda", requires_grad=True)
            w = torch.ones(2 * hd, device="cuda", requires_grad=True)
            do = torch.randn(Bn, N, C, device="cuda", generator=g).bfloat16()
            ours = lambda: att.pooled_diff_attention(q, kv, lam, w, h, hd, hd ** -0.5)
            ref = lambda: fa2_pooled(q * hd ** -0.5, kv[..., :C], kv[..., C:], lam, w, h, hd)
            with torch.no_grad():
                a, b = ours().float(), ref().float()
            rel = float((a - b).abs().max() / b.abs().max())
            fo, fr = t_ms(ours), t_ms(ref)
            bo = t_ms(lambda: torch.autograd.grad(ours(), [q, kv, lam, w], do))
            br = t_ms(lambda: torch.autograd.grad(ref(), [q, kv, lam, w], do))
            rows.append({"config": cfg, "N": N, "C": C, "P": P, "ours_fwd_ms": round(fo, 4), "fa2_fwd_ms": round(fr, 4),
                         "fwd_speedup": round(fr / fo, 2), "ours_fwd_bwd_ms": round(bo, 4), "fa2_fwd_bwd_ms": round(br, 4),
                         "fwd_bwd_speedup": round(br / bo, 2), "rel_diff_bf16": rel})
            print(json.dumps(rows[-1]), flush=True)
    return rows


def run_block():
    from mlagg_unet_b200.mlagg import MLLABlock
    orig = o_mlagg.pooled_diff_attention

    def pooled_fa2(q, kp, vp, lam, subln_w):          # oracle signature: q (B,N,h,2,hd) scaled once, kp (B,P,h,2,hd), vp (B,P,h,2hd)
        Bn, N, h, _, hd = q.shape
        P = kp.shape[1]
        return fa2_pooled(q.reshape(Bn, N, -1), kp.reshape(Bn, P, -1), vp.reshape(Bn, P, -1), lam, subln_w, h, hd)

    rows = []
    torch.manual_seed(0)
    for side in (32, 64, 128, 256):
        blk = MLLABlock(dim=256, input_resolution=(side, side), num_heads=4, mlp_ratio=2, sr_ratio=side // 8).cuda()
        p = dict(blk.named_parameters())
        x = torch.randn(10, 256, side, side, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_()
        xr = x.detach().contiguous().requires_grad_()              # the reference runs NCHW-contiguous

        def ours_f():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return blk(x)

        def ref_f():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return o_mlagg.mlla_block_forward(p, xr, blk.num_heads, blk.sr_ratio)

        o_mlagg.pooled_diff_attention = pooled_fa2
        try:
            with torch.no_grad():
                a, b = ours_f().float(), ref_f().float()
            row = {"L": side * side, "rel_diff_bf16": float((a - b).abs().max() / b.abs().max()),
                   "ours_fwd_ms": round(t_ms(ours_f), 3), "ref_fwd_ms": round(t_ms(ref_f), 3)}
            row["ours_fwd_bwd_ms"] = round(t_ms(lambda: ours_f().float().sum().backward()), 3)
            row["ref_fwd_bwd_ms"] = round(t_ms(lambda: ref_f().float().sum().backward()), 3)
        finally:
            o_mlagg.pooled_diff_attention = orig
        row["fwd_speedup"] = round(row["ref_fwd_ms"] / row["ours_fwd_ms"], 2)
        row["fwd_bwd_speedup"] = round(row["ref_fwd_bwd_ms"] / row["ours_fwd_bwd_ms"], 2)
        rows.append(row)
        print(json.dumps(row), flush=True)
    return rows


def run_msmm():
    try:
        from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn as vllm_scan
    except Exception as e:
        print("vllm selective scan unavailable:", repr(e))
        return []
    from mlagg_unet_b200.mamba_skip import VSS_Conv_Layer

    def scan(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False):
        state = torch.zeros(u.shape[0], u.shape[1], A.shape[1], device=u.device)
        return vllm_scan(u, state, delta, A, B, C, D, None, delta_bias, delta_softplus)

    rows = []
    torch.manual_seed(0)
    for Bn, size in ((10, 320), (4, 512)):
        dims = [96, 192, 384, 768]
        hw = [(size // 2 // 2 ** i,) * 2 for i in range(4)]
        layer = VSS_Conv_Layer(dims, 48, depth=1, drop_path=0.1).cuda().eval()
        p = {"blocks.0." + k: v for k, v in layer.blocks[0].named_parameters()}
        xs = [torch.randn(Bn, c, h, w, device="cuda").contiguous(memory_format=torch.channels_last) for c, (h, w) in zip(dims, hw)]
        xr = [t.contiguous() for t in xs]
        with torch.no_grad():
            def ours():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return layer(xs)

            def ref():
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    return o_msmm.vss_conv_block_forward(p, xr, 48, scan=scan, prefix="blocks.0.")
            a, b = ours(), ref()
            rel = max(float((u.float() - v.float()).abs().max() / v.float().abs().max()) for u, v in zip(a, b))
            row = {"input": f"{Bn} x {size}^2", "L_cat": sum(h * w for h, w in hw), "rel_diff_bf16": rel,
                   "ours_fwd_ms": round(t_ms(ours, n=5), 3), "ref_fwd_ms": round(t_ms(ref, n=5), 3)}
        row["fwd_speedup"] = round(row["ref_fwd_ms"] / row["ours_fwd_ms"], 2)
        rows.append(row)
        print(json.dumps(row), flush=True)
    return rows


if __name__ == "__main__":
    which = sys.argv[1:] or ["pooled", "block", "msmm"]
    out = {}
    for w in which:
        print("==", w, flush=True)
        out[w] = {"pooled": run_pooled, "block": run_block, "msmm": run_msmm}[w]()
