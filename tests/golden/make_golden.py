"""Generate tests/golden/*.pt by RUNNING THE REFERENCE'S OWN MODULE SOURCE in this container.

Usage (only where /root/reference exists -- it does not exist on the GPU box):
    python tests/golden/make_golden.py            # all round-1 fixtures
    python tests/golden/make_golden.py --hd24     # round 2: the hd = 24 / P = 100 attention + block fixtures only

Method: the reference package cannot be imported (batchgenerators, timm, monai,
mamba_ssm ... are absent), but the three hot-path files are pure torch once their third-party import
lines are dropped.  Their source text is read from /root/reference at generation time, exec'd in a
namespace holding stand-ins for the missing symbols, and the resulting classes are run on seeded
inputs.  Nothing is copied into this repo: only the produced tensors are committed.

Stand-ins injected (all un-vendored third-party code, so outside /root/reference anyway):
    DropPath / to_2tuple / trunc_normal_ / UnetrBasicBlock / UnetrUpBlock -> mlagg_unet_b200.thirdparty_shims
    selective_scan_fn  -> oracle.scan.selective_scan_oracle (fp32 C restatement; PARITY UNPINNED vs mamba-ssm)
    flash_attn_func    -> softmax(q k^T / sqrt(d)) v  via torch SDPA (flash-attn's published definition)
So these fixtures pin everything the REFERENCE REPO itself computes around those two calls: index maps,
projections, masks, lambda / sub-LN conventions, the double softmax scale, norms, gates, LePE.
"""
from __future__ import annotations

import os
import re
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/mlagg/nnunetv2/training/nnUNetTrainer"

from mlagg_unet_b200 import thirdparty_shims as shims  # noqa: E402
from oracle.scan import selective_scan_oracle  # noqa: E402


def flash_attn_func(q, k, v, causal=False):
    assert not causal
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    return o.transpose(1, 2)


def load_reference(path, start_marker=None, extra=None):
    src = open(path).read()
    if start_marker is not None:
        src = src[src.index(start_marker):]
    src = src[: src.index("if __name__")]
    src = "\n".join(l for l in src.split("\n")
                    if not re.match(r"^\s*from (timm|monai|nnunetv2|flash_attn|dynamic_network_architectures)", l)
                    and not re.match(r"^import (timm|thop)", l))
    ns = {
        "DropPath": shims.DropPath, "to_2tuple": shims.to_2tuple, "trunc_normal_": shims.trunc_normal_,
        "UnetrBasicBlock": shims.UnetrBasicBlock, "UnetrUpBlock": shims.UnetrUpBlock,
        "selective_scan_fn": selective_scan_oracle, "flash_attn_func": flash_attn_func,
        "__name__": "reference_exec",
        "torch": torch, "nn": torch.nn, "F": F, "np": __import__("numpy"), "math": __import__("math"),
    }
    ns.update(extra or {})
    exec(compile(src, path, "exec"), ns)
    return ns


def sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def grads(out, wrt):
    torch.manual_seed(99)
    if isinstance(out, (list, tuple)):
        loss = sum((o * torch.randn_like(o)).sum() for o in out)
    else:
        loss = (out * torch.randn_like(out)).sum()
    return [g.detach().clone() for g in torch.autograd.grad(loss, wrt, allow_unused=True)]


def hd24(mlagg):
    """Round 2: fixtures at the SHIPPED attention geometry (hd = 24, P = 100; reference :71-89 gives every stage hd 24):
    the two AggregatedAttention branches with h = 2 and one stage-0-shaped MLLABlock (dim 96, heads 2 -> h = 1)."""
    torch.manual_seed(24)
    H, W, dim, h = 10, 10, 96, 2                       # branch dim 96 = 2 * h * 24; sr 1 -> P = N = 100
    for local in (True, False):
        att = mlagg["AggregatedAttention"](dim, (H, W), num_heads=h, local=local, sr_ratio=1).eval()
        x = torch.randn(2, H * W, dim, requires_grad=True)
        y = att(x, H, W)
        g = grads(y, [x] + list(att.parameters()))
        torch.save({"H": H, "W": W, "dim": dim, "num_heads": h, "sr_ratio": 1, "local": local,
                    "state": sd(att), "input": x.detach(), "output": y.detach(), "grad_input": g[0],
                    "grad_params": {n_: g_ for (n_, _), g_ in zip(att.named_parameters(), g[1:])}},
                   f"{HERE}/mlagg_attention_{'local' if local else 'pooled'}_hd24.pt")
    H, W, dim, heads, sr = 20, 20, 96, 2, 2            # stage-0 block of the shipped network: pooled 10 x 10 = 100
    blk = mlagg["MLLABlock"](dim, (H, W), heads, mlp_ratio=2, sr_ratio=sr, drop_path=0.05).eval()
    x = torch.randn(2, dim, H, W, requires_grad=True)
    y = blk(x)
    g = grads(y, [x] + list(blk.parameters()))
    torch.save({"H": H, "W": W, "dim": dim, "num_heads": heads, "sr_ratio": sr, "state": sd(blk),
                "input": x.detach(), "output": y.detach(), "grad_input": g[0],
                "grad_params": {n_: g_ for (n_, _), g_ in zip(blk.named_parameters(), g[1:])}},
               f"{HERE}/mlagg_block_hd24.pt")


def main():
    torch.set_num_threads(8)
    if "--hd24" in sys.argv:          # add the round-2 fixtures without rewriting the round-1 files
        mamba = load_reference(f"{REF}/variants/mamba/MambaSkip.py")
        hd24(load_reference(f"{REF}/nnUNetTrainer_MLAgg_2D_dt_MS.py", "import sys\nimport torch.utils.checkpoint",
                            {"VSS_Conv_Layer": mamba["VSS_Conv_Layer"]}))
        return
    mamba = load_reference(f"{REF}/variants/mamba/MambaSkip.py")
    mlagg = load_reference(f"{REF}/nnUNetTrainer_MLAgg_2D_dt_MS.py", "import sys\nimport torch.utils.checkpoint",
                           {"VSS_Conv_Layer": mamba["VSS_Conv_Layer"]})
    mlla = load_reference(f"{REF}/nnUNetTrainer_MLLA_UNet.py", "import torch.utils.checkpoint")

    # ---------------- MSMM: SS2D_skip + VSS_Conv_Layer, three NON-square stages ----------------
    torch.manual_seed(1)
    hw = [(6, 5), (3, 4), (2, 2)]
    dims, hidden = [12, 16, 20], 8
    layer = mamba["VSS_Conv_Layer"](dims, hidden, depth=1, drop_path=0.1).eval()
    with torch.no_grad():  # break the symmetric init so every parameter matters
        for n_, p_ in layer.named_parameters():
            if "A_logs" in n_:
                p_.add_(0.3 * torch.randn_like(p_))
            elif "Ds" in n_:
                p_.add_(0.5 * torch.randn_like(p_))
    xs = [torch.randn(2, c, h, w, requires_grad=True) for c, (h, w) in zip(dims, hw)]
    outs = layer(xs)
    params = list(layer.parameters())
    g = grads(outs, xs + params)
    torch.save({"hw": hw, "dims": dims, "hidden": hidden, "state": sd(layer), "inputs": [x.detach() for x in xs],
                "outputs": [o.detach() for o in outs], "grad_inputs": g[: len(xs)],
                "grad_params": {n_: g_ for (n_, _), g_ in zip(layer.named_parameters(), g[len(xs):])}},
               f"{HERE}/msmm_vss_conv_layer.pt")

    ss = layer.blocks[0].self_attention
    L = sum(h * w for h, w in hw)
    xt = torch.randn(2, L, hidden, requires_grad=True)
    yt = ss(xt, 2, [h for h, _ in hw], [w for _, w in hw], [h * w for h, w in hw])
    torch.save({"hw": hw, "state": sd(ss), "input": xt.detach(), "output": yt.detach(),
                "grad_input": grads(yt, [xt])[0]}, f"{HERE}/msmm_ss2d_skip.pt")

    # ---------------- MLAgg: AggregatedAttention (local / pooled) and MLLABlock ----------------
    torch.manual_seed(2)
    H, W, dim, heads, sr = 8, 6, 32, 4, 2  # branch dim 16, h=2, hd=4; pooled 4x3
    for local in (True, False):
        att = mlagg["AggregatedAttention"](dim // 2, (H, W), num_heads=heads // 2, local=local, sr_ratio=sr).eval()
        x = torch.randn(2, H * W, dim // 2, requires_grad=True)
        y = att(x, H, W)
        g = grads(y, [x] + list(att.parameters()))
        torch.save({"H": H, "W": W, "dim": dim // 2, "num_heads": heads // 2, "sr_ratio": sr, "local": local,
                    "state": sd(att), "input": x.detach(), "output": y.detach(), "grad_input": g[0],
                    "grad_params": {n_: g_ for (n_, _), g_ in zip(att.named_parameters(), g[1:])}},
                   f"{HERE}/mlagg_attention_{'local' if local else 'pooled'}.pt")
    blk = mlagg["MLLABlock"](dim, (H, W), heads, mlp_ratio=2, sr_ratio=sr, drop_path=0.05).eval()
    x = torch.randn(2, dim, H, W, requires_grad=True)
    y = blk(x)
    g = grads(y, [x] + list(blk.parameters()))
    torch.save({"H": H, "W": W, "dim": dim, "num_heads": heads, "sr_ratio": sr, "state": sd(blk),
                "input": x.detach(), "output": y.detach(), "grad_input": g[0],
                "grad_params": {n_: g_ for (n_, _), g_ in zip(blk.named_parameters(), g[1:])}},
               f"{HERE}/mlagg_block.pt")

    # ---------------- MLLA-UNet: RoPE, LinearAttention (elu+1), MLLABlock ----------------
    torch.manual_seed(3)
    H, W, dim, heads = 6, 8, 32, 4
    la = mlla["LinearAttention"](dim, (H, W), heads).eval()
    x = torch.randn(2, H * W, dim, requires_grad=True)
    y = la(x)
    g = grads(y, [x] + list(la.parameters()))
    torch.save({"H": H, "W": W, "dim": dim, "num_heads": heads,
                "state": {k: v for k, v in sd(la).items() if "rotations" not in k},
                "rope_in": x.detach(), "rope_out": la.rope(x.detach().reshape(2, H, W, dim)).reshape(2, H * W, dim),
                "input": x.detach(), "output": y.detach(), "grad_input": g[0],
                "grad_params": {n_: g_ for (n_, _), g_ in zip(la.named_parameters(), g[1:])}},
               f"{HERE}/mlla_linear_attention.pt")
    blk = mlla["MLLABlock"](dim, (H, W), heads, mlp_ratio=2.0, drop_path=0.05).eval()
    x = torch.randn(2, H * W, dim, requires_grad=True)
    y = blk(x)
    torch.save({"H": H, "W": W, "dim": dim, "num_heads": heads,
                "state": {k: v for k, v in sd(blk).items() if "rotations" not in k},
                "input": x.detach(), "output": y.detach(), "grad_input": grads(y, [x])[0]},
               f"{HERE}/mlla_block.pt")

    # ---------------- full network, narrow (embed 8) so the fixture stays small ----------------
    torch.manual_seed(4)
    net = mlagg["MLLA_Uper"](img_size=[64, 64], patch_size=2, in_channels=1, out_channels=5, embed_dim=8,
                             depths=[2, 2, 2, 2], num_heads=[2, 4, 8, 16], mlp_ratio=2, qkv_bias=True,
                             drop_rate=0.0, dropout_path_rate=0.1, sr_ratio=[16, 8, 4, 2], deep_supervision=True).eval()
    x = torch.randn(2, 1, 64, 64)
    with torch.no_grad():
        outs = net(x)
    state = sd(net)
    torch.save({"state": state, "input": x, "logits0": outs[0], "argmax0": outs[0].argmax(1).to(torch.uint8),
                "ds_shapes": [tuple(o.shape) for o in outs], "ds_sums": [float(o.double().sum()) for o in outs],
                "n_params": sum(p.numel() for p in net.parameters())}, f"{HERE}/mlla_uper_embed8.pt")
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
