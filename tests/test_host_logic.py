"""CPU: host-side logic of the product (no kernels are launched): C-ABI symbols, index maps, state_dict / API
compatibility with the reference, scheduler + DropPath semantics, batch split, and that the product path refuses to
run on CPU tensors (there is no fallback)."""
import inspect

import numpy as np
import os
import re

import pytest
import torch

from conftest import ROOT, load_golden


def test_library_exports_every_declared_symbol():
    from mlagg_unet_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    hdr = open(os.path.join(ROOT, "include", "mlagg_b200.h")).read()
    declared = set(re.findall(r"\b(mlagg_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.mlagg_version() >= 100
    assert L.mlagg_scan_ckpt_bytes(10, 384, 34000, 16) == 10 * 2125 * 384 * 16 * 4
    assert L.mlagg_error_string(-2).decode() == "unsupported configuration"


def test_c_abi_rejects_bad_arguments_without_a_gpu():
    from mlagg_unet_b200 import _lib
    L = _lib.lib()
    assert L.mlagg_selective_scan_fwd(None, None, None, None, None, None, None, None, None, None,
                                      1, 8, 16, 16, 1, 1, None) == -3          # NULL
    assert L.mlagg_dwconv3x3_fwd(None, None, None, None, 1, 4, 4, 8, 0, 0, None) == -3
    assert L.mlagg_local_diffattn_ws_bytes(2, 4, 4, 1, 24) == 2 * 16 * (48 + 27) * 4


def test_cross_scan_maps_agree_with_oracle_and_are_inverse():
    from mlagg_unet_b200.mamba_skip import _scan_maps_cpu
    from oracle.msmm import cross_scan_maps
    hw = ((6, 5), (3, 4), (2, 2))
    idx, inv = _scan_maps_cpu(hw)
    assert torch.equal(idx, cross_scan_maps(list(hw)))
    L = idx.shape[1]
    for k in range(4):
        assert torch.equal(idx[k][inv[k]], torch.arange(L))


def test_state_dict_compatibility_with_reference():
    from mlagg_unet_b200 import mlla
    from mlagg_unet_b200.mamba_skip import VSS_Conv_Layer
    from mlagg_unet_b200.mlagg import MLLA_Uper, MLLABlock
    g = load_golden("mlla_uper_embed8.pt")
    net = MLLA_Uper(img_size=[64, 64], patch_size=2, in_channels=1, out_channels=5, embed_dim=8, depths=[2, 2, 2, 2],
                    num_heads=[2, 4, 8, 16], mlp_ratio=2, sr_ratio=[16, 8, 4, 2], dropout_path_rate=0.1)
    net.load_state_dict(g["state"], strict=True)
    assert sum(p.numel() for p in net.parameters()) == g["n_params"]
    assert not net.dummy_tensor.requires_grad  # F6: unused parameter must not take part in DDP reduction
    # parameter ORDER equals the reference's registration order (the golden's state_dict was written by the reference's
    # own MLLA_Uper): AdamW's `optimizer_state` indexes parameters by position, so checkpoints interchange only then
    ours = [k for k, _ in net.named_parameters()]
    ref_order = [k for k in g["state"] if k in set(ours)]
    assert ours == ref_order
    g = load_golden("msmm_vss_conv_layer.pt")
    VSS_Conv_Layer(g["dims"], g["hidden"], depth=1, drop_path=0.1).load_state_dict(g["state"], strict=True)
    g = load_golden("mlagg_block.pt")
    MLLABlock(g["dim"], (g["H"], g["W"]), g["num_heads"], mlp_ratio=2, sr_ratio=g["sr_ratio"]).load_state_dict(
        g["state"], strict=True)
    g = load_golden("mlla_block.pt")
    res = mlla.MLLABlock(g["dim"], (g["H"], g["W"]), g["num_heads"], mlp_ratio=2.0).load_state_dict(g["state"], strict=False)
    assert res.unexpected_keys == [] and res.missing_keys == ["attn.rope.rotations"]


def test_full_size_network_has_reference_parameter_count():
    from mlagg_unet_b200.trainer import SyntheticPlan, nnUNetTrainer_MLAgg_2D_dt_MS
    plan = SyntheticPlan()
    net = nnUNetTrainer_MLAgg_2D_dt_MS.build_network_architecture(plan, {}, plan, 1, True)
    n = sum(p.numel() for p in net.parameters())
    assert abs(n - 27.1e6) < 0.1e6  # ~27.1 M
    sig = inspect.signature(nnUNetTrainer_MLAgg_2D_dt_MS.build_network_architecture)
    assert list(sig.parameters) == ["plans_manager", "dataset_json", "configuration_manager", "num_input_channels",
                                    "enable_deep_supervision"]


def test_rope_buffer_matches_reference_layout():
    from mlagg_unet_b200.mlla import RoPE
    g = load_golden("mlla_linear_attention.pt")
    # rotations (H, W, C/2, 2): the rotated tensor of the golden input reproduces the reference's output on CPU
    r = RoPE((g["H"], g["W"], g["dim"]))
    assert r.rotations.shape == (g["H"], g["W"], g["dim"] // 2, 2)
    out = r(g["rope_in"].reshape(2, g["H"], g["W"], g["dim"])).reshape(g["rope_out"].shape)
    assert (out - g["rope_out"]).abs().max() < 1e-5


def test_cosine_scheduler_timm_semantics():
    from mlagg_unet_b200.thirdparty_shims import CosineLRScheduler
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.AdamW([p], 5e-4)
    s = CosineLRScheduler(opt, t_initial=500, lr_min=1e-6, warmup_t=10, warmup_lr_init=1e-4)
    assert opt.param_groups[0]["lr"] == pytest.approx(1e-4)
    s.step(5)
    assert opt.param_groups[0]["lr"] == pytest.approx(1e-4 + 5 * (5e-4 - 1e-4) / 10)
    s.step(250)
    assert opt.param_groups[0]["lr"] == pytest.approx(1e-6 + 0.5 * (5e-4 - 1e-6))
    s.step(500)
    assert opt.param_groups[0]["lr"] == pytest.approx(1e-6)


def test_droppath_semantics():
    from mlagg_unet_b200.thirdparty_shims import DropPath
    d = DropPath(0.5)
    x = torch.ones(1000, 3, 2)
    assert torch.equal(d.eval()(x), x)
    y = d.train()(x)
    per_sample = y.flatten(1)
    assert ((per_sample == 0).all(1) | (per_sample == 2).all(1)).all()  # whole sample dropped or scaled by 1/keep
    assert 0.35 < float((per_sample[:, 0] == 0).float().mean()) < 0.65


def test_split_batch_is_the_reference_arithmetic():
    from mlagg_unet_b200.trainer import split_batch
    assert split_batch(10, 2) == [5, 5]
    assert split_batch(10, 4) == [3, 3, 3, 1]
    assert split_batch(10, 8) == [2, 2, 2, 2, 2, 0, -2, -4]  # invalid in the reference itself
    with pytest.raises(AssertionError):
        split_batch(4, 8)


def test_product_path_has_no_cpu_fallback():
    from mlagg_unet_b200._lib import MlaggError
    from mlagg_unet_b200.attention import local_diff_attention, pooled_diff_attention
    from mlagg_unet_b200.ops import causal_conv1d_fn, dwconv3x3_tokens
    from mlagg_unet_b200.selective_scan_interface import selective_scan_fn
    x = torch.randn(1, 16, 8)
    with pytest.raises(MlaggError):
        dwconv3x3_tokens(x, torch.randn(8, 1, 3, 3), None, 4, 4)
    with pytest.raises(MlaggError):
        causal_conv1d_fn(torch.randn(1, 4, 8), torch.randn(4, 3))
    with pytest.raises(MlaggError):
        selective_scan_fn(torch.randn(1, 4, 8), torch.randn(1, 4, 8), -torch.rand(4, 16), torch.randn(1, 16, 8),
                          torch.randn(1, 16, 8))
    with pytest.raises(MlaggError):
        local_diff_attention(x, torch.randn(1, 16, 16), torch.tensor(0.5), torch.ones(8), 4, 4, 1, 4, 0.5)
    with pytest.raises(MlaggError):
        pooled_diff_attention(x, torch.randn(1, 4, 16), torch.tensor(0.5), torch.ones(8), 1, 4, 0.5)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mlagg-unet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "scan_ref" not in src, f


def test_sliding_window_steps_and_gaussian_follow_the_reference():
    """compute_steps_for_sliding_window / compute_gaussian (reference sliding_window_prediction.py:13-58): the worked
    example from the reference's own comment (image 110, tile 64, step 0.5 -> 0, 23, 46) and map properties."""
    from mlagg_unet_b200 import inference as inf
    assert inf.compute_steps_for_sliding_window((110,), (64,), 0.5) == [[0, 23, 46]]
    assert inf.compute_steps_for_sliding_window((64, 64), (64, 64), 0.5) == [[0], [0]]
    assert inf.compute_steps_for_sliding_window((320, 400), (320, 320), 0.5) == [[0], [0, 80]]
    g = inf.compute_gaussian((32, 48))
    assert g.shape == (32, 48) and g.max() == 1.0 and g[16, 24] == 1.0 and g.min() > 0
    assert inf._mirror_sets(None) == [()] and inf._mirror_sets((0, 1)) == [(), (2,), (3,), (2, 3)]


def test_rows2d_views_and_padded_xproj_layout():
    """Host logic of the strided-GEMM path and of the tokens-major x_proj layout (no kernels involved)."""
    from mlagg_unet_b200.ops import _rows2d
    from mlagg_unet_b200.selective_scan_interface import xdbl_pad
    t = torch.arange(3 * 10 * 16, dtype=torch.float32).reshape(3, 10, 16)
    a, b = t.chunk(2, dim=-1)
    v = _rows2d(b)
    assert v.shape == (30, 8) and v.stride() == (16, 1) and v.data_ptr() == b.data_ptr()
    assert torch.equal(v, b.reshape(30, 8))
    assert _rows2d(t).shape == (30, 16) and _rows2d(t[:, :5]) is None          # batch stride does not collapse
    assert _rows2d(t.transpose(1, 2)) is None and _rows2d(t[0]).shape == (10, 16)
    assert xdbl_pad(35) == 72 and xdbl_pad(34) == 68 and xdbl_pad(33) == 68     # 2 directions x (R + 2N), 4-aligned


def test_walk_token_formula_is_the_cross_scan_map():
    """csrc/walk.cu visits token soff + (q % H) * W + q / H at column-walk position soff + q; that is direction 1 of the
    cross-scan index maps (reference MambaSkip.py:414-422), which the golden fixtures already pin."""
    from mlagg_unet_b200.mamba_skip import _scan_maps_cpu
    hw = ((5, 3), (2, 4), (1, 1), (3, 2))
    idx, inv = _scan_maps_cpu(hw)
    toks, off = [], 0
    for H, W in hw:
        toks += [off + (q % H) * W + q // H for q in range(H * W)]
        off += H * W
    assert idx[1].tolist() == toks and idx[0].tolist() == list(range(off))
    assert idx[3].tolist() != toks and sorted(toks) == list(range(off))
    assert all(inv[1][t] == p for p, t in enumerate(toks))


def test_stride_probes_for_in_place_operands():
    """`_tok_strides` / `_rows3` decide whether a view is handed to the kernels in place or copied first: channel slices
    and stage segments qualify, transposed or misaligned views do not."""
    from mlagg_unet_b200.ops import _rows3, _tok_strides
    t = torch.zeros(2, 30, 16)
    assert _tok_strides(t, 16) == (16, 480)
    assert _tok_strides(t[..., 8:], 8) == (16, 480)                 # v half of a kv projection
    assert _tok_strides(t[:, 5:15], 16) == (16, 480)                # a stage segment keeps the parent's image stride
    assert _tok_strides(t[..., 2:10], 8) is None                    # 4-channel vectors would be misaligned
    assert _tok_strides(t[..., 2:9], 7) == (16, 480)                # scalar kernels (C % 4 != 0) take any offset
    assert _tok_strides(t.transpose(1, 2), 30) is None
    x = torch.zeros(2, 12, 5, 6).contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)    # (B, H, W, C) view
    r = _rows3(x[..., 4:])
    assert r is not None and r[1:] == (2, 30, 8, 12, 360)
    assert _rows3(x.permute(0, 2, 1, 3)) is None                    # rows no longer collapse to one stride
    assert _rows3(torch.zeros(2, 7, 3)[:, 1:5])[1:] == (2, 4, 3, 3, 21)


def test_overlay_files_define_what_the_reference_looks_up():
    """overlay/ mirrors the two reference paths run_training.py:39-40 / the trainer's own import resolve; nnunetv2 is
    absent in this image, so the check is structural: the files parse, sit at the reference's relative paths, and define
    / re-export the names the reference imports."""
    import ast
    base = os.path.join(ROOT, "overlay", "nnunetv2", "training", "nnUNetTrainer")
    tr = ast.parse(open(os.path.join(base, "nnUNetTrainer_MLAgg_2D_dt_MS.py")).read())
    cls = [n for n in tr.body if isinstance(n, ast.ClassDef) and n.name == "nnUNetTrainer_MLAgg_2D_dt_MS"]
    assert len(cls) == 1 and [b.id for b in cls[0].bases] == ["nnUNetTrainer"]
    defined = {n.name for n in cls[0].body if isinstance(n, ast.FunctionDef)} | {
        t.id for n in cls[0].body if isinstance(n, ast.Assign) for t in n.targets if isinstance(t, ast.Name)}
    assert {"build_network_architecture", "set_deep_supervision_enabled", "_get_deep_supervision_scales", "_build_loss",
            "configure_optimizers", "train_step"} <= defined
    ms = ast.parse(open(os.path.join(base, "variants", "mamba", "MambaSkip.py")).read())
    names = {a.name for n in ms.body if isinstance(n, ast.ImportFrom) for a in n.names}
    assert {"SS2D_skip", "DWConv", "ConvolutionalGLU", "VSS_Conv_Block", "VSS_Conv_Layer", "selective_scan_fn"} <= names


def test_training_logger_and_epoch_bookkeeping_follow_the_reference():
    """TrainingLogger == nnUNetLogger's list-per-key bookkeeping incl. the 0.9 / 0.1 EMA (nnunet_logger.py:17-51); the
    scheduler is stepped once per epoch with the epoch index (nnUNetTrainer.py:825)."""
    from mlagg_unet_b200.trainer import TrainingLogger
    lg = TrainingLogger()
    for e, v in enumerate([0.2, 0.6, 0.4]):
        lg.log("mean_fg_dice", v, e)
    ema = lg.my_fantastic_logging["ema_fg_dice"]
    assert ema[0] == 0.2 and abs(ema[1] - (0.2 * 0.9 + 0.06)) < 1e-12 and abs(ema[2] - (ema[1] * 0.9 + 0.04)) < 1e-12
    lg.log("mean_fg_dice", 0.5, 2)                       # re-logging the same epoch overwrites
    assert len(lg.my_fantastic_logging["mean_fg_dice"]) == 3 and lg.my_fantastic_logging["mean_fg_dice"][2] == 0.5
    with pytest.raises(AssertionError):
        lg.log("train_losses", 1.0, 5)                   # exactly one value per epoch


def test_leaf_param_resolves_parameters_and_dense_views_only():
    """_lib.leaf_param decides which weight gradients may be produced on the side stream and handed to the trainer
    (never through autograd): the Parameter itself or a same-size dense view of it; not a padded / permuted copy, not a
    frozen parameter, not a slice."""
    import torch
    from mlagg_unet_b200 import _lib
    lin = torch.nn.Linear(6, 4)
    conv = torch.nn.Conv2d(6, 4, 1).to(memory_format=torch.channels_last)
    assert _lib.leaf_param(lin.weight) is lin.weight and _lib.leaf_param(lin.bias) is lin.bias
    assert _lib.leaf_param(lin.weight.view(24)) is lin.weight
    assert _lib.leaf_param(conv.weight.view(4, 6)) is conv.weight
    assert _lib.leaf_param(None) is None
    assert _lib.leaf_param(torch.nn.functional.pad(lin.weight, (0, 0, 0, 4))) is None      # the padded segmentation head
    assert _lib.leaf_param(lin.weight.t().reshape(24)) is None                              # a copy
    assert _lib.leaf_param(lin.weight[:2]) is None                                          # a slice
    assert _lib.leaf_param(lin.weight.detach()) is None
    lin.weight.requires_grad_(False)
    assert _lib.leaf_param(lin.weight) is None


def test_side_stream_helpers_are_inert_outside_a_trainer_step():
    import torch
    from mlagg_unet_b200 import _lib
    assert not _lib.side_active()
    with _lib.side_launch(torch.zeros(1)) as on_side:
        assert on_side is False
    p = torch.nn.Parameter(torch.zeros(2, 3))
    _lib.stash_grad(p, torch.ones(6))
    (q, g), = _lib.take_stashed_grads()
    assert q is p and g.shape == p.shape and _lib.take_stashed_grads() == []
