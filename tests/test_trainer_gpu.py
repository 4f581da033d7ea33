"""GPU: the epoch loop of the trainer (run_training and its hooks, reference nnUNetTrainer.py:784-1223) on a small
synthetic plan: learning-rate schedule stepped once per epoch, train / validation bookkeeping, pseudo dice from hard
tp / fp / fn, checkpoints written and resumable."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _trainer(tmp_path=None, epochs=2):
    from mlagg_unet_b200.trainer import SyntheticPlan, nnUNetTrainer_MLAgg_2D_dt_MS
    torch.manual_seed(0)
    tr = nnUNetTrainer_MLAgg_2D_dt_MS(SyntheticPlan(patch_size=(64, 64), batch_size=2, num_classes=5))
    tr.num_epochs, tr.num_iterations_per_epoch, tr.num_val_iterations_per_epoch = epochs, 5, 2
    tr.output_folder = None if tmp_path is None else str(tmp_path)
    os.environ["MLAGG_QUIET"] = "1"
    return tr


def test_run_training_epoch_loop(tmp_path):
    tr = _trainer(tmp_path, epochs=3)
    tr.save_every = 1
    tr.run_training()
    lg = tr.logger.my_fantastic_logging
    assert tr.current_epoch == 3 and all(len(lg[k]) == 3 for k in lg)
    # warm-up schedule of timm's CosineLRScheduler(warmup_t=10, warmup_lr_init=1e-4), stepped with the epoch index
    assert np.allclose(lg["lrs"], [1e-4 + e * (5e-4 - 1e-4) / 10 for e in range(3)], rtol=1e-6)
    assert all(np.isfinite(lg["train_losses"])) and all(np.isfinite(lg["val_losses"]))
    assert len(lg["dice_per_class_or_region"][0]) == 4            # 5 classes, background dropped
    assert 0.0 <= lg["mean_fg_dice"][-1] <= 1.0
    assert tr._graph is not None                                   # steps 4+ of epoch 0 replayed the captured graph
    # like the reference: final + best exist, latest (written after epochs 0 and 1) is removed by on_train_end
    assert os.path.isfile(tmp_path / "checkpoint_final.pth") and os.path.isfile(tmp_path / "checkpoint_best.pth")
    assert not os.path.isfile(tmp_path / "checkpoint_latest.pth")
    assert torch.load(tmp_path / "checkpoint_final.pth", weights_only=False)["current_epoch"] == 4   # the reference's + 1


def test_resume_from_checkpoint_latest(tmp_path):
    tr = _trainer(tmp_path, epochs=3)
    tr.save_every = 1
    tr.on_train_end = lambda: None                                 # keep checkpoint_latest (written after epoch 1)
    tr.run_training()
    tr2 = _trainer(tmp_path, epochs=3)
    tr2.initialize()
    tr2.load_checkpoint(str(tmp_path / "checkpoint_latest.pth"))
    assert tr2.current_epoch == 2 and len(tr2.logger.my_fantastic_logging["train_losses"]) == 2
    for (k, a), (_, b) in zip(tr2.optimizer.state_dict()["state"].items(), tr2.optimizer.state_dict()["state"].items()):
        assert float(a["step"]) == 10.0                            # two epochs of five iterations
    tr2.run_training()                                             # epoch 2 only
    lg = tr2.logger.my_fantastic_logging
    assert tr2.current_epoch == 3 and len(lg["lrs"]) == 3
    assert abs(lg["lrs"][2] - (1e-4 + 2 * 4e-4 / 10)) < 1e-9


def test_validation_step_counts_match_a_direct_evaluation():
    tr = _trainer().initialize()
    batch = tr.synthetic_batch(2, seed=3)
    tr.network.eval()
    out = tr.validation_step(batch)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        logits = tr.network(batch["data"].cuda())[0]
    pred = logits.argmax(1).cpu()
    gt = batch["target"][0][:, 0].long()
    # two forward passes of a freshly initialised bf16 network: a handful of near-tie pixels may flip between them
    for c in range(1, 5):
        assert abs(out["tp_hard"][c - 1] - int(((pred == c) & (gt == c)).sum())) <= 8
        assert abs(out["fp_hard"][c - 1] - int(((pred == c) & (gt != c)).sum())) <= 8
        assert abs(out["fn_hard"][c - 1] - int(((pred != c) & (gt == c)).sum())) <= 8
    assert out["tp_hard"].sum() + out["fn_hard"].sum() == int((gt > 0).sum())          # every foreground pixel counted once
    assert np.isfinite(out["loss"])


@pytest.mark.gpu
@pytest.mark.parametrize("K,cl,dtype,tdtype,batch_dice", [(14, True, torch.bfloat16, torch.float32, True),
                                                          (14, False, torch.float32, torch.float32, True),
                                                          (5, True, torch.float32, torch.int64, False),
                                                          (3, False, torch.bfloat16, torch.int64, True),
                                                          (20, True, torch.float32, torch.float32, True)])
def test_fused_dice_ce_matches_the_torch_formulation(K, cl, dtype, tdtype, batch_dice):
    """mlagg_dice_ce_stats_* + the dice arithmetic on the statistics against cross_entropy + soft_dice_loss (the
    restatement of the reference's DC_and_CE_loss) evaluated in float64 on the same logits: value and d loss / d logits,
    NCHW and channels_last heads, float and int64 labels, ragged pixel count."""
    from mlagg_unet_b200.trainer import DeepSupervisionDiceCE, soft_dice_loss
    torch.manual_seed(K)
    Bn, H, W = 3, 37, 29
    logits = (2 * torch.randn(Bn, K, H, W, device="cuda")).to(dtype)
    if cl:
        logits = logits.contiguous(memory_format=torch.channels_last)
    target = torch.randint(0, K, (Bn, 1, H, W), device="cuda").to(tdtype)
    lf = DeepSupervisionDiceCE(1, batch_dice=batch_dice)
    x = logits.clone().requires_grad_()
    loss = lf.one(x, target)
    loss.backward()
    if K == 14 and dtype == torch.bfloat16:       # the padded-head layout: rows of 16, vector row access in the kernels
        wide = torch.zeros(Bn, H, W, 16, device="cuda", dtype=dtype)
        wide[..., :K] = logits.permute(0, 2, 3, 1)
        xv = wide[..., :K].permute(0, 3, 1, 2).detach().requires_grad_()
        lv = lf.one(xv, target)
        gv, = torch.autograd.grad(lv, xv)
        assert abs(float(lv) - float(loss)) < 1e-6 * max(1.0, abs(float(loss)))
        assert torch.equal(gv, x.grad)
    xd = logits.double().requires_grad_()
    ref = torch.nn.functional.cross_entropy(xd, target[:, 0].long()) + soft_dice_loss_f64(xd, target, batch_dice)
    ref.backward()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert abs(float(loss) - float(ref)) < 1e-5 * max(1.0, abs(float(ref)))
    assert x.grad.dtype == dtype and x.grad.stride() == x.stride()
    err = float((x.grad.double() - xd.grad).abs().max() / xd.grad.abs().max())
    assert err < tol, err


def soft_dice_loss_f64(x64, target, batch_dice, smooth=1e-5):
    """trainer.soft_dice_loss without its .float() casts"""
    x = torch.softmax(x64, dim=1)
    onehot = torch.zeros_like(x, dtype=torch.bool).scatter_(1, target.long(), 1)
    x, onehot = x[:, 1:], onehot[:, 1:]
    inter, spred, sgt = (x * onehot).sum((2, 3)), x.sum((2, 3)), onehot.sum((2, 3)).double()
    if batch_dice:
        inter, spred, sgt = inter.sum(0), spred.sum(0), sgt.sum(0)
    return -((2 * inter + smooth) / torch.clip(sgt + spred + smooth, 1e-8)).mean()


def test_side_stream_gradients_equal_the_single_stream_step(monkeypatch):
    """Weight gradients produced on the side stream (tcgen05 dW GEMMs, cuDNN wgrad) and attached by the trainer give the
    same training trajectory as the single-stream step (MLAGG_SIDE_STREAM=0): losses and parameters over four steps,
    three eager and one graph replay, stochastic depth off.  (Float atomics reorder sums between runs: tolerance, not
    equality.)"""
    from mlagg_unet_b200.thirdparty_shims import DropPath
    from mlagg_unet_b200.trainer import SyntheticPlan, nnUNetTrainer_MLAgg_2D_dt_MS
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("MLAGG_SIDE_STREAM", mode)
        torch.manual_seed(0)
        tr = nnUNetTrainer_MLAgg_2D_dt_MS(SyntheticPlan(patch_size=(128, 128), batch_size=2, num_classes=14)).initialize()
        for m in tr.network.modules():
            if isinstance(m, DropPath):
                m.drop_prob = 0.0
        batch = tr.synthetic_batch(2, seed=1)
        losses = [float(tr.train_step(batch)["loss"]) for _ in range(5)]
        res[mode] = (losses, [p.detach().clone() for p in tr.network.parameters()])
        assert tr._graph is not None
    (l1, p1), (l0, p0) = res["1"], res["0"]
    assert max(abs(a - b) for a, b in zip(l1, l0)) < 2e-3, (l1, l0)
    # AdamW moves every element by ~lr per step whatever the gradient's size: compare with the total movement
    num = sum(float((a - b).pow(2).sum()) for a, b in zip(p1, p0)) ** 0.5
    tr0 = nnUNetTrainer_MLAgg_2D_dt_MS(SyntheticPlan(patch_size=(128, 128), batch_size=2, num_classes=14))
    torch.manual_seed(0)
    tr0.initialize()
    den = sum(float((a - b.detach()).pow(2).sum()) for a, b in zip(p0, tr0.network.parameters())) ** 0.5
    assert num < 0.05 * den, (num, den)


@pytest.mark.parametrize("batch_dice", [True, False])
def test_deep_supervision_loss_all_scales_at_once_equals_the_per_scale_sum(batch_dice):
    """DeepSupervisionDiceCE.all_scales (stacked statistics, one set of tiny kernels) == sum_i w_i * one(scale i): value and
    the gradient of every scale's logits."""
    from mlagg_unet_b200.trainer import DeepSupervisionDiceCE
    torch.manual_seed(5)
    K, Bn = 14, 2
    outs = [torch.randn(Bn, K, 64 >> i, 48 >> i, device="cuda").contiguous(memory_format=torch.channels_last).requires_grad_()
            for i in range(5)]
    tgts = [torch.randint(0, K, (Bn, 1, 64 >> i, 48 >> i), device="cuda").float() for i in range(5)]
    lf = DeepSupervisionDiceCE(5, batch_dice=batch_dice)
    l1 = lf(outs, tgts)
    g1 = torch.autograd.grad(l1, outs)
    l2 = sum(w * lf.one(o, t) for w, o, t in zip(lf.weights, outs, tgts))
    g2 = torch.autograd.grad(l2, outs)
    assert abs(float(l1) - float(l2)) < 1e-6 * max(1.0, abs(float(l2)))
    for a, b in zip(g1, g2):
        assert float((a - b).abs().max()) <= 1e-6 * float(b.abs().max()) + 1e-12
