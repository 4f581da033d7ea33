"""a reference-format checkpoint (dict keys of nnUNetTrainer.save_checkpoint, :1007-1022) built
from the state_dict the reference's own MLLA_Uper produced loads strictly and reproduces the reference logits / argmax
mask; the batched sliding-window predictor equals a tile-by-tile restatement of sliding_window_prediction.py:110-197."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu


def _net():
    from mlagg_unet_b200.mlagg import MLLA_Uper
    return MLLA_Uper(img_size=[64, 64], patch_size=2, in_channels=1, out_channels=5, embed_dim=8, depths=[2, 2, 2, 2],
                     num_heads=[2, 4, 8, 16], mlp_ratio=2, qkv_bias=True, drop_rate=0., dropout_path_rate=0.1,
                     sr_ratio=[16, 8, 4, 2], deep_supervision=True)


def test_reference_format_checkpoint_round_trip(tmp_path):
    from mlagg_unet_b200.trainer import SyntheticPlan, nnUNetTrainer_MLAgg_2D_dt_MS
    g = load_golden("mlla_uper_embed8.pt")
    ref_ckpt = {"network_weights": {"module." + k: v for k, v in g["state"].items()},   # as saved from a DDP run
                "optimizer_state": None, "grad_scaler_state": None, "logging": {}, "_best_ema": None,
                "current_epoch": 7, "init_args": {}, "trainer_name": "nnUNetTrainer_MLAgg_2D_dt_MS",
                "inference_allowed_mirroring_axes": (0, 1)}
    path = str(tmp_path / "checkpoint_final.pth")
    torch.save(ref_ckpt, path)
    tr = nnUNetTrainer_MLAgg_2D_dt_MS(SyntheticPlan(patch_size=(64, 64), batch_size=2, num_classes=5))
    tr.build_network_architecture = staticmethod(lambda *a, **k: _net())
    tr.initialize()
    tr.load_checkpoint(path)
    assert tr.current_epoch == 7
    tr.network.eval()
    with torch.no_grad():
        outs = tr.network(g["input"].cuda())
    assert rel_err(outs[0].cpu(), g["logits0"]) < 1e-4
    assert torch.equal(outs[0].argmax(1).to(torch.uint8).cpu(), g["argmax0"])
    # and back: what we write has the reference's keys and reloads into a fresh trainer with identical weights
    out_path = str(tmp_path / "ours.pth")
    tr.save_checkpoint(out_path)
    ck = torch.load(out_path, weights_only=False)
    assert set(ref_ckpt) == set(ck) and set(ck["network_weights"]) == set(g["state"])
    tr2 = nnUNetTrainer_MLAgg_2D_dt_MS(SyntheticPlan(patch_size=(64, 64), batch_size=2, num_classes=5))
    tr2.build_network_architecture = staticmethod(lambda *a, **k: _net())
    tr2.initialize()
    tr2.load_checkpoint(out_path)
    for (k, a), (_, b) in zip(tr.network.state_dict().items(), tr2.network.state_dict().items()):
        assert torch.equal(a, b), k


def test_reference_optimizer_state_round_trip(tmp_path):
    """A reference checkpoint carries `optimizer_state` of torch.optim.AdamW over ALL network parameters (reference
    nnUNetTrainer_MLAgg_2D_dt_MS.py:137-140), `dummy_tensor` included, with a python-float lr and no fused / capturable
    flags.  It must load, the restored step must still be graph-capturable with the learning rate as a device scalar,
    and what we save must load back into an optimizer built over the reference's parameter list."""
    from mlagg_unet_b200.trainer import SyntheticPlan, nnUNetTrainer_MLAgg_2D_dt_MS
    g = load_golden("mlla_uper_embed8.pt")
    ref_net = _net()
    ref_net.load_state_dict(g["state"], strict=True)
    ref_net.dummy_tensor.requires_grad_(True)                    # as in the reference (:1362)
    ref_params = list(ref_net.parameters())
    ref_opt = torch.optim.AdamW(ref_params, 5e-4, weight_decay=3e-5, eps=1e-4)
    gen = torch.Generator().manual_seed(0)
    for p in ref_params:
        if p is not ref_net.dummy_tensor:                        # never used in forward: no gradient, no state
            p.grad = 1e-3 * torch.randn(p.shape, generator=gen)
    ref_opt.step()
    ck = {"network_weights": ref_net.state_dict(), "optimizer_state": ref_opt.state_dict(), "grad_scaler_state": None,
          "logging": {}, "_best_ema": None, "current_epoch": 3, "init_args": {},
          "trainer_name": "nnUNetTrainer_MLAgg_2D_dt_MS", "inference_allowed_mirroring_axes": (0, 1)}
    path = str(tmp_path / "checkpoint_latest.pth")
    torch.save(ck, path)

    tr = nnUNetTrainer_MLAgg_2D_dt_MS(SyntheticPlan(patch_size=(64, 64), batch_size=2, num_classes=5))
    tr.build_network_architecture = staticmethod(lambda *a, **k: _net())
    tr.initialize()
    tr.load_checkpoint(path)                                     # default load_optimizer=True
    grp = tr.optimizer.param_groups[0]
    assert len(grp["params"]) == len(ref_params)
    assert torch.is_tensor(grp["lr"]) and grp["lr"].is_cuda and grp["capturable"] and grp["fused"]
    own = list(tr.network.parameters())
    for i, p in enumerate(ref_params):
        if p in ref_opt.state:
            assert torch.equal(tr.optimizer.state[own[i]]["exp_avg"].cpu(), ref_opt.state[p]["exp_avg"]), i
            assert float(tr.optimizer.state[own[i]]["step"]) == 1.0
    batch = tr.synthetic_batch(2, device=tr.device)
    losses = [float(tr.train_step(batch)["loss"]) for _ in range(6)]     # 3 eager steps, capture, replays
    assert tr._graph is not None and all(np.isfinite(losses))
    tr.lr_scheduler.step(5)                                      # the scheduler's write reaches the captured step
    assert abs(float(tr.optimizer.param_groups[0]["lr"]) - (1e-4 + 5 * (5e-4 - 1e-4) / 10)) < 1e-9
    out = str(tmp_path / "ours.pth")
    tr.save_checkpoint(out)
    back = torch.load(out, weights_only=False, map_location="cpu")
    ref_opt2 = torch.optim.AdamW(ref_params, 5e-4, weight_decay=3e-5, eps=1e-4)
    ref_opt2.load_state_dict(back["optimizer_state"])            # same group size, same indices
    # (index 0 is `dummy_tensor`: a root-level parameter comes first in .parameters(); it has no state on either side)
    assert ref_params[0] is ref_net.dummy_tensor and len(ref_opt2.state[ref_params[0]]) == 0
    assert float(ref_opt2.state[ref_params[1]]["step"]) == 7.0


def test_sliding_window_matches_tilewise_restatement():
    from mlagg_unet_b200 import inference as inf
    g = load_golden("mlla_uper_embed8.pt")
    net = _net().cuda().eval()
    net.load_state_dict(g["state"], strict=True)
    net.deep_supervision = False
    gen = torch.Generator().manual_seed(21)
    img = torch.randn(1, 2, 100, 90, generator=gen)             # (c, slices, H, W): larger than the 64 x 64 tile
    tile, mirror = (64, 64), (0, 1)
    ours = inf.predict_sliding_window_return_logits(net, img, 5, tile, mirror_axes=mirror, tile_step_size=0.5,
                                                    tiles_per_batch=3, autocast_dtype=None)
    # restatement of sliding_window_prediction.py:110-197 + :82-107, one tile and one network call at a time, fp32
    gauss = torch.from_numpy(inf.compute_gaussian(tile)).cuda()
    steps = inf.compute_steps_for_sliding_window((100, 90), tile, 0.5)
    assert steps == [[0, 18, 36], [0, 26]]
    acc = torch.zeros(5, 2, 100, 90, device="cuda")
    cnt = torch.zeros(2, 100, 90, device="cuda")
    data = img.cuda()
    with torch.no_grad():
        for d in range(2):
            for sx in steps[0]:
                for sy in steps[1]:
                    x = data[:, d, sx:sx + 64, sy:sy + 64][None]
                    p = net(x)
                    p = p + torch.flip(net(torch.flip(x, (2,))), (2,)) + torch.flip(net(torch.flip(x, (3,))), (3,)) \
                        + torch.flip(net(torch.flip(x, (2, 3))), (2, 3))
                    acc[:, d, sx:sx + 64, sy:sy + 64] += p[0] / 4 * gauss
                    cnt[d, sx:sx + 64, sy:sy + 64] += gauss
    ref = acc / cnt
    assert ours.shape == (5, 2, 100, 90)
    assert rel_err(ours.cpu(), ref.cpu()) < 1e-4
    assert (ours.argmax(0) == ref.argmax(0)).float().mean() > 0.999
    # an image smaller than the tile is padded and cropped back
    small = inf.predict_sliding_window_return_logits(net, img[:, :1, :50, :40], 5, tile, autocast_dtype=None)
    assert small.shape == (5, 1, 50, 40) and torch.isfinite(small).all()
    # bf16 autocast path (what the trainer's validation uses): same masks up to near-ties
    half = inf.predict_sliding_window_return_logits(net, img, 5, tile, mirror_axes=mirror)
    assert rel_err(half.cpu(), ref.cpu()) < 5e-2
