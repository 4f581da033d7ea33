"""Sliding-window inference around the hot path -- SURVEY.md 8f-3 / 8f-4.

Mirrors the reference's `nnunetv2/inference/sliding_window_prediction.py`:
  compute_gaussian                      :13-29     (same scipy call, so the importance map is bit-identical)
  compute_steps_for_sliding_window      :32-58
  maybe_mirror_and_predict              :82-107
  predict_sliding_window_return_logits  :110-197   (4-D input (c, x, y, z); for the 2-D configurations x = slices)

B200-first differences: the tiles of an image and their mirrored copies are stacked into ONE batch per forward call
(`tiles_per_batch`, mirror variants included) instead of one tile and up to 4 network calls per step, and the
accumulators are fp32 (the reference accumulates fp16 logits under autocast).  The network call itself is the module
built by `nnUNetTrainer_MLAgg_2D_dt_MS.build_network_architecture`; in eval / no-grad mode the scan forward writes no
state checkpoints.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch


def compute_gaussian(tile_size: Sequence[int], sigma_scale: float = 1. / 8, dtype=np.float32) -> np.ndarray:
    from scipy.ndimage import gaussian_filter
    tmp = np.zeros(tile_size)
    tmp[tuple(i // 2 for i in tile_size)] = 1
    g = gaussian_filter(tmp, [i * sigma_scale for i in tile_size], 0, mode="constant", cval=0)
    g = (g / np.max(g)).astype(dtype)
    g[g == 0] = np.min(g[g != 0])
    return g


def compute_steps_for_sliding_window(image_size: Sequence[int], tile_size: Sequence[int], tile_step_size: float) -> List[List[int]]:
    assert all(i >= j for i, j in zip(image_size, tile_size)), "image size must be as large or larger than patch_size"
    assert 0 < tile_step_size <= 1, "step_size must be larger than 0 and smaller or equal to 1"
    target = [i * tile_step_size for i in tile_size]
    num_steps = [int(np.ceil((i - k) / j)) + 1 for i, j, k in zip(image_size, target, tile_size)]
    steps = []
    for dim in range(len(tile_size)):
        max_step = image_size[dim] - tile_size[dim]
        actual = max_step / (num_steps[dim] - 1) if num_steps[dim] > 1 else 99999999999
        steps.append([int(np.round(actual * i)) for i in range(num_steps[dim])])
    return steps


def _mirror_sets(mirror_axes):
    """The flips maybe_mirror_and_predict applies, as tuples of tensor dims of an (N, C, H, W) batch."""
    if mirror_axes is None:
        return [()]
    assert max(mirror_axes) <= 1, "2-D network: mirror axes 0 / 1"
    sets = [()]
    if 0 in mirror_axes:
        sets.append((2,))
    if 1 in mirror_axes:
        sets.append((3,))
    if 0 in mirror_axes and 1 in mirror_axes:
        sets.append((2, 3))
    return sets


@torch.no_grad()
def predict_sliding_window_return_logits(network, input_image, num_segmentation_heads: int, tile_size: Tuple[int, int],
                                         mirror_axes: Tuple[int, ...] | None = None, tile_step_size: float = 0.5,
                                         use_gaussian: bool = True, precomputed_gaussian: torch.Tensor | None = None,
                                         device: torch.device = torch.device("cuda"), tiles_per_batch: int = 8,
                                         autocast_dtype=torch.bfloat16) -> torch.Tensor:
    """input_image (c, slices, H, W) -> logits (num_segmentation_heads, slices, H, W), fp32 on `device`."""
    network = network.to(device).eval()
    if not isinstance(input_image, torch.Tensor):
        input_image = torch.from_numpy(np.ascontiguousarray(input_image))
    assert input_image.dim() == 4, "input_image must be 4-D (c, x, y, z)"
    c, S, H, W = input_image.shape
    # pad to at least one tile (constant 0, centred like acvl_utils.pad_nd_image)
    ph, pw = max(tile_size[0] - H, 0), max(tile_size[1] - W, 0)
    data = torch.nn.functional.pad(input_image.float(), (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))
    data = data.to(device)
    Hp, Wp = data.shape[2:]
    if use_gaussian:
        g = precomputed_gaussian if precomputed_gaussian is not None else torch.from_numpy(compute_gaussian(tile_size))
        g = g.to(device=device, dtype=torch.float32)
    else:
        g = torch.ones(tile_size, device=device)
    steps = compute_steps_for_sliding_window((Hp, Wp), tile_size, tile_step_size)
    origins = [(d, sx, sy) for d in range(S) for sx in steps[0] for sy in steps[1]]
    logits = torch.zeros(num_segmentation_heads, S, Hp, Wp, device=device)
    weight = torch.zeros(S, Hp, Wp, device=device)
    flips = _mirror_sets(mirror_axes)
    th, tw = tile_size
    enabled = device.type == "cuda" and autocast_dtype is not None
    for i0 in range(0, len(origins), tiles_per_batch):
        chunk = origins[i0:i0 + tiles_per_batch]
        tiles = torch.stack([data[:, d, sx:sx + th, sy:sy + tw] for d, sx, sy in chunk])          # (n, c, th, tw)
        batch = torch.cat([torch.flip(tiles, f) if f else tiles for f in flips])                   # mirrors ride along
        with torch.autocast(device.type, dtype=autocast_dtype, enabled=enabled):
            out = network(batch)
        out = (out[0] if isinstance(out, (list, tuple)) else out).float()
        n = len(chunk)
        pred = sum(torch.flip(out[k * n:(k + 1) * n], f) if f else out[k * n:(k + 1) * n] for k, f in enumerate(flips))
        pred = pred / len(flips)
        for j, (d, sx, sy) in enumerate(chunk):
            logits[:, d, sx:sx + th, sy:sy + tw] += pred[j] * g
            weight[d, sx:sx + th, sy:sy + tw] += g
    logits /= weight
    return logits[:, :, ph // 2:ph // 2 + H, pw // 2:pw // 2 + W]
