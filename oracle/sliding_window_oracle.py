"""TEST INFRASTRUCTURE ONLY — CPU/torch restatement of the reference's sliding-window predictor.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s baseline legs may import this; the product
(`dinounet_b200/`) never does.

Follows (reference file:line):
  compute_gaussian                        inference/sliding_window_prediction.py:11-32
  compute_steps_for_sliding_window        inference/sliding_window_prediction.py:35-61
  get_sliding_window_slicers              inference/predict_from_raw_data.py:502-536
  maybe_mirror_and_predict                inference/predict_from_raw_data.py:538-553
  predict_sliding_window_return_logits    inference/predict_from_raw_data.py:572-621 (+ driver :680-726)
Third-party piece not under /root/reference: `acvl_utils.cropping_and_padding.padding.pad_nd_image`
(acvl-utils >=0.2.3,<0.3, requirements.txt:2) — its published algorithm is restated in `pad_nd_image` below; the
reference's only call site is predict_from_raw_data.py:703-705 (new_shape = patch size, 'constant', value 0,
return_slicer=True).
Pinned against the REAL reference class (`nnUNetPredictor`, imported with stubbed third-party packages by
`oracle/ref_predictor_loader.py`) in tests/test_sliding_window_cpu.py, and against `scipy.ndimage.gaussian_filter`.
"""
import itertools
from typing import List, Sequence, Tuple

import numpy as np
import torch


def pad_nd_image(image: torch.Tensor, new_shape: Sequence[int]):
    """Centre-pad the trailing dims with zeros up to `new_shape` (never crops); returns (padded, revert slicer)."""
    old = np.array(image.shape)
    new_shape = list(new_shape)
    if len(new_shape) < len(old):
        new_shape = list(old[:len(old) - len(new_shape)]) + new_shape
    new = np.array([max(n, o) for n, o in zip(new_shape, old)])
    diff = new - old
    below, above = diff // 2, diff // 2 + diff % 2
    if diff.any():
        pads = []
        for b, a in zip(below[::-1], above[::-1]):      # F.pad wants the last dim first
            pads += [int(b), int(a)]
        res = torch.nn.functional.pad(image, pads, mode="constant", value=0)
    else:
        res = image
    slicer = tuple(slice(int(b), int(s - a)) for b, a, s in zip(below, above, res.shape))
    return res, slicer


def _gaussian_kernel1d(sigma: float, radius: int) -> np.ndarray:
    """scipy.ndimage._filters._gaussian_kernel1d (order 0)."""
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


def compute_gaussian(tile_size: Sequence[int], sigma_scale: float = 1. / 8, value_scaling_factor: float = 1,
                     dtype=torch.float16, device=torch.device("cpu")) -> torch.Tensor:
    """gaussian_filter(delta at the tile centre, sigma = size/8, truncate 4, constant 0) == outer product of the
    per-axis truncated kernels; / max * scale; cast; zeros replaced by the smallest non-zero entry."""
    axes = []
    for n in tile_size:
        sigma = n * sigma_scale
        radius = int(4.0 * sigma + 0.5)
        k = _gaussian_kernel1d(sigma, radius)
        c = n // 2
        line = np.zeros(n)
        for i in range(n):
            o = c - i                      # correlate1d: out[i] = sum_t k[t + r] * in[i + t]; in is a delta at c
            if -radius <= o <= radius:
                line[i] = k[o + radius]
        axes.append(line)
    g = axes[0]
    for a in axes[1:]:
        g = np.multiply.outer(g, a)
    g = torch.from_numpy(g)
    g = g / torch.max(g) * value_scaling_factor
    g = g.type(dtype).to(device)
    g[g == 0] = torch.min(g[g != 0])
    return g


def compute_steps_for_sliding_window(image_size, tile_size, tile_step_size: float) -> List[List[int]]:
    assert 0 < tile_step_size <= 1
    target = [t * tile_step_size for t in tile_size]
    num_steps = [int(np.ceil((i - k) / j)) + 1 for i, j, k in zip(image_size, target, tile_size)]
    steps = []
    for dim in range(len(tile_size)):
        max_step = image_size[dim] - tile_size[dim]
        actual = max_step / (num_steps[dim] - 1) if num_steps[dim] > 1 else 99999999999
        steps.append([int(np.round(actual * i)) for i in range(num_steps[dim])])
    return steps


def get_sliding_window_slicers(image_size, patch_size, tile_step_size: float):
    """2D patch on a (d, H, W) image: every slice d x every (sx, sy) step, d-major."""
    slicers = []
    if len(patch_size) < len(image_size):
        assert len(patch_size) == len(image_size) - 1
        steps = compute_steps_for_sliding_window(image_size[1:], patch_size, tile_step_size)
        for d in range(image_size[0]):
            for sx in steps[0]:
                for sy in steps[1]:
                    slicers.append((slice(None), d, slice(sx, sx + patch_size[0]), slice(sy, sy + patch_size[1])))
    else:
        steps = compute_steps_for_sliding_window(image_size, patch_size, tile_step_size)
        for sx in steps[0]:
            for sy in steps[1]:
                for sz in steps[2]:
                    slicers.append((slice(None), *[slice(s, s + t) for s, t in zip((sx, sy, sz), patch_size)]))
    return slicers


def maybe_mirror_and_predict(network, x: torch.Tensor, mirror_axes) -> torch.Tensor:
    prediction = network(x)
    if mirror_axes is not None:
        assert max(mirror_axes) <= x.ndim - 3
        combos = [c for i in range(len(mirror_axes)) for c in itertools.combinations([m + 2 for m in mirror_axes], i + 1)]
        for axes in combos:
            prediction += torch.flip(network(torch.flip(x, (*axes,))), (*axes,))
        prediction /= (len(combos) + 1)
    return prediction


def predict_sliding_window_return_logits(network, input_image: torch.Tensor, patch_size: Sequence[int],
                                         num_segmentation_heads: int, tile_step_size: float = 0.5,
                                         use_gaussian: bool = True, mirror_axes=(0, 1),
                                         results_device=torch.device("cpu")) -> torch.Tensor:
    """input_image [c, x, y, z] -> fp16 logits [heads, x, y, z] (padding reverted)."""
    assert input_image.ndim == 4
    with torch.no_grad():
        data, revert = pad_nd_image(input_image, patch_size)
        slicers = get_sliding_window_slicers(data.shape[1:], patch_size, tile_step_size)
        data = data.to(results_device)
        logits = torch.zeros((num_segmentation_heads, *data.shape[1:]), dtype=torch.half, device=results_device)
        n_pred = torch.zeros(data.shape[1:], dtype=torch.half, device=results_device)
        gaussian = compute_gaussian(tuple(patch_size), 1. / 8, 10, device=results_device) if use_gaussian else None
        for sl in slicers:
            workon = data[sl][None]
            prediction = maybe_mirror_and_predict(network, workon, mirror_axes)[0].to(results_device)
            logits[sl] += (prediction * gaussian if use_gaussian else prediction)
            n_pred[sl[1:]] += (gaussian if use_gaussian else 1)
        logits /= n_pred
        if torch.any(torch.isinf(logits)):
            raise RuntimeError("Encountered inf in predicted array")
        return logits[tuple([slice(None), *revert[1:]])]
