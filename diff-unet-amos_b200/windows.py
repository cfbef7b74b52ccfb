"""Sliding-window grid: integer host logic of the window driver (bit-exact part of the path).

Same scan-interval / start-corner / ordering rules as the driver the reference calls at engine.py:173-177
(MONAI sliding_window_inference, constant blend), plus the contiguous per-rank sharding used for
multi-GPU inference (the SequentialDistributedSampler idea, light_training/sampler.py:5-41, applied to windows).
"""
from __future__ import annotations

import itertools
import math
from typing import List, Sequence, Tuple

import numpy as np


def scan_intervals(image_size: Sequence[int], roi: Sequence[int], overlap: float) -> Tuple[int, ...]:
    res = []
    for full, win in zip(image_size, roi):
        if win == full:
            res.append(int(win))
        else:
            step = int(win * (1 - overlap))  # truncation toward zero, e.g. 96 * 0.2 -> 19
            res.append(max(step, 1))
    return tuple(res)


def axis_starts(full: int, win: int, step: int) -> List[int]:
    count = 1
    if step > 0:
        upper = int(math.ceil(float(full) / step))
        for k in range(upper):
            if k * step + win >= full:
                count = k + 1
                break
    # the last window is shifted back so it ends at the border
    return [k * step - max(k * step + win - full, 0) for k in range(count)]


def window_starts(image_size: Sequence[int], roi: Sequence[int], overlap: float) -> np.ndarray:
    """int64 [n_windows, 3] start corners; first spatial axis slowest."""
    steps = scan_intervals(image_size, roi, overlap)
    per_axis = [axis_starts(f, w, s) for f, w, s in zip(image_size, roi, steps)]
    return np.array(list(itertools.product(*per_axis)), dtype=np.int64).reshape(-1, len(per_axis))


def axis_counts(image_size: Sequence[int], roi: Sequence[int], overlap: float) -> List[np.ndarray]:
    """Per-axis coverage counts; the full count map is their outer product (the grid is a Cartesian product)."""
    steps = scan_intervals(image_size, roi, overlap)
    out = []
    for f, w, s in zip(image_size, roi, steps):
        c = np.zeros(f, dtype=np.int32)
        for a in axis_starts(f, w, s):
            c[a:a + w] += 1
        out.append(c)
    return out


def gaussian_importance_map(roi: Sequence[int], sigma_scale: float = 0.125, floor: float = 1e-3):
    """MONAI ``compute_importance_map(roi, mode="gaussian", sigma_scale)`` followed by the clamp sliding_window_inference
    applies (``min_non_zero = max(map.min(), 1e-3)``): a separable product of 1-D gaussians centred on the window,
    sigma_d = sigma_scale * roi_d, built in fp32 in the same op order (monai/data/utils.py, monai/inferers/utils.py >= 1.2).
    Host logic (a [roi] tensor built once per volume); fp32 torch CPU ops."""
    import torch

    imp = None
    for i, n in enumerate(roi):
        x = torch.arange(start=-(n - 1) / 2.0, end=(n - 1) / 2.0 + 1, dtype=torch.float32)
        x = torch.exp(x ** 2 / (-2 * (n * sigma_scale) ** 2))
        imp = x if imp is None else imp.unsqueeze(-1) * x[(None,) * i]
    lo = max(float(imp.min()), floor)
    return torch.clamp(imp, min=lo)


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) share of ``n_items`` for ``rank``; the first ``n_items % world`` ranks get one extra."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)
