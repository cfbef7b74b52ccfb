"""CPU restatement of MONAI ``sliding_window_inference`` (constant blend) as the reference calls it.

TEST INFRASTRUCTURE ONLY (oracle).

The algorithm lives in a THIRD-PARTY dependency that is absent from /root/reference: MONAI
(``monai.inferers.sliding_window_inference``), un-vendored and version-unpinned (inferred >=1.1,<1.5, SURVEY
section 8c).  This file restates its published algorithm (monai/inferers/utils.py
``sliding_window_inference`` / ``_get_scan_interval``; monai/data/utils.py ``dense_patch_slices``,
``compute_importance_map`` with mode="constant") and is anchored on the reference's own call site
engine.py:173-177: ``sliding_window_inference(image, roi, sw_batch_size, model, overlap, pred_type="ddim_sample")``.

PARITY PIN: no golden vectors exist in the reference for this function.  The restatement is
pinned by the known answers of SURVEY Appendix B (window counts / start lists for the reference's cfg values) in
tests/test_oracle_sliding.py; parity for this function is therefore "known-answer pinned", not reference-run pinned.
"""
from __future__ import annotations

import math
from typing import Callable, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def scan_interval(image_size: Sequence[int], roi: Sequence[int], overlap: float) -> Tuple[int, ...]:
    out = []
    for im, r in zip(image_size, roi):
        if r == im:
            out.append(int(r))
        else:
            iv = int(r * (1 - overlap))  # float truncation, e.g. 96*(1-0.8) = 19.199.. -> 19
            out.append(iv if iv > 0 else 1)
    return tuple(out)


def window_starts(image_size: Sequence[int], roi: Sequence[int], interval: Sequence[int]) -> List[List[int]]:
    starts = []
    for im, r, iv in zip(image_size, roi, interval):
        if iv == 0:
            n = 1
        else:
            num = int(math.ceil(float(im) / iv))
            first = next((d for d in range(num) if d * iv + r >= im), None)
            n = first + 1 if first is not None else 1
        dim = []
        for k in range(n):
            s = k * iv
            s -= max(s + r - im, 0)
            dim.append(s)
        starts.append(dim)
    return starts


def window_grid(image_size: Sequence[int], roi: Sequence[int], overlap: float) -> np.ndarray:
    """All window start corners, int64 [n_windows, 3], first spatial dim slowest (meshgrid indexing='ij')."""
    iv = scan_interval(image_size, roi, overlap)
    starts = window_starts(image_size, roi, iv)
    grid = np.asarray([g.flatten() for g in np.meshgrid(*starts, indexing="ij")]).T
    return grid.astype(np.int64)


def count_map(image_size: Sequence[int], roi: Sequence[int], grid: np.ndarray) -> np.ndarray:
    cnt = np.zeros(tuple(image_size), dtype=np.int32)
    for s in grid:
        cnt[s[0]:s[0] + roi[0], s[1]:s[1] + roi[1], s[2]:s[2] + roi[2]] += 1
    return cnt


def importance_map(roi: Sequence[int], mode: str = "constant", sigma_scale: float = 0.125) -> torch.Tensor:
    """monai/data/utils.py ``compute_importance_map`` (constant: ones; gaussian: separable product of 1-D gaussians,
    sigma_d = sigma_scale * roi_d) followed by the clamp of monai/inferers/utils.py (>= 1.2):
    ``min_non_zero = max(map.min(), 1e-3); map = clamp(map, min=min_non_zero)``.  Not exercised by the reference
    (engine.py:173-177 uses the default mode); restated for the gaussian-blend extension, parity unpinned."""
    if mode == "constant":
        return torch.ones(tuple(roi), dtype=torch.float32)
    imp = None
    for i, n in enumerate(roi):
        x = torch.arange(start=-(n - 1) / 2.0, end=(n - 1) / 2.0 + 1, dtype=torch.float32)
        x = torch.exp(x ** 2 / (-2 * (n * sigma_scale) ** 2))
        imp = x if imp is None else imp.unsqueeze(-1) * x[(None,) * i]
    return torch.clamp(imp, min=max(float(imp.min()), 1e-3))


def scale_intensity_range(x: torch.Tensor, a_min=-175.0, a_max=250.0, b_min=0.0, b_max=1.0, clip=True) -> torch.Tensor:
    """monai.transforms.ScaleIntensityRange as the reference configures it (utils.py:167-170)."""
    y = (x - a_min) / (a_max - a_min)
    y = y * (b_max - b_min) + b_min
    return torch.clamp(y, b_min, b_max) if clip else y


def sliding_window_inference(inputs: torch.Tensor, roi_size: Sequence[int], sw_batch_size: int,
                             predictor: Callable[..., torch.Tensor], overlap: float = 0.25, mode: str = "constant",
                             sigma_scale: float = 0.125, **kwargs) -> torch.Tensor:
    """Sliding window, constant blend (the reference's) or gaussian blend.  ``predictor(window_batch, window_indices=..., **kwargs)`` -> [b, C, *roi].

    ``window_indices`` (flat indices n*num_win + w of the windows in the batch) is an oracle-side addition so the
    predictor can pick the explicit per-window noise; MONAI itself forwards only ``**kwargs``.
    """
    batch, _, *img = inputs.shape
    roi = tuple(int(r) for r in roi_size)
    size = tuple(max(i, r) for i, r in zip(img, roi))
    pad = []
    for k in range(len(inputs.shape) - 1, 1, -1):
        diff = max(roi[k - 2] - inputs.shape[k], 0)
        half = diff // 2
        pad.extend([half, diff - half])
    if any(pad):
        inputs = F.pad(inputs, pad=pad, mode="constant", value=0.0)
    grid = window_grid(size, roi, overlap)
    num_win = len(grid)
    total = num_win * batch
    out = None
    cnt = torch.zeros((1, 1) + size, dtype=torch.float32)
    imp = importance_map(roi, mode, sigma_scale)
    for g in range(0, total, sw_batch_size):
        idxs = list(range(g, min(g + sw_batch_size, total)))
        crops = []
        for idx in idxs:
            n, s = idx // num_win, grid[idx % num_win]
            crops.append(inputs[n:n + 1, :, s[0]:s[0] + roi[0], s[1]:s[1] + roi[1], s[2]:s[2] + roi[2]])
        pred = predictor(torch.cat(crops), window_indices=idxs, **kwargs)
        if out is None:
            out = torch.zeros((batch, pred.shape[1]) + size, dtype=torch.float32)
        for j, idx in enumerate(idxs):
            n, s = idx // num_win, grid[idx % num_win]
            if mode == "constant":
                out[n, :, s[0]:s[0] + roi[0], s[1]:s[1] + roi[1], s[2]:s[2] + roi[2]] += pred[j]
            else:
                out[n, :, s[0]:s[0] + roi[0], s[1]:s[1] + roi[1], s[2]:s[2] + roi[2]] += imp * pred[j]
            if n == 0:
                cnt[0, 0, s[0]:s[0] + roi[0], s[1]:s[1] + roi[1], s[2]:s[2] + roi[2]] += imp
    out = out / cnt
    if any(pad):  # crop the padding back off (pad list is ordered last dim first)
        sl = [slice(None), slice(None)]
        for d in range(3):
            lo = pad[(2 - d) * 2]
            sl.append(slice(lo, lo + img[d]))
        out = out[tuple(sl)]
    return out


def engine_infer_labels(stitched: torch.Tensor) -> torch.Tensor:
    """engine.py:179-180: ``(sigmoid(outputs) > 0.5).float()``."""
    return (torch.sigmoid(stitched) > 0.5).float()
