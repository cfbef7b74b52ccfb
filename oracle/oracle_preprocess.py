"""CPU restatement of the array transforms of the reference's validation pipeline (utils.py:165-181) and of upstream
Diff-UNet's uncertainty-weighted step fusion.

TEST INFRASTRUCTURE ONLY (oracle).  Only tests/, __graft_entry__.smoke() and bench.py's checker legs may import this module.

PARITY UNPINNED: these transforms live in MONAI (module ``monai.transforms``, version unpinned, absent from
/root/reference and from this image) and the fusion lives in upstream Diff-UNet (not vendored either); the reference
holds no golden vectors for them.  They are restated here from the published algorithms:

  * ``foreground_bbox``   monai.transforms.utils.generate_spatial_bounding_box(img, select_fn=is_positive, margin=0)
                          as used by CropForegroundd(source_key="image") at utils.py:171
  * ``spacing_resample``  monai.transforms.Spacing for an axis-aligned affine (utils.py:173-177): output shape from
                          compute_shape_offset(scale_extent=False) = round((n - 1) * s_in / s_out) + 1, output index i
                          maps to input index i * s_out / s_in, bilinear / nearest, border padding
  * ``uncertainty_fuse``  the test-time fusion of upstream Diff-UNet (BraTS test script): per step, uncertainty of the
                          run-averaged output weights the sum of the runs' clamped predictions
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np
import torch


def foreground_bbox(image: torch.Tensor) -> Tuple[Tuple[int, ...], Tuple[int, ...]]:
    """image [C, D, H, W]; (start, end) with end exclusive; empty foreground -> ((0,0,0),(0,0,0))."""
    mask = (image > 0).any(dim=0).numpy()
    if not mask.any():
        return (0, 0, 0), (0, 0, 0)
    start, end = [], []
    for ax in range(3):
        other = tuple(a for a in range(3) if a != ax)
        idx = np.nonzero(mask.any(axis=other))[0]
        start.append(int(idx[0]))
        end.append(int(idx[-1]) + 1)
    return tuple(start), tuple(end)


def resampled_shape(shape: Sequence[int], spacing_in: Sequence[float], pixdim: Sequence[float]) -> Tuple[int, ...]:
    return tuple(int(np.round((n - 1) * float(si) / float(so))) + 1 for n, si, so in zip(shape, spacing_in, pixdim))


def spacing_resample(t: torch.Tensor, spacing_in: Sequence[float], pixdim: Sequence[float], mode: str) -> torch.Tensor:
    """t [C, D, H, W] fp32.  Coordinates in float64, interpolation in fp32: lerp along x, then y, then z, each as
    a + t * (b - a)."""
    C, D, H, W = t.shape
    out_shape = resampled_shape((D, H, W), spacing_in, pixdim)
    coords = []
    for n_out, n_in, si, so in zip(out_shape, (D, H, W), spacing_in, pixdim):
        c = np.arange(n_out, dtype=np.float64) * (float(so) / float(si))
        coords.append(np.clip(c, 0.0, float(n_in - 1)))
    if mode == "nearest":
        iz, iy, ix = [torch.from_numpy(np.rint(c).astype(np.int64)) for c in coords]
        return t[:, iz][:, :, iy][:, :, :, ix].contiguous()
    lo = [np.floor(c).astype(np.int64) for c in coords]
    hi = [np.minimum(l + 1, n - 1) for l, n in zip(lo, (D, H, W))]
    fr = [torch.from_numpy((c - l).astype(np.float32)) for c, l in zip(coords, lo)]
    lo = [torch.from_numpy(v) for v in lo]
    hi = [torch.from_numpy(v) for v in hi]

    def lerp(a, b, w):
        return a + w * (b - a)

    def take(zi, yi):
        plane = t[:, zi][:, :, yi]                      # [C, OD, OH, W]
        return lerp(plane[..., lo[2]], plane[..., hi[2]], fr[2][None, None, None, :])

    fy = fr[1][None, None, :, None]
    fz = fr[0][None, :, None, None]
    c0 = lerp(take(lo[0], lo[1]), take(lo[0], hi[1]), fy)
    c1 = lerp(take(hi[0], lo[1]), take(hi[0], hi[1]), fy)
    return lerp(c0, c1, fz).contiguous()


def uncertainty_fuse(per_step: torch.Tensor) -> torch.Tensor:
    """per_step [R, N, ...] raw model outputs of R runs in loop order -> fused tensor [...] (fp32 torch ops)."""
    R, N = per_step.shape[:2]
    out = torch.zeros_like(per_step[0, 0])
    for k in range(N):
        m = per_step[:, k].sum(0) / R
        p = torch.sigmoid(m).clamp_min(0.001)
        u = -p * torch.log(p)
        w = torch.exp(torch.sigmoid(torch.tensor((k + 1) / N)) * (1 - u))
        out = out + w * per_step[:, k].clamp(-1, 1).sum(0)
    return out
