"""CPU pins of oracle/oracle_preprocess.py (the MONAI transforms themselves are absent: known-answer checks only)."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import oracle_preprocess as op


def test_foreground_bbox_known_answers():
    img = torch.zeros(2, 9, 10, 11)
    assert op.foreground_bbox(img) == ((0, 0, 0), (0, 0, 0))
    img[0, 2, 3, 4] = 1.0
    img[1, 7, 8, 9] = 0.5
    img[0, 5, 1, 10] = -3.0  # not positive: ignored (select_fn = is_positive)
    assert op.foreground_bbox(img) == ((2, 3, 4), (8, 9, 10))


def test_spacing_resample_matches_corner_aligned_interpolation():
    """compute_shape_offset(scale_extent=False): voxel 0 -> voxel 0 and (n-1)*s_in/s_out + 1 output voxels; when the ratio
    makes the last voxels coincide this is exactly F.interpolate(align_corners=True)."""
    torch.manual_seed(0)
    t = torch.rand(2, 9, 13, 5)
    for factor in (2, 4):
        out = op.spacing_resample(t, (float(factor),) * 3, (1.0, 1.0, 1.0), "bilinear")
        shape = tuple((n - 1) * factor + 1 for n in t.shape[1:])
        ref = F.interpolate(t[None], size=shape, mode="trilinear", align_corners=True)[0]
        assert out.shape == ref.shape and float((out - ref).abs().max()) < 1e-6
    same = op.spacing_resample(t, (1.5, 1.5, 2.0), (1.5, 1.5, 2.0), "bilinear")
    assert torch.equal(same, t) and torch.equal(op.spacing_resample(t, (1, 1, 1), (1, 1, 1), "nearest"), t)
    assert op.resampled_shape((37, 45, 52), (0.8, 0.8, 5.0), (1.5, 1.5, 2.0)) == (20, 24, 129)
    near = op.spacing_resample(torch.arange(10.0).view(1, 10, 1, 1), (1.0, 1.0, 1.0), (2.5, 1.0, 1.0), "nearest")
    assert near.flatten().tolist() == [0.0, 2.0, 5.0, 8.0, 9.0]  # rint(2.5) = 2 (half to even), rint(7.5) = 8, 10.0 clamps to the border


def test_uncertainty_fuse_reduces_to_weighted_sum_for_one_run():
    torch.manual_seed(1)
    steps = torch.randn(1, 10, 6, 7)
    out = op.uncertainty_fuse(steps)
    ref = torch.zeros(6, 7)
    for k in range(10):
        p = torch.sigmoid(steps[0, k]).clamp_min(0.001)
        w = torch.exp(torch.sigmoid(torch.tensor((k + 1) / 10)) * (1 + p * torch.log(p)))
        ref += w * steps[0, k].clamp(-1, 1)
    assert float((out - ref).abs().max()) < 1e-5
