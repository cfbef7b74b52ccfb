"""Known-answer pins for the restated MONAI sliding-window driver + the
reference-model-driven whole-volume golden."""
import numpy as np
import torch

from oracle import oracle_ddim, oracle_model, oracle_sliding
from tests.util import SMALL, load_golden, rel_l2, seeded_image, seeded_noise


def test_known_window_counts():
    g = oracle_sliding.window_grid((512, 512, 160), (96, 96, 96), 0.25)
    assert len(g) == 98
    assert sorted(set(g[:, 0])) == [0, 72, 144, 216, 288, 360, 416]
    assert sorted(set(g[:, 2])) == [0, 64]
    assert g[0].tolist() == [0, 0, 0] and g[1].tolist() == [0, 0, 64] and g[2].tolist() == [0, 72, 0]  # dim 0 slowest
    assert oracle_sliding.scan_interval((512, 512, 160), (96, 96, 96), 0.8) == (19, 19, 19)
    assert len(oracle_sliding.window_grid((512, 512, 160), (96, 96, 96), 0.8)) == 2645
    assert len(oracle_sliding.window_grid((512, 512, 160), (96, 96, 96), 0.1)) == 72
    assert len(oracle_sliding.window_grid((512, 512, 160), (96, 96, 96), 0.5)) == 300
    assert len(oracle_sliding.window_grid((512, 512, 160), (128, 128, 128), 0.25)) == 50
    g = oracle_sliding.window_grid((48, 48, 40), (32, 32, 32), 0.25)
    assert len(g) == 8
    cnt = oracle_sliding.count_map((48, 48, 40), (32, 32, 32), g)
    assert set(np.unique(cnt).tolist()) == {1, 2, 4, 8}
    # roi == image along a dim -> interval = roi, one window
    assert oracle_sliding.scan_interval((96, 200, 96), (96, 96, 96), 0.25) == (96, 72, 96)


def test_small_image_is_padded_and_cropped():
    img = torch.arange(20 * 40 * 40, dtype=torch.float32).reshape(1, 1, 20, 40, 40)
    out = oracle_sliding.sliding_window_inference(img, (32, 32, 32), 2, lambda b, window_indices=None: b * 2.0, 0.25)
    assert out.shape == img.shape
    assert torch.equal(out, img * 2.0)


def test_volume_matches_reference_driven_golden():
    g = load_golden("volume_48x48x40_C2_small.npz")
    cout, roi = 2, (32, 32, 32)
    sd = oracle_model.init_state_dict(1, cout, SMALL, seed=0)
    image = seeded_image((1, 1, 48, 48, 40))
    nwin = int(g["nwin"])
    noise = seeded_noise((nwin, cout) + roi)
    sched = oracle_ddim.SpacedSchedule(10)

    def predictor(batch, window_indices=None, pred_type=None):
        assert pred_type == "ddim_sample"
        res = []
        for j, w in enumerate(window_indices):
            img = batch[j:j + 1]
            emb = oracle_model.encoder_forward(sd, img)
            fn = lambda x, t: oracle_model.denoiser_forward(sd, x, t, img, emb)
            res.append(oracle_ddim.ddim_sample_window(fn, noise[w:w + 1], sched)["sample_return"])
        return torch.cat(res)

    with torch.no_grad():
        out = oracle_sliding.sliding_window_inference(image, roi, 4, predictor, 0.25, pred_type="ddim_sample")
    assert rel_l2(out, g["stitched"]) < 1e-4
    lab = oracle_sliding.engine_infer_labels(out).numpy().astype(np.uint8)
    assert (lab == g["labels"]).mean() > 0.9999


def test_gaussian_importance_map_known_answers():
    """MONAI gaussian blend (extension): centre weight exp(-0.5 (0.5/sigma)^2)^3, corner clamped to 1e-3,
    symmetric; a constant predictor is reproduced exactly up to rounding; product host logic == oracle."""
    import diff_unet_amos_b200 as pkg

    imp = oracle_sliding.importance_map((96, 96, 96), "gaussian", 0.125)
    c = float(np.exp(np.float32(0.5) ** 2 / np.float32(-2 * 12.0 ** 2)))
    assert abs(float(imp[47, 47, 47]) - c ** 3) < 1e-6
    assert float(imp.min()) == np.float32(1e-3) and float(imp[0, 0, 0]) == np.float32(1e-3)
    assert torch.equal(imp, imp.flip(0)) and torch.allclose(imp, imp.permute(1, 2, 0), rtol=1e-6)  # (gz*gy)*gx: order matters bitwise
    assert torch.equal(pkg.gaussian_importance_map((96, 96, 96), 0.125), imp)
    assert torch.equal(pkg.gaussian_importance_map((32, 48, 64), 0.125), oracle_sliding.importance_map((32, 48, 64), "gaussian"))
    img = torch.rand(1, 1, 40, 48, 36)
    out = oracle_sliding.sliding_window_inference(img, (32, 32, 32), 2, lambda b, window_indices=None: torch.full_like(b, 3.0),
                                                  0.25, mode="gaussian")
    assert torch.allclose(out, torch.full_like(out, 3.0), rtol=1e-6)


def test_scale_intensity_range_known_answers():
    x = torch.tensor([-1000.0, -175.0, 37.5, 250.0, 3000.0])
    y = oracle_sliding.scale_intensity_range(x)
    assert y.tolist() == [0.0, 0.0, 0.5, 1.0, 1.0]
