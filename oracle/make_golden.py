"""Generate tests/golden/* by RUNNING THE UNMODIFIED REFERENCE (build container only).

TEST INFRASTRUCTURE ONLY (oracle).  Usage:  python oracle/make_golden.py
Needs /root/reference; the outputs are committed so the GPU box (which has no reference tree) can check
against them.  Seeds follow SURVEY section 8(d): weights manual_seed(0), image manual_seed(1) + rand,
noise manual_seed(2) + randn.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle_sliding  # noqa: E402
from oracle.oracle_model import state_dict_fingerprint  # noqa: E402
from oracle.ref_loader import build_reference_model, reference_spaced_diffusion  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def seeded_image(shape, seed=1):
    torch.manual_seed(seed)
    return torch.rand(*shape)


def seeded_noise(shape, seed=2):
    torch.manual_seed(seed)
    return torch.randn(*shape)


def tables():
    out = {}
    for n in (10, 25):
        d = reference_spaced_diffusion(n)
        out[str(n)] = {
            "timestep_map": [int(t) for t in d.timestep_map],
            "alphas_cumprod": [float(v).hex() for v in d.alphas_cumprod],
            "alphas_cumprod_prev": [float(v).hex() for v in d.alphas_cumprod_prev],
            "sqrt_recip_alphas_cumprod": [float(v).hex() for v in d.sqrt_recip_alphas_cumprod],
            "sqrt_recipm1_alphas_cumprod": [float(v).hex() for v in d.sqrt_recipm1_alphas_cumprod],
        }
    with open(os.path.join(GOLD, "ddim_tables.json"), "w") as f:
        json.dump(out, f, indent=1)


def fingerprints():
    out = {}
    for name, (cout, feats) in {"C16_default": (16, None), "C3_default": (3, None),
                                "C2_small": (2, [8, 8, 16, 32, 64, 8]),
                                "C16_wide": (16, [64, 128, 256, 512, 1024, 64])}.items():
        m = build_reference_model(1, cout, feats, seed=0)
        sd = m.state_dict()
        out[name] = {"n_tensors": len(sd), "n_params": int(sum(v.numel() for v in sd.values())),
                     "fingerprint": state_dict_fingerprint(sd)}
    with open(os.path.join(GOLD, "weights_fingerprint.json"), "w") as f:
        json.dump(out, f)


def window_case(tag, cout, S, feats=None, sub=1):
    """One window through the reference's own API: Diffusion.forward(pred_type='ddim_sample') and, with explicit
    noise, SpacedDiffusion.ddim_sample_loop(model.model, ..., noise=...)."""
    torch.set_num_threads(os.cpu_count())
    m = build_reference_model(1, cout, feats, seed=0)
    image = seeded_image((1, 1, S, S, S), 1)
    noise = seeded_noise((1, cout, S, S, S), 2)
    with torch.no_grad():
        torch.manual_seed(2)  # first RNG draw inside ddim_sample_loop_progressive is randn(shape) == noise
        acc_api = m(image=image, pred_type="ddim_sample")
        emb = m.embed_model(image)
        out = m.sample_diffusion.ddim_sample_loop(m.model, (1, cout, S, S, S), noise=noise,
                                                  model_kwargs={"image": image, "embeddings": emb})
        acc = sum(s for s in out["all_samples"])
        assert torch.equal(acc, acc_api), "explicit-noise path differs from the forward(pred_type) path"
        # single denoiser call at t = 999 on the raw noise (the model(x, t, image=, embeddings=) seam)
        logits999 = m.model(noise, torch.tensor([999]), image=image, embeddings=emb)
    s = slice(None, None, sub)
    np.savez_compressed(
        os.path.join(GOLD, f"window_{tag}.npz"),
        cout=cout, S=S, sub=sub, features=np.array(feats if feats else [64, 64, 128, 256, 512, 64]),
        acc=acc[:, :, s, s, s].numpy(),
        final_x=out["sample"][:, :, s, s, s].numpy(),
        logits999=logits999[:, :, s, s, s].numpy(),
        step_out_sum=np.array([float(o.double().sum()) for o in out["all_model_outputs"]]),
        step_out_abs=np.array([float(o.double().abs().sum()) for o in out["all_model_outputs"]]),
        emb_sum=np.array([float(e.double().sum()) for e in emb]),
        emb_abs=np.array([float(e.double().abs().sum()) for e in emb]),
        emb0=emb[0][:, ::8, ::4, ::4, ::4].numpy(),
        emb4=emb[4].numpy() if emb[4].numel() < 70000 else emb[4][:, ::8].numpy(),
    )
    print(tag, "acc range", float(acc.min()), float(acc.max()))


def volume_case():
    """Whole-volume: restated MONAI driver (oracle_sliding) around the REFERENCE model as predictor.
    48x48x40 volume, roi 32, overlap 0.25 -> 8 windows (SURVEY Appendix B small check)."""
    cout, roi = 2, (32, 32, 32)
    m = build_reference_model(1, cout, [8, 8, 16, 32, 64, 8], seed=0)
    image = seeded_image((1, 1, 48, 48, 40), 1)
    nwin = len(oracle_sliding.window_grid((48, 48, 40), roi, 0.25))
    noise = seeded_noise((nwin, cout) + roi, 2)

    def predictor(batch, window_indices=None, pred_type=None):
        res = []
        for j, w in enumerate(window_indices):
            img = batch[j:j + 1]
            emb = m.embed_model(img)
            out = m.sample_diffusion.ddim_sample_loop(m.model, (1, cout) + roi, noise=noise[w:w + 1],
                                                      model_kwargs={"image": img, "embeddings": emb})
            res.append(sum(out["all_samples"]))
        return torch.cat(res)

    with torch.no_grad():
        stitched = oracle_sliding.sliding_window_inference(image, roi, 4, predictor, 0.25, pred_type="ddim_sample")
    np.savez_compressed(os.path.join(GOLD, "volume_48x48x40_C2_small.npz"), stitched=stitched.numpy(), nwin=nwin,
                        labels=oracle_sliding.engine_infer_labels(stitched).numpy().astype(np.uint8))
    print("volume", stitched.shape, nwin)


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    tables()
    fingerprints()
    window_case("S32_C3_default", 3, 32)
    window_case("S32_C16_default", 16, 32, sub=2)
    window_case("S32_C2_small", 2, 32, [8, 8, 16, 32, 64, 8])
    window_case("S48_C16_small", 16, 48, [8, 8, 16, 32, 64, 8], sub=2)
    volume_case()
