"""Pin the oracle (oracle/) against goldens produced by running the UNMODIFIED reference
(oracle/make_golden.py) and, when /root/reference is present, against the live reference."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import oracle_ddim, oracle_model
from oracle.ref_loader import reference_available
from tests.util import GOLDEN, SMALL, load_golden, rel_l2, seeded_image, seeded_noise

FP32_TOL = 2e-5  # same fp32 algorithm, different summation order across CPU kernels


def test_ddim_tables_match_reference():
    gold = json.load(open(os.path.join(GOLDEN, "ddim_tables.json")))
    for n in (10, 25):
        s = oracle_ddim.SpacedSchedule(n)
        g = gold[str(n)]
        assert s.timestep_map == g["timestep_map"]
        for key in ("alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"):
            ref = np.array([float.fromhex(h) for h in g[key]])
            assert np.array_equal(getattr(s, key), ref), key  # bit-exact float64
    assert oracle_ddim.SpacedSchedule(10).timestep_map == [0, 111, 222, 333, 444, 555, 666, 777, 888, 999]


@pytest.mark.parametrize("name,cout,feats", [("C16_default", 16, None), ("C3_default", 3, None),
                                             ("C2_small", 2, SMALL), ("C16_wide", 16, (64, 128, 256, 512, 1024, 64))])
def test_weight_init_matches_reference(name, cout, feats):
    gold = json.load(open(os.path.join(GOLDEN, "weights_fingerprint.json")))[name]
    sd = oracle_model.init_state_dict(1, cout, feats or oracle_model.DEFAULT_FEATURES, seed=0)
    assert len(sd) == gold["n_tensors"]
    assert sum(v.numel() for v in sd.values()) == gold["n_params"]
    fp = oracle_model.state_dict_fingerprint(sd)
    assert list(fp.keys()) == list(gold["fingerprint"].keys())
    for k, v in fp.items():
        assert v == gold["fingerprint"][k], k
    if name == "C16_default":
        assert gold["n_params"] == 38405520 and gold["n_tensors"] == 144


def _oracle_window(cout, S, feats):
    sd = oracle_model.init_state_dict(1, cout, feats, seed=0)
    image = seeded_image((1, 1, S, S, S))
    noise = seeded_noise((1, cout, S, S, S))
    with torch.no_grad():
        emb = oracle_model.encoder_forward(sd, image)
        fn = lambda x, t: oracle_model.denoiser_forward(sd, x, t, image, emb)
        res = oracle_ddim.ddim_sample_window(fn, noise, oracle_ddim.SpacedSchedule(10), collect=True)
        logits999 = fn(noise, torch.tensor([999]))
    return emb, res, logits999


@pytest.mark.parametrize("tag,cout,S,feats", [("S32_C2_small", 2, 32, SMALL), ("S48_C16_small", 16, 48, SMALL),
                                              ("S32_C3_default", 3, 32, oracle_model.DEFAULT_FEATURES)])
def test_window_matches_reference_golden(tag, cout, S, feats):
    g = load_golden(f"window_{tag}.npz")
    sub = int(g["sub"])
    s = slice(None, None, sub)
    emb, res, logits999 = _oracle_window(cout, S, feats)
    assert rel_l2(logits999[:, :, s, s, s], g["logits999"]) < FP32_TOL
    assert rel_l2(res["sample_return"][:, :, s, s, s], g["acc"]) < 1e-4
    assert rel_l2(res["final_x"][:, :, s, s, s], g["final_x"]) < 1e-4
    for i, e in enumerate(emb):
        assert abs(float(e.double().abs().sum()) - g["emb_abs"][i]) <= 1e-5 * g["emb_abs"][i]
    assert rel_l2(emb[0][:, ::8, ::4, ::4, ::4], g["emb0"]) < FP32_TOL
    for k, o in enumerate(res["model_outputs"]):
        assert abs(float(o.double().abs().sum()) - g["step_out_abs"][k]) <= 1e-4 * g["step_out_abs"][k]
    acc = res["sample_return"]
    assert float(acc.min()) >= -10.0 and float(acc.max()) <= 10.0  # sum of 10 clamped x0 (diffusion.py:94-98)


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_oracle_matches_live_reference():
    from oracle.ref_loader import build_reference_model

    cout, S = 2, 32
    m = build_reference_model(1, cout, list(SMALL), seed=0)
    sd = oracle_model.init_state_dict(1, cout, SMALL, seed=0)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k]), k
    image = seeded_image((1, 1, S, S, S))
    noise = seeded_noise((1, cout, S, S, S))
    with torch.no_grad():
        emb_ref = m.embed_model(image)
        emb = oracle_model.encoder_forward(sd, image)
        for a, b in zip(emb, emb_ref):
            assert rel_l2(a, b) < 1e-6
        t = torch.tensor([555])
        assert rel_l2(oracle_model.denoiser_forward(sd, noise, t, image, emb),
                      m.model(noise, t, image=image, embeddings=emb_ref)) < 1e-6
        out = m.sample_diffusion.ddim_sample_loop(m.model, tuple(noise.shape), noise=noise,
                                                  model_kwargs={"image": image, "embeddings": emb_ref})
        res = oracle_ddim.ddim_sample_window(lambda x, tt: oracle_model.denoiser_forward(sd, x, tt, image, emb), noise)
    assert rel_l2(res["sample_return"], sum(out["all_samples"])) < 1e-5
