"""CPU restatement of the reference's DDIM-respaced sampler (schedule tables + loop).

TEST INFRASTRUCTURE ONLY (oracle): imported by tests/, smoke() and bench.py's CPU legs, never by the product.

Follows, line by line:
  * linear beta schedule                guided_diffusion/gaussian_diffusion.py:18-35
  * cumulative tables (float64)         guided_diffusion/gaussian_diffusion.py:131-147
  * respacing                           guided_diffusion/respace.py:7-60 (space_timesteps), :72-86 (new betas)
  * timestep remap                      guided_diffusion/respace.py:123-129
  * one DDIM step, eta = 0              guided_diffusion/gaussian_diffusion.py:537-586 with p_mean_variance
                                        :231-326 (START_X, clip_denoised) and _predict_eps_from_xstart :345-349
  * table lookup cast float64 -> fp32   guided_diffusion/gaussian_diffusion.py:904-917
  * loop high -> low + sum of x0        gaussian_diffusion.py:667-716, models/diffusion/diffusion.py:86-102

Parity pin: tests/golden/ddim_tables.json (written by oracle/make_golden.py from the reference objects).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import numpy as np
import torch


def linear_betas(num_timesteps: int = 1000) -> np.ndarray:
    scale = 1000 / num_timesteps
    return np.linspace(scale * 0.0001, scale * 0.02, num_timesteps, dtype=np.float64)


def space_timesteps(num_timesteps: int, section_counts: List[int]) -> List[int]:
    """respace.py:7-60 for the list form (the reference passes ``[10]``)."""
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start = 0
    steps: List[int] = []
    for i, count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        cur = 0.0
        for _ in range(count):
            steps.append(start + round(cur))
            cur += stride
        start += size
    return sorted(set(steps))


class SpacedSchedule:
    """The float64 tables of ``SpacedDiffusion(space_timesteps(T, [N]), linear betas)``."""

    def __init__(self, num_steps: int = 10, timesteps: int = 1000):
        base_ac = np.cumprod(1.0 - linear_betas(timesteps), axis=0)
        use = set(space_timesteps(timesteps, [num_steps]))
        last = 1.0
        new_betas, tmap = [], []
        for i, ac in enumerate(base_ac):
            if i in use:
                new_betas.append(1 - ac / last)
                last = ac
                tmap.append(i)
        betas = np.array(new_betas, dtype=np.float64)
        self.timestep_map: List[int] = tmap
        self.num_timesteps = len(tmap)
        self.alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)

    def f32(self, arr: np.ndarray, i: int) -> torch.Tensor:
        """_extract_into_tensor: index the float64 table, then ``.float()`` (gaussian_diffusion.py:914)."""
        return torch.from_numpy(arr)[i].float()


def ddim_step(sched: SpacedSchedule, i: int, x_t: torch.Tensor, model_output: torch.Tensor):
    """One eta=0 DDIM update at loop index ``i`` -> (x_prev, pred_xstart), in the reference's op order."""
    x0 = model_output.clamp(-1, 1)                                              # :292-297
    eps = (sched.f32(sched.sqrt_recip_alphas_cumprod, i) * x_t - x0) / sched.f32(sched.sqrt_recipm1_alphas_cumprod, i)
    ab_prev = sched.f32(sched.alphas_cumprod_prev, i)
    # sigma = eta * ... = 0 ; noise term multiplied by 0                         :570-584
    x_prev = x0 * torch.sqrt(ab_prev) + torch.sqrt(1 - ab_prev - 0.0) * eps
    return x_prev, x0


def ddim_sample_window(model_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], noise: torch.Tensor,
                       sched: Optional[SpacedSchedule] = None, collect: bool = False) -> Dict[str, object]:
    """``Diffusion.ddim_sample`` for one window batch given explicit ``noise``.

    ``model_fn(x_t, t_original)`` is the denoiser with image/embeddings bound.  Returns the SUM over steps
    of the clamped x0 predictions (models/diffusion/diffusion.py:94-98) and, optionally, per-step outputs.
    """
    sched = sched or SpacedSchedule(10)
    x = noise
    acc = torch.zeros_like(noise)
    outs, x0s = [], []
    for i in reversed(range(sched.num_timesteps)):
        t = torch.full((noise.shape[0],), sched.timestep_map[i], dtype=torch.int64)
        out = model_fn(x, t)
        x, x0 = ddim_step(sched, i, x, out)
        acc = acc + x0
        if collect:
            outs.append(out)
            x0s.append(x0)
    return {"sample_return": acc, "final_x": x, "model_outputs": outs, "pred_xstarts": x0s}
