"""Respaced DDIM schedule (host side, float64 numpy) for the B200 path.

Product code: computes what the reference's ``SpacedDiffusion(space_timesteps(T, [N]), linear betas)`` holds
(guided_diffusion/respace.py:7-86, gaussian_diffusion.py:18-35,131-147) and hands the fp32-cast tables to the
C ABI (``dunet_plan_set_schedule``).  The kernels use them exactly as gaussian_diffusion.py:345-349,566-584 do.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np


def kept_timesteps(total: int, n: int) -> List[int]:
    """Evenly strided subset of range(total) with both ends kept (single-section ``space_timesteps``)."""
    if n > total:
        raise ValueError(f"cannot divide section of {total} steps into {n}")
    if n <= 1:
        return [0]
    stride = (total - 1) / (n - 1)
    pos, out = 0.0, []
    for _ in range(n):
        out.append(round(pos))
        pos += stride
    return sorted(set(out))


@dataclass
class DdimSchedule:
    timestep_map: List[int]
    alphas_cumprod: np.ndarray
    alphas_cumprod_prev: np.ndarray
    sqrt_recip_alphas_cumprod: np.ndarray
    sqrt_recipm1_alphas_cumprod: np.ndarray

    @property
    def num_timesteps(self) -> int:
        return len(self.timestep_map)

    @staticmethod
    def build(num_steps: int = 10, train_timesteps: int = 1000) -> "DdimSchedule":
        scale = 1000 / train_timesteps
        betas = np.linspace(scale * 0.0001, scale * 0.02, train_timesteps, dtype=np.float64)
        base = np.cumprod(1.0 - betas, axis=0)
        keep = kept_timesteps(train_timesteps, num_steps)
        # respacing: beta'_k = 1 - abar[t_k] / abar[t_{k-1}]  ->  cumprod(1 - beta') (kept in the reference's op order)
        new_betas, last = [], 1.0
        for t in keep:
            new_betas.append(1 - base[t] / last)
            last = base[t]
        ac = np.cumprod(1.0 - np.array(new_betas, dtype=np.float64), axis=0)
        return DdimSchedule(
            timestep_map=keep,
            alphas_cumprod=ac,
            alphas_cumprod_prev=np.append(1.0, ac[:-1]),
            sqrt_recip_alphas_cumprod=np.sqrt(1.0 / ac),
            sqrt_recipm1_alphas_cumprod=np.sqrt(1.0 / ac - 1),
        )

    def closed_form(self):
        """x_prev = a_i * x0 + b_i * x_t  (eta = 0); documentation / tests only."""
        a = np.sqrt(self.alphas_cumprod_prev) - np.sqrt(1 - self.alphas_cumprod_prev) / self.sqrt_recipm1_alphas_cumprod
        b = np.sqrt(1 - self.alphas_cumprod_prev) * self.sqrt_recip_alphas_cumprod / self.sqrt_recipm1_alphas_cumprod
        return a, b
