"""Seeded synthetic inputs shared by the tests: image seed 1 rand, noise seed 2 randn."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL = (8, 8, 16, 32, 64, 8)


def seeded_image(shape, seed=1):
    torch.manual_seed(seed)
    return torch.rand(*shape)


def seeded_noise(shape, seed=2):
    torch.manual_seed(seed)
    return torch.randn(*shape)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_l2(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
