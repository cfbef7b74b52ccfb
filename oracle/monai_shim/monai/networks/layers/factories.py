"""Shim of monai.networks.layers.factories.{Conv, Pool}: ``Conv["conv", 3]`` / ``Pool["MAX", 3]`` lookups
(reference: models/basic_unet/denoiser.py:100, :282)."""
import torch.nn as nn


class _Factory:
    def __init__(self, table):
        self._table = table

    def __getitem__(self, key):
        name, dims = key
        return self._table[(name.lower(), int(dims))]


Conv = _Factory({("conv", 3): nn.Conv3d, ("convtrans", 3): nn.ConvTranspose3d})
Pool = _Factory({("max", 3): nn.MaxPool3d, ("avg", 3): nn.AvgPool3d})
