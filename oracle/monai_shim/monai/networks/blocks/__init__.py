"""Shim of monai.networks.blocks.{Convolution, ADN, UpSample} (3-D, the arguments the reference uses).

Reference call sites: models/basic_unet/denoiser.py:56-58 (Convolution(..., act, norm, dropout, bias,
padding=1)) and :161-170 (UpSample(spatial_dims, in, out, 2, mode="deconv", ...)).
"""
import torch.nn as nn


def _make_act(act):
    name, kwargs = (act, {}) if isinstance(act, str) else act
    if name.lower() != "leakyrelu":
        raise NotImplementedError(f"shim only models LeakyReLU, got {name}")
    return nn.LeakyReLU(**kwargs)


def _make_norm(norm, channels):
    name, kwargs = (norm, {}) if isinstance(norm, str) else norm
    if name.lower() != "instance":
        raise NotImplementedError(f"shim only models instance norm, got {name}")
    return nn.InstanceNorm3d(channels, **kwargs)


class ADN(nn.Sequential):
    """ordering "NDA": norm -> dropout -> activation (lab.ipynb:320-323)."""

    def __init__(self, in_channels, act, norm, dropout):
        super().__init__()
        self.add_module("N", _make_norm(norm, in_channels))
        if dropout is not None:
            self.add_module("D", nn.Dropout(float(dropout)))
        self.add_module("A", _make_act(act))


class Convolution(nn.Sequential):
    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, act=None, norm=None,
                 dropout=None, bias=True, padding=None, **_unused):
        super().__init__()
        assert spatial_dims == 3
        if padding is None:
            padding = kernel_size // 2
        self.add_module("conv", nn.Conv3d(in_channels, out_channels, kernel_size=kernel_size, stride=strides,
                                          padding=padding, bias=bias))
        self.add_module("adn", ADN(out_channels, act, norm, dropout))


class UpSample(nn.Sequential):
    def __init__(self, spatial_dims, in_channels, out_channels, scale_factor=2, mode="deconv", **_unused):
        super().__init__()
        assert spatial_dims == 3 and mode == "deconv"
        self.add_module("deconv", nn.ConvTranspose3d(in_channels, out_channels, kernel_size=scale_factor,
                                                     stride=scale_factor, bias=True))
