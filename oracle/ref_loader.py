"""Import the UNMODIFIED reference hot-path modules from /root/reference.

TEST INFRASTRUCTURE ONLY (oracle). Works only in the build container: /root/reference does not exist on
the GPU box, so nothing reachable from ``pytest -m gpu``, ``smoke()`` or ``bench.py`` may call this at run
time.  It is used by ``oracle/make_golden.py`` (to generate tests/golden/*) and by the ``not gpu`` tests that
pin the restatement in ``oracle/oracle_model.py`` against the real reference when the tree is present.

Recipe:
  1. ``models/__init__.py`` star-imports every model family (Swin-UNETR, MDT, ...) and would need real MONAI,
     timm, ...  ->  register an empty namespace package named ``models`` whose ``__path__`` is the reference's
     ``models/`` directory, so ``models/__init__.py`` is never executed.
  2. put ``oracle/monai_shim`` on sys.path so ``import monai...`` resolves to the shim.
  3. ``from models.diff_unet import DiffUNet`` then runs models/diff_unet.py, models/diffusion/*,
     models/basic_unet/{denoiser,pretrained/basic_unet}.py and guided_diffusion/{gaussian_diffusion,respace,
     resample,nn,losses}.py unmodified.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DUNET_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "monai_shim")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "diff_unet.py"))


def _install():
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "monai" not in sys.modules and _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)
    if REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)
    if "models" not in sys.modules:
        pkg = types.ModuleType("models")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "models")]
        sys.modules["models"] = pkg
    # models/basic_unet/__init__.py star-imports the non-diffusion BasicUNet too; those only need the shim.


def load_reference_diffunet():
    """Return the reference's ``DiffUNet`` class (models/diff_unet.py:9)."""
    _install()
    from models.diff_unet import DiffUNet  # type: ignore

    return DiffUNet


def build_reference_model(in_channels=1, out_channels=16, features=None, seed=0):
    """``torch.manual_seed(seed)`` then construct the reference DiffUNet (eval mode), as SURVEY 8(d) prescribes."""
    import torch

    DiffUNet = load_reference_diffunet()
    torch.manual_seed(seed)
    kw = {} if features is None else {"features": list(features)}
    with contextlib.redirect_stdout(io.StringIO()):  # encoder ctor prints its feature list
        model = DiffUNet(in_channels=in_channels, out_channels=out_channels, **kw)
    return model.eval()


def reference_spaced_diffusion(num_steps=10, timesteps=1000):
    """The reference's own respaced sampler object (models/diffusion/diffusion.py:38-45)."""
    _install()
    from guided_diffusion.gaussian_diffusion import LossType, ModelMeanType, ModelVarType, get_named_beta_schedule
    from guided_diffusion.respace import SpacedDiffusion, space_timesteps

    betas = get_named_beta_schedule("linear", timesteps)
    return SpacedDiffusion(
        use_timesteps=space_timesteps(timesteps, [num_steps]),
        betas=betas,
        model_mean_type=ModelMeanType.START_X,
        model_var_type=ModelVarType.FIXED_LARGE,
        loss_type=LossType.RESCALED_KL,
    )
