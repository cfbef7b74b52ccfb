"""Importable alias of the product package, whose directory is named ``diff-unet-amos_b200`` (not a valid Python
identifier).  ``import diff_unet_amos_b200`` resolves sub-modules from that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "diff-unet-amos_b200")
__path__.insert(0, _real)

from .api import *  # noqa: E402,F401,F403
from .api import __all__  # noqa: E402,F401
