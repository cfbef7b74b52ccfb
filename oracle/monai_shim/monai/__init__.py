"""Minimal stand-in for the MONAI names the reference's hot-path files import.

TEST INFRASTRUCTURE ONLY (oracle). MONAI is an un-vendored, un-pinned third-party
dependency of aarchiiive/diff-unet-amos and is not installed in this image. This shim
re-creates, from plain ``torch.nn`` modules, exactly the module tree (sub-module names
``conv`` / ``adn.N`` / ``adn.D`` / ``adn.A`` / ``deconv``) that real MONAI prints in the
reference's own notebook (lab.ipynb:318-330, 361-373), so that the reference files
``models/basic_unet/denoiser.py`` and ``models/basic_unet/pretrained/basic_unet.py``
execute unmodified and produce the reference's checkpoint keys.
It is only ever put on sys.path by oracle/ref_loader.py.
"""
