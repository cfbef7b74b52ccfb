"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's Diff-UNet DDIM sliding-window inference path, used as the checker in
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.  The product package
(diff-unet-amos_b200/) never imports anything from here and fails loudly without its CUDA library.

  ref_loader.py      imports the UNMODIFIED reference from /root/reference through a small MONAI shim
                     (build container only; used to generate tests/golden/* and to pin the restatement)
  oracle_model.py    networks (encoder, denoiser, time embedding) as functional torch fp32
  oracle_ddim.py     respaced DDIM schedule tables + sampling loop + sum of x0
  oracle_sliding.py  MONAI sliding_window_inference (constant blend) restated; known-answer pinned
  make_golden.py     generator of tests/golden/* (runs the reference; committed with its outputs)

The reference is pure Python/PyTorch (no native sources), so there is no C restatement and no oracle/_ref
binary: the checker is the torch-fp32 port above, pinned against reference-run goldens.
"""
