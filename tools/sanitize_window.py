"""One small window (32^3, default features, 1 DDIM step, 2 windows) for compute-sanitizer:
    compute-sanitizer --tool initcheck  python tools/sanitize_window.py
    compute-sanitizer --tool memcheck   python tools/sanitize_window.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import diff_unet_amos_b200 as pkg

S, C = int(os.environ.get("S", "32")), 3
torch.manual_seed(0)
m = pkg.DiffUNetB200(in_channels=1, out_channels=C, image_size=S, spatial_size=S, batch_max=2, num_steps=int(os.environ.get("STEPS", "1"))).cuda().eval()
image = torch.rand(2, 1, S, S, S, device="cuda")
noise = torch.randn(2, C, S, S, S, device="cuda")
with torch.no_grad():
    out = m(image=image, pred_type="ddim_sample", noise=noise)
torch.cuda.synchronize()
print("done", float(out.abs().max()))
