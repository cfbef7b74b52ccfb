"""NOT a test: the "library bar" of SURVEY 8d -- the reference's network (oracle port of its PyTorch code) run EAGERLY
with torch/cuDNN on the B200, one 96^3 window, C=16, DDIM-10, in fp32 (TF32 off / on) and bf16 autocast.  Lives under tests/
because only tests/ may import oracle/.  Prints patches/s next to the B200-native path for the same window.

    python tests/library_bar.py [--batch B]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import diff_unet_amos_b200 as pkg
from oracle import oracle_ddim, oracle_model

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
a = ap.parse_args()
S, C, B = 96, 16, a.batch
dev = "cuda"
sd = {k: v.to(dev) for k, v in oracle_model.init_state_dict(1, C, oracle_model.DEFAULT_FEATURES, seed=0).items()}
torch.manual_seed(1)
image = torch.rand(B, 1, S, S, S, device=dev)
torch.manual_seed(2)
noise = torch.randn(B, C, S, S, S, device=dev)
sched = oracle_ddim.SpacedSchedule(10)


def window(autocast_dtype=None):
    ctx = torch.autocast("cuda", dtype=autocast_dtype) if autocast_dtype else torch.autocast("cuda", enabled=False)
    with torch.no_grad(), ctx:
        emb = oracle_model.encoder_forward(sd, image)
        x, acc = noise, torch.zeros_like(noise)
        for i in reversed(range(10)):
            t = torch.full((B,), sched.timestep_map[i], dtype=torch.int64, device=dev)
            out = oracle_model.denoiser_forward(sd, x, t, image, emb).float()
            x, x0 = oracle_ddim.ddim_step(sched, i, x, out)
            acc = acc + x0
    return acc


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


rows = []
for name, tf32, dt in (("eager torch fp32 (TF32 off)", False, None), ("eager torch fp32 (TF32 on)", True, None),
                       ("eager torch bf16 autocast", True, torch.bfloat16)):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True
    s = timeit(lambda: window(dt))
    rows.append((name, s))
for prec in ("fp16", "bf16"):
    m = pkg.DiffUNetB200(in_channels=1, out_channels=C, image_size=S, spatial_size=S, batch_max=B, precision=prec).to(dev).eval()
    m.load_state_dict({k: v for k, v in sd.items()})
    s = timeit(lambda: m(image=image, pred_type="ddim_sample", noise=noise), reps=5)
    rows.append((f"libdunet_b200 ({prec}, this repo)", s))
    del m
m32 = pkg.DiffUNetB200(in_channels=1, out_channels=C, image_size=S, spatial_size=S, batch_max=B, precision="fp32x3").to(dev).eval()
m32.load_state_dict({k: v for k, v in sd.items()})
s = timeit(lambda: m32(image=image, pred_type="ddim_sample", noise=noise), reps=3)
rows.append(("libdunet_b200 (fp32x3, this repo)", s))
print(f"one DDIM-10 call, {B} window(s) of 96^3, C=16, default features, {torch.cuda.get_device_name(0)}")
for name, s in rows:
    print(f"  {name:36s} {1e3 * s:9.2f} ms per call   {B / s:8.2f} patches/s")
