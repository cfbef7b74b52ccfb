"""NOT a pytest test: where does the reduced-precision error of a 96^3 window come from?  (run by hand on a B200)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import diff_unet_amos_b200 as pkg
from oracle import oracle_ddim, oracle_model
from tests.util import seeded_image, seeded_noise, rel_l2

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
C, S = 16, 96


def build(prec, flags=0):
    torch.manual_seed(0)
    return pkg.DiffUNetB200(in_channels=1, out_channels=C, image_size=S, spatial_size=S, batch_max=1, precision=prec, debug_flags=flags).cuda().eval()


def oracle(sd, image, noise):
    sched = oracle_ddim.SpacedSchedule(10)
    e = oracle_model.encoder_forward(sd, image)
    x, ref = noise, torch.zeros_like(noise)
    for i in reversed(range(10)):
        t = torch.full((1,), sched.timestep_map[i], dtype=torch.int64, device="cuda")
        x, x0 = oracle_ddim.ddim_step(sched, i, x, oracle_model.denoiser_forward(sd, x, t, image, e))
        ref = ref + x0
    return ref, e


def report(tag, acc, ref):
    err = rel_l2(acc.cpu(), ref.cpu())
    sign = ((acc > 0) == (ref > 0)).float().mean().item()
    am = (acc.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"{tag:40s} rel-l2 {err:.3e}  sign {sign:.5f}  argmax {am:.5f}", flush=True)


with torch.no_grad():
    for seed in (1, 11):
        image, noise = seeded_image((1, 1, S, S, S), seed).cuda(), seeded_noise((1, C, S, S, S), seed + 1).cuda()
        m16 = build("fp16")
        sd = {k: v.detach() for k, v in m16.state_dict().items()}
        ref, emb_ref = oracle(sd, image, noise)
        report(f"seed {seed} fp16", m16(image=image, pred_type="ddim_sample", noise=noise), ref)
        out = m16.sample_diffusion.ddim_sample_loop(m16.model, noise.shape, noise=noise, model_kwargs={"image": image, "embeddings": emb_ref})
        report(f"seed {seed} fp16 + fp32 embeddings", sum(o.clamp(-1, 1) for o in out["all_model_outputs"]).cuda(), ref)
        mp = build("fp16", 256)
        report(f"seed {seed} fp16, plain fp16 encoder", mp(image=image, pred_type="ddim_sample", noise=noise), ref)
        del mp
        mb = build("bf16")
        report(f"seed {seed} bf16", mb(image=image, pred_type="ddim_sample", noise=noise), ref)
        out = mb.sample_diffusion.ddim_sample_loop(mb.model, noise.shape, noise=noise, model_kwargs={"image": image, "embeddings": emb_ref})
        report(f"seed {seed} bf16 + fp32 embeddings", sum(o.clamp(-1, 1) for o in out["all_model_outputs"]).cuda(), ref)
        del m16, mb
