import sys, os
sys.path.insert(0, '/root/repo')
import torch
import diff_unet_amos_b200 as pkg
from diff_unet_amos_b200 import _lib
from oracle import oracle_model
from tests.util import seeded_image, seeded_noise
def build(cout, S, feats, **kw):
    torch.manual_seed(0)
    return pkg.DiffUNetB200(in_channels=1, out_channels=cout, image_size=S, spatial_size=S, features=feats, **kw).to("cuda").eval()
cout, S = 3, 64
image, noise = seeded_image((2, 1, S, S, S)).cuda(), seeded_noise((2, cout, S, S, S)).cuda()
ma = build(cout, S, oracle_model.DEFAULT_FEATURES, num_steps=3)
mb = build(cout, S, oracle_model.DEFAULT_FEATURES, num_steps=3, debug_flags=_lib.DUNET_FLAG_NO_FUSED_NORM)
for rep in range(3):
    a = ma(image=image, pred_type="ddim_sample", noise=noise)
    a2 = ma(image=image, pred_type="ddim_sample", noise=noise)
    b = mb(image=image, pred_type="ddim_sample", noise=noise)
    print("rep", rep, "a==a2", torch.equal(a, a2), "a==b", torch.equal(a, b), "max|a-b|", (a - b).abs().max().item(), "n diff", (a != b).sum().item(), "of", a.numel())
