"""Soak test of the bit-identity properties (run by hand on a B200): normalise-on-load vs the unfused path, run-to-run determinism and
batch transparency, many repetitions in one process with other work interleaved.  Prints the number of mismatching repetitions."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import diff_unet_amos_b200 as pkg
from diff_unet_amos_b200 import _lib
from oracle import oracle_model
from tests.util import seeded_image, seeded_noise

REPS = int(os.environ.get("REPS", "40"))


def build(cout, S, **kw):
    torch.manual_seed(0)
    return pkg.DiffUNetB200(in_channels=1, out_channels=cout, image_size=S, spatial_size=S, features=oracle_model.DEFAULT_FEATURES, **kw).to("cuda").eval()


bad = {"fused_vs_unfused": 0, "rerun": 0, "batching": 0}
for S in (64, 96):
    cout = 3
    ma = build(cout, S, num_steps=3, batch_max=2)
    mb = build(cout, S, num_steps=3, batch_max=2, debug_flags=_lib.DUNET_FLAG_NO_FUSED_NORM)
    image, noise = seeded_image((2, 1, S, S, S)).cuda(), seeded_noise((2, cout, S, S, S)).cuda()
    ref = ma(image=image, pred_type="ddim_sample", noise=noise).clone()
    junk = torch.empty(64 << 20, device="cuda")
    for rep in range(REPS):
        junk.normal_()  # unrelated traffic between the calls (L2 contents, allocator state, clocks)
        a = ma(image=image, pred_type="ddim_sample", noise=noise)
        b = mb(image=image, pred_type="ddim_sample", noise=noise)
        one = ma(image=image[1:2], pred_type="ddim_sample", noise=noise[1:2])
        bad["rerun"] += int(not torch.equal(a, ref))
        bad["fused_vs_unfused"] += int(not torch.equal(a, b))
        bad["batching"] += int(not torch.equal(one[0], ref[1]))
    print(f"S={S}: {REPS} repetitions, mismatches so far {bad}", flush=True)
print("SOAK", "CLEAN" if not any(bad.values()) else "MISMATCH", bad)
