"""Staged bring-up probe for the B200 box: each stage runs in its own subprocess (a trapped kernel poisons the CUDA
context) with a timeout, and prints diagnostics that tell WHAT is wrong, not just that something is.

usage: python tools/gpu_probe.py [stage ...]      (no args: all stages)
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _conv_op(src0, w, src1=None, ref=False):
    import ctypes

    import torch

    from diff_unet_amos_b200 import _lib

    lib = _lib.load()
    B, c0, D, H, W = src0.shape
    c1 = 0 if src1 is None else src1.shape[1]
    cout = w.shape[0]
    out = torch.empty((B, cout, D, H, W), device="cuda", dtype=torch.float32)
    p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
    _lib.check(lib.dunet_op_conv3x3x3(p(src0), c0, p(src1), c1, p(w.contiguous()), cout, p(out), B, _lib.i32x3((D, H, W)),
                                      1 if ref else 0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return out


def _bf(t):
    return t.to(__import__("torch").bfloat16).float()


def stage_load():
    import torch

    from diff_unet_amos_b200 import _lib

    lib = _lib.load()
    print("version", lib.dunet_version(), "device", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))


def stage_coord(ref=False):
    """one-hot weights: out channel 0/1/2 must return the x/y/z coordinate (+1) of the tap's source voxel."""
    import torch

    D, H, W = 4, 16, 8
    x = torch.zeros(1, 64, D, H, W, device="cuda")
    zz, yy, xx = torch.meshgrid(torch.arange(D), torch.arange(H), torch.arange(W), indexing="ij")
    x[0, 0], x[0, 1], x[0, 2] = (xx + 1).float().cuda(), (yy + 1).float().cuda(), (zz + 1).float().cuda()
    bad_total = 0
    for tap in (13, 0, 1, 3, 9, 26, 14):
        tz, ty, tx = tap // 9, (tap // 3) % 3, tap % 3
        w = torch.zeros(64, 64, 3, 3, 3, device="cuda")
        for c in range(3):
            w[c, c, tz, ty, tx] = 1.0
        out = _conv_op(x, w, ref=ref)
        exp = torch.zeros(3, D, H, W)
        for c, (g, n, t) in enumerate(((xx, W, tx), (yy, H, ty), (zz, D, tz))):
            pass
        sx, sy, sz = xx + tx - 1, yy + ty - 1, zz + tz - 1
        ok = (sx >= 0) & (sx < W) & (sy >= 0) & (sy < H) & (sz >= 0) & (sz < D)
        exp[0], exp[1], exp[2] = (sx + 1) * ok, (sy + 1) * ok, (sz + 1) * ok
        got = out[0, :3].cpu()
        bad = (got != exp).sum().item()
        rest = out[0, 3:].abs().max().item()
        bad_total += bad
        print(f"tap {tap} (tz{tz} ty{ty} tx{tx}): mismatches {bad}/{exp.numel()}  other-channels max {rest}")
        if bad:
            print(" expected x-src row z=1:", exp[0, 1, :3].int().tolist())
            print(" got      x-src row z=1:", got[0, 1, :3].int().tolist())
            print(" expected y-src        :", exp[1, 1, :3].int().tolist())
            print(" got      y-src        :", got[1, 1, :3].int().tolist())
            print(" expected z-src z=0..3 :", exp[2, :, 5, 3].int().tolist(), " got:", got[2, :, 5, 3].int().tolist())
    print("COORD", "REF" if ref else "TC", "OK" if bad_total == 0 else "FAIL")


def stage_coord_ref():
    stage_coord(ref=True)


def _cmp(name, got, exp):
    import torch

    err = (got - exp).norm() / exp.norm().clamp_min(1e-20)
    print(f"{name}: rel-l2 {err.item():.3e}  max-abs {(got - exp).abs().max().item():.3e}  ref-absmax {exp.abs().max().item():.3e}")
    return err.item()


def stage_conv_random():
    import torch
    import torch.nn.functional as F

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    cases = [("64->64 16^3", 64, 0, 64, 1, (16, 16, 16)), ("17->64 (pad32) 16^3", 17, 0, 64, 1, (16, 16, 16)),
             ("1->64 (pad32) 16^3", 1, 0, 64, 1, (16, 16, 16)), ("64+64->64 16^3", 64, 64, 64, 1, (16, 16, 16)),
             ("64->128 8^3 B2", 64, 0, 128, 2, (8, 8, 8)), ("128->128 12^3", 128, 0, 128, 1, (12, 12, 12)),
             ("256+256->256 6^3", 256, 256, 256, 1, (6, 6, 6)), ("512->512 2^3 B3", 512, 0, 512, 3, (2, 2, 2)),
             ("64->64 32x48x40", 64, 0, 64, 1, (32, 48, 40)), ("8->8 (small feats) 16^3", 8, 0, 8, 1, (16, 16, 16))]
    worst = 0.0
    for name, c0, c1, cout, B, dims in cases:
        s0 = torch.randn(B, c0, *dims, device="cuda")
        s1 = torch.randn(B, c1, *dims, device="cuda") if c1 else None
        w = torch.randn(cout, c0 + c1, 3, 3, 3, device="cuda") / (27 * (c0 + c1)) ** 0.5
        xin = _bf(s0 if s1 is None else torch.cat([s0, s1], 1))
        exp = F.conv3d(xin.double(), _bf(w).double(), padding=1).float()
        e1 = _cmp(name + " [tc ]", _conv_op(s0, w, s1, ref=False), exp)
        e2 = _cmp(name + " [ref]", _conv_op(s0, w, s1, ref=True), exp)
        worst = max(worst, e1)
    print("CONV_RANDOM", "OK" if worst < 5e-3 else "FAIL", worst)


def stage_timeouts():
    import ctypes

    from diff_unet_amos_b200 import _lib

    f = ctypes.c_uint32(0)
    _lib.check(_lib.load().dunet_debug_barrier_timeouts(ctypes.byref(f)))
    print("barrier timeout flag", f.value)


STAGES = {"load": stage_load, "coord_ref": stage_coord_ref, "coord": stage_coord, "conv_random": stage_conv_random,
          "timeouts": stage_timeouts}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--stage":
        STAGES[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(STAGES)
    for n in names:
        print(f"===== stage {n}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--stage", n], timeout=240, capture_output=True, text=True)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print(f"stage {n} exit code {r.returncode}\n{r.stderr[-3000:]}")
        except subprocess.TimeoutExpired as e:
            print(f"stage {n} TIMED OUT\n{(e.stdout or b'')[-2000:]}")
        sys.stdout.flush()
