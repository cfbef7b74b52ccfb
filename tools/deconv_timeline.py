"""Per-unit timeline of the persistent transposed-conv kernel (clock64 stamps)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diff_unet_amos_b200 import _lib
lib = _lib.load()
p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
B, cin, cout, dims = int(sys.argv[1]) if len(sys.argv) > 1 else 1, 64, 64, (48, 48, 48)
x = torch.randn(B, cin, *dims, device="cuda"); w = torch.randn(cin, cout, 2, 2, 2, device="cuda") * 0.1; b = torch.randn(cout, device="cuda")
out = torch.empty(B, cout, *(2 * d for d in dims), device="cuda")
dbg = torch.zeros(148 * 64, dtype=torch.int64, device="cuda")
for rep in range(2):
    dbg.zero_()
    lib.dunet_debug_set_conv_timeline(p(dbg))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(lib.dunet_op_deconv2x2x2(p(x), cin, p(w), p(b), cout, p(out), B, _lib.i32x3(dims), 0, st()))
    e1.record(); torch.cuda.synchronize()
    lib.dunet_debug_set_conv_timeline(None)
print("op total (incl. pack/unpack) ms", e0.elapsed_time(e1))
t = dbg.view(148, 64).cpu().double()
for cta in (0, 1, 77, 147):
    r = t[cta]
    units = [(r[u] - r[62]).item() for u in range(60) if r[u] > 0]
    print(f"CTA {cta}: {len(units)} units, kernel cycles {(r[63]-r[62]).item():.0f}; unit-done stamps:", [int(u) for u in units[:14]])
g0, g1 = t[:, 60], t[:, 61]
print(f"globaltimer: kernel span {(g1.max() - g0.min()).item() / 1e3:.1f} us; CTA start spread {(g0.max() - g0.min()).item() / 1e3:.1f} us; "
      f"per-CTA duration min {(g1 - g0).min().item() / 1e3:.1f} max {(g1 - g0).max().item() / 1e3:.1f} us; clock64 per CTA {((t[:,63]-t[:,62]).mean()).item():.0f} cycles")
