"""Per-CTA phase timeline of the tcgen05 conv kernel (clock64 stamps): where do the cycles of one CTA go?"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diff_unet_amos_b200 import _lib

lib = _lib.load()
p = lambda t: None if t is None else ctypes.c_void_p(t.data_ptr())
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, c0, c1, cout, dims in [("64->64 96^3", 64, 0, 64, (96, 96, 96)), ("64+64->64 96^3", 64, 64, 64, (96, 96, 96)),
                                 ("17->64 96^3", 17, 0, 64, (96, 96, 96)), ("128->128 24^3", 128, 0, 128, (24, 24, 24))]:
    x0 = torch.randn(1, c0, *dims, device="cuda")
    x1 = torch.randn(1, c1, *dims, device="cuda") if c1 else None
    w = torch.randn(cout, c0 + c1, 3, 3, 3, device="cuda") * 0.05
    out = torch.empty(1, cout, *dims, device="cuda")
    dbg = torch.zeros(8 * 4096, dtype=torch.int64, device="cuda")
    for rep in range(2):
        lib.dunet_debug_set_conv_timeline(p(dbg))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        _lib.check(lib.dunet_profile_enable(1))
        _lib.check(lib.dunet_op_conv3x3x3(p(x0), c0, p(x1), c1, p(w), cout, p(out), 1, _lib.i32x3(dims), 0, st()))
        torch.cuda.synchronize()
        ms, n, fl = ctypes.c_double(), ctypes.c_uint64(), ctypes.c_double()
        _lib.check(lib.dunet_profile_read(ctypes.byref(ms), ctypes.byref(n), ctypes.byref(fl)))
        lib.dunet_profile_enable(0)
        lib.dunet_debug_set_conv_timeline(None)
    t = dbg.view(-1, 8).cpu()
    t = t[t[:, 0] != 0][:, :5].double()
    flops = 2.0 * x0[0, 0].numel() * cout * 27 * (c0 + c1)
    print(f"{name}: kernel {ms.value * 1e3:.1f} us  {flops / ms.value / 1e9:.0f} TFLOP/s  ctas {len(t)}")
    if len(t):
        d = t[:, 1:] - t[:, :-1]
        print("   mean cycles: start->first MMA issued %.0f | MMA issue span %.0f | issue-end->acc complete %.0f | epilogue %.0f | total %.0f"
              % (d[:, 0].mean(), d[:, 1].mean(), d[:, 2].mean(), d[:, 3].mean(), (t[:, 4] - t[:, 0]).mean()))
