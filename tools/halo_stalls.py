#!/usr/bin/env python
"""Where does the persistent halo kernel wait?  Per-role blocked cycles (diagnostic hook in tapconv_halo.cu)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
from b200np import ops
from b200np.lib import LIB
prec = {"tf32x3": 1, "tf32": 2}[sys.argv[1] if len(sys.argv) > 1 else "tf32x3"]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1140
g = torch.Generator().manual_seed(0)
x = torch.rand(N, 64, 64, 64, generator=g).cuda(); h = torch.rand(N, 32, 32, 64, generator=g).cuda()
w2 = ops.pack_conv_weight((torch.randn(64, 64, 3, 3, generator=g) * 0.04).cuda())
ws = ops.pack_conv_weight((torch.randn(64, 64, 1, 1, generator=g) * 0.1).cuda())
b = torch.zeros(64, device="cuda")
dbg = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
LIB.b200np_debug_set_halo_timing.argtypes = [ctypes.c_void_p]
flag_list = [int(f) for f in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
ev = lambda: torch.cuda.Event(enable_timing=True)
def report(name, fl):
    d = dbg.view(148, 8).double().mean(0).tolist()
    tiles = N * 8 / 148
    print(f"  [{fl}] cycles per tile: MMA-lane total {d[6]/tiles:.0f} = wait acc_empty {d[3]/tiles:.0f} + wait a_full {d[4]/tiles:.0f} "
          f"+ wait b_full {d[5]/tiles:.0f} + issue/other {(d[6]-d[3]-d[4]-d[5])/tiles:.0f} | producer total {d[1]/tiles:.0f}, blocked on a_empty {d[0]/tiles:.0f} "
          f"| weight warp blocked on b_empty {d[2]/tiles:.0f} | epilogue blocked on acc_full {d[7]/tiles:.0f}")


for name, fn in (("fwd conv1 s2", lambda: ops.conv_fwd(x, w2, b, 2, 1, prec)),
                 ("fwd conv2+skip", lambda: ops.conv_fwd(h, w2, b, 1, 1, prec, skip=(x, ws, b, 2))),
                 ("dgrad s1", lambda: ops.conv_dgrad(h, w2, h.shape, 1, prec, mask_src=h)),
                 ("dgrad s2 + skip (4 classes fused)", lambda: ops.conv_dgrad(h, w2, x.shape, 2, prec, mask_src=x, skip=(h, ws, 2)))):
    for fl in flag_list:  # diagnostic flags, see tapconv_halo.cu (1 weights, 2 planes, 4 epilogue, 8 cross MMAs, 16 MMAs)
        LIB.b200np_debug_set_halo_flags(fl)
        fn(); torch.cuda.synchronize()
        e0, e1 = ev(), ev(); e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name}: flags {fl:2d}: {e0.elapsed_time(e1) / 5:.3f} ms")
        LIB.b200np_debug_set_halo_timing(ctypes.c_void_p(dbg.data_ptr()))
        dbg.zero_(); fn(); torch.cuda.synchronize()
        LIB.b200np_debug_set_halo_timing(ctypes.c_void_p(0))
        report(name, fl)
    LIB.b200np_debug_set_halo_flags(0)
