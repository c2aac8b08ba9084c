import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
from b200np import ops
from b200np.lib import LIB
N = 1140
g = torch.Generator().manual_seed(0)
x = torch.rand(N, 64, 64, 64, generator=g).cuda(); dz = torch.randn(N, 32, 32, 64, generator=g).cuda()
w1 = ops.pack_conv_weight((torch.randn(64, 64, 3, 3, generator=g) * 0.04).cuda())
ws = ops.pack_conv_weight((torch.randn(64, 64, 1, 1, generator=g) * 0.1).cuda())
ref = None
for mt in (4, 2, 1):
    LIB.b200np_debug_set_halo_min_taps(mt)
    fn = lambda: ops.conv_dgrad(dz, w1, x.shape, 2, 1, mask_src=x, skip=(dz, ws, 2))
    out = fn(); torch.cuda.synchronize()
    if ref is None: ref = out
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"halo for classes with >= {mt} taps: dgrad 3x3 s2 + skip {e0.elapsed_time(e1)/5:.3f} ms, max diff vs first {float((out-ref).abs().max()):.2e}")
