import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
from b200np import ops
from b200np.lib import LIB
N = 1140
g = torch.Generator().manual_seed(0)
x = torch.rand(N, 64, 64, 64, generator=g).cuda(); h = torch.rand(N, 32, 32, 64, generator=g).cuda()
w2 = ops.pack_conv_weight((torch.randn(64, 64, 3, 3, generator=g) * 0.04).cuda())
ws = ops.pack_conv_weight((torch.randn(64, 64, 1, 1, generator=g) * 0.1).cuda())
b = torch.zeros(64, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
for fl in (0, 256, 512, 768):
    LIB.b200np_debug_set_halo_flags(fl)
    print("flags", fl, "fwd conv2+skip ms", t(lambda: ops.conv_fwd(h, w2, b, 1, 1, 1, skip=(x, ws, b, 2))), "dgrad s1", t(lambda: ops.conv_dgrad(h, w2, h.shape, 1, 1, mask_src=h)))
LIB.b200np_debug_set_halo_flags(0)
