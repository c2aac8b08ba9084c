"""Prototype check + timing (see wgrad_halo_proto.cu / wgrad_halo_proto2.cu):
    python tools/probe/run_wgrad_halo_proto.py [N] [v2]
Build first (in tools/probe):
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -shared -Xcompiler -fPIC -I../../include \
         -I../../what-matters-for-meta-learning_b200/csrc --expt-relaxed-constexpr -o libwgrad_halo_proto.so wgrad_halo_proto.cu
(same for wgrad_halo_proto2.cu -> libwgrad_halo_proto2.so)"""
import ctypes, os, sys, torch
here = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(here))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
from b200np import ops  # noqa: E402
V2 = "v2" in sys.argv[1:]
lib = ctypes.CDLL(os.path.join(here, "libwgrad_halo_proto2.so" if V2 else "libwgrad_halo_proto.so"))
entry = lib.wgrad_halo_proto2 if V2 else lib.wgrad_halo_proto
entry.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 5 + [ctypes.c_void_p]
CH = 74


def proto(x, dy, x3):
    N, H, W, _ = x.shape
    part = torch.zeros(2, CH, 3, 128, 64, device="cuda")
    rc = entry(x.data_ptr(), dy.data_ptr(), part.data_ptr(), N, H, W, CH, x3, None)
    assert rc == 0, rc
    P = part.sum(1)                                            # [role][dx + 1][m][co]
    dw = torch.zeros(2, 64, 64, 3, 3, device="cuda")           # [role][co][ci][ky][kx]; row ky = 1 comes from both roles
    for role in range(2):
        for t2 in range(2):
            ky = role + t2
            dw[role, :, :, ky, :] = P[role, :, t2 * 64:(t2 + 1) * 64, :].permute(2, 1, 0)
    return dw


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g = torch.Generator().manual_seed(0)
for N in (4, int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1140):
    x = torch.rand(N, 32, 32, 64, generator=g).cuda()
    dy = torch.randn(N, 32, 32, 64, generator=g).cuda()
    ref, _, _ = ops.conv_wgrad(x, dy, 3, 1, 0)                 # CUDA-core fp32 kernel: [co][ci][ky][kx]
    for x3 in (1, 0):
        dw = proto(x, dy, x3)
        prod, _, _ = ops.conv_wgrad(x, dy, 3, 1, 1 if x3 else 2)
        rel = lambda a, b: float((a - b).norm() / b.norm())
        msg = (f"N={N} {'tf32x3' if x3 else 'tf32  '}: rel-L2 vs fp32 kernel  proto rows ky=0,1 (role 0) {rel(dw[0][:, :, :2], ref[:, :, :2]):.3e}  "
               f"ky=1,2 (role 1) {rel(dw[1][:, :, 1:], ref[:, :, 1:]):.3e}  product kernel {rel(prod, ref):.3e}")
        part = torch.zeros(2, CH, 3, 128, 64, device="cuda")
        t_p = timed(lambda: entry(x.data_ptr(), dy.data_ptr(), part.data_ptr(), N, 32, 32, CH, x3, None))
        t_k = timed(lambda: ops.conv_wgrad(x, dy, 3, 1, 1 if x3 else 2))
        print(msg + f" | proto (12 tap-products) {t_p:.3f} ms, product kernel (9 taps + reduction) {t_k:.3f} ms", flush=True)
