"""Prototype 3 (wgrad_halo_proto3.cu) against the product's b200np_conv_wgrad: parity of dW, dW_skip, db and timing.
    python tools/probe/run_wgrad_halo_proto3.py [N]
Build first (in tools/probe):
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -shared -Xcompiler -fPIC -I../../include \\
         -I../../what-matters-for-meta-learning_b200/csrc --expt-relaxed-constexpr -o libwgrad_halo_proto3.so wgrad_halo_proto3.cu"""
import ctypes, os, sys, torch
here = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(here))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
from b200np import ops  # noqa: E402
lib = ctypes.CDLL(os.path.join(here, "libwgrad_halo_proto3.so"))
lib.wgrad_halo_proto3.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 4 + [ctypes.c_void_p]


def proto(x, dy, xs, x3):
    N, H, W, _ = x.shape
    nt = 10 if xs is not None else 9
    part = torch.zeros(148, nt, 64, 64, device="cuda")       # [chunk][tap][co][ci]
    pdb = torch.zeros(148, 64, device="cuda")
    ch = lib.wgrad_halo_proto3(x.data_ptr(), dy.data_ptr(), xs.data_ptr() if xs is not None else None, part.data_ptr(),
                               pdb.data_ptr(), N, H, W, x3, None)
    assert ch > 0, ch
    P = part[:ch].sum(0)                                      # [tap][co][ci]
    dw = P[:9].permute(1, 2, 0).reshape(64, 64, 3, 3)         # [co][ci][ky][kx]
    return dw, (P[9] if nt == 10 else None), pdb[:ch].sum(0), (part, pdb)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rel = lambda a, b: float((a - b).norm() / b.norm())
g = torch.Generator().manual_seed(0)
for N, HW in ((4, 32), (3, 16), (int(sys.argv[1]) if len(sys.argv) > 1 else 1140, 32)):
    x = torch.rand(N, HW, HW, 64, generator=g).cuda()
    xs = torch.rand(N, 2 * HW, 2 * HW, 64, generator=g).cuda()
    dy = torch.randn(N, HW, HW, 64, generator=g).cuda()
    ref, ref_db, _ = ops.conv_wgrad(x, dy, 3, 1, 0)                        # CUDA-core fp32 kernels
    ref_s, _, _ = ops.conv_wgrad(xs, dy, 1, 2, 0, want_db=False)
    for x3 in (1, 0):
        dw, dws, db, bufs = proto(x, dy, xs, x3)
        pw, pdb_, pws = ops.conv_wgrad(x, dy, 3, 1, 1 if x3 else 2, skip=(xs, 2))
        t_p = timed(lambda: lib.wgrad_halo_proto3(x.data_ptr(), dy.data_ptr(), xs.data_ptr(), bufs[0].data_ptr(),
                                                  bufs[1].data_ptr(), N, HW, HW, x3, None))
        t_k = timed(lambda: ops.conv_wgrad(x, dy, 3, 1, 1 if x3 else 2, skip=(xs, 2)))
        print(f"N={N} {HW}x{HW} {'tf32x3' if x3 else 'tf32  '}: rel-L2 vs fp32 kernels  proto dW {rel(dw, ref):.3e} dWskip "
              f"{rel(dws.reshape(ref_s.shape), ref_s):.3e} db {rel(db, ref_db):.3e} | product dW {rel(pw, ref):.3e} dWskip "
              f"{rel(pws, ref_s):.3e} | proto {t_p:.3f} ms (without the partial reduction), product {t_k:.3f} ms", flush=True)
