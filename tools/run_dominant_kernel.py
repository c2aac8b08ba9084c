#!/usr/bin/env python
"""Runs the dominant kernels of the step alone at the bench's size (for `ncu --set full`):
layer1 conv2 3x3 s1 + 1x1 s2 skip (forward), its data gradients and weight gradients."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
from b200np import ops  # noqa: E402

if "B200NP_WGRAD_WAVES" in os.environ:   # diagnostic entry point: pixel chunks per weight-gradient launch
    ops.LIB.b200np_debug_set_wgrad_waves(int(os.environ["B200NP_WGRAD_WAVES"]))

prec = {"fp32": 0, "tf32x3": 1, "tf32": 2}[sys.argv[1] if len(sys.argv) > 1 else "tf32x3"]
what = sys.argv[2] if len(sys.argv) > 2 else "all"
N = int(sys.argv[3]) if len(sys.argv) > 3 else 1140
dev = "cuda"
g = torch.Generator().manual_seed(0)
x = torch.rand(N, 64, 64, 64, generator=g).to(dev)
h = torch.rand(N, 32, 32, 64, generator=g).to(dev)
dz = torch.randn(N, 32, 32, 64, generator=g).to(dev)
w2 = (torch.randn(64, 64, 3, 3, generator=g) * 0.04).to(dev)
w1 = (torch.randn(64, 64, 3, 3, generator=g) * 0.04).to(dev)
ws = (torch.randn(64, 64, 1, 1, generator=g) * 0.1).to(dev)
b = torch.zeros(64, device=dev)
wf2 = wd2 = ops.pack_conv_weight(w2)
wf1 = wd1 = ops.pack_conv_weight(w1)
wfs = wds = ops.pack_conv_weight(ws)
ev = lambda: torch.cuda.Event(enable_timing=True)


def timed(name, fn, flops, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:28s} {ms:8.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s (algorithmic)", flush=True)


M1 = N * 32 * 32
if what in ("all", "fwd"):
    timed("conv2 3x3 s1 + skip fwd", lambda: ops.conv_fwd(h, wf2, b, 1, 1, prec, skip=(x, wfs, b, 2)), 2.0 * M1 * 64 * 640)
    timed("conv1 3x3 s2 fwd", lambda: ops.conv_fwd(x, wf1, b, 2, 1, prec), 2.0 * M1 * 64 * 576)
if what in ("all", "dgrad"):
    timed("dgrad 3x3 s1 (mask)", lambda: ops.conv_dgrad(dz, wd2, h.shape, 1, prec, mask_src=h), 2.0 * M1 * 64 * 576)
    timed("dgrad 3x3 s2 + skip (mask)", lambda: ops.conv_dgrad(dz, wd1, x.shape, 2, prec, mask_src=x, skip=(dz, wds, 2)), 2.0 * M1 * 64 * 640)
if what in ("all", "dgrad") and prec != 0:
    def pack_bits(act):
        g_ = (act > 0).to(torch.int64).reshape(*act.shape[:-1], 2, 32)
        w_ = (g_ << torch.arange(32, device=act.device, dtype=torch.int64)).sum(-1)
        return torch.where(w_ >= 2 ** 31, w_ - 2 ** 32, w_).to(torch.int32).contiguous()
    bh, bx = pack_bits(h), pack_bits(x)
    timed("dgrad 3x3 s1 (packed gates)", lambda: ops.conv_dgrad(dz, wd2, h.shape, 1, prec, mask_bits=bh), 2.0 * M1 * 64 * 576)
    timed("dgrad 3x3 s2 + skip (gates)", lambda: ops.conv_dgrad(dz, wd1, x.shape, 2, prec, mask_bits=bx, skip=(dz, wds, 2)), 2.0 * M1 * 64 * 640)
if what in ("all", "wgrad"):
    timed("wgrad 3x3 s1", lambda: ops.conv_wgrad(h, dz, 3, 1, prec), 2.0 * M1 * 64 * 576)
    timed("wgrad 3x3 s2", lambda: ops.conv_wgrad(x, dz, 3, 2, prec), 2.0 * M1 * 64 * 576)
    timed("wgrad 1x1 s2", lambda: ops.conv_wgrad(x, dz, 1, 2, prec, want_db=False), 2.0 * M1 * 64 * 64)
if what in ("all", "stem"):
    img = torch.rand(N, 1, 128, 128, generator=g).to(dev)
    wst = (torch.randn(64, 1, 5, 5, generator=g) * 0.2).to(dev)
    dy0 = torch.randn(N, 64, 64, 64, generator=g).to(dev)
    timed("stem 5x5 s2 fwd", lambda: ops.conv_small_fwd(img, wst, b, prec=prec), 2.0 * N * 4096 * 64 * 25)
    timed("stem 5x5 s2 wgrad", lambda: ops.conv_small_wgrad(img, dy0, wst.shape, prec), 2.0 * N * 4096 * 64 * 25)
if what in ("all", "wgrad"):
    # accuracy of the long pixel contraction at full size: tensor-core modes against the CUDA-core fp32 kernel
    ref, _, _ = ops.conv_wgrad(h, dz, 3, 1, 0)
    got, _, _ = ops.conv_wgrad(h, dz, 3, 1, prec)
    print(f"wgrad 3x3 s1 full size: rel-L2 vs fp32 CUDA-core kernel = {float((got - ref).norm() / ref.norm()):.3e}", flush=True)
