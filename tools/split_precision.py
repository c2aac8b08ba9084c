#!/usr/bin/env python
"""rel-L2 against fp64 of one layer1-sized conv in its three autograd roles (forward, data gradient, weight gradient)
for the tensor-core precision modes -- the calibration behind the `B200NP_LO_RN` choice in csrc/umma.cuh."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
from b200np import ops  # noqa: E402

g = torch.Generator().manual_seed(0)
N, H = 16, 32
x = torch.rand(N, 64, H, H, generator=g, dtype=torch.float64)
w = torch.randn(64, 64, 3, 3, generator=g, dtype=torch.float64) * 0.04
b = torch.randn(64, generator=g, dtype=torch.float64) * 0.1
dy = torch.randn(N, 64, H, H, generator=g, dtype=torch.float64)
x.requires_grad_(True), w.requires_grad_(True)
y = F.conv2d(x, w, b, padding=1)
y.backward(dy)
nh = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().float().cuda()
rel = lambda a, r: float((a.double().cpu() - r).norm() / r.norm())
pw = ops.pack_conv_weight(w.detach().float().cuda())
xg, dyg, bg = nh(x), nh(dy), b.float().cuda()
for name, prec in (("fp32", 0), ("tf32x3", 1), ("tf32", 2)):
    yo = ops.conv_fwd(xg, pw, bg, 1, 0, prec)
    dx = ops.conv_dgrad(dyg, pw, xg.shape, 1, prec, mask_src=None)
    dw, db, _ = ops.conv_wgrad(xg, dyg, 3, 1, prec)
    print(f"{name:7s} fwd {rel(yo.permute(0, 3, 1, 2), y.detach()):.3e}  dgrad {rel(dx.permute(0, 3, 1, 2), x.grad):.3e}  "
          f"wgrad {rel(dw, w.grad):.3e}", flush=True)
