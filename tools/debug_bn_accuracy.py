#!/usr/bin/env python
"""Accuracy of the batch-norm block kernels (forward, backward, second-order backward) against fp64 for inputs with a
large mean / std ratio and few-valued inputs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
import torch
from b200np import ops

rel = lambda a, b: float((a.double().cpu() - b.cpu()).norm() / b.cpu().norm().clamp_min(1e-30))
g = torch.Generator().manual_seed(0)
for R, C, mean, std in ((61440, 32, 0.0, 1.0), (61440, 32, 0.1, 0.02), (61440, 32, 1.0, 0.01), (960, 256, 0.1, 0.02)):
    x = (torch.randn(R, C, generator=g, dtype=torch.float64) * std + mean).float().double()   # exactly representable in fp32
    dy = torch.randn(R, C, generator=g, dtype=torch.float64).float().double()
    sc = (torch.randn(C, generator=g, dtype=torch.float64) * 0.3).float().double()
    sh = (torch.randn(C, generator=g, dtype=torch.float64) * 0.3).float().double()
    vx = torch.randn(R, C, generator=g, dtype=torch.float64).float().double()
    xr, dyr, scr, shr = (t.clone().cuda().requires_grad_() for t in (x, dy, sc, sh))
    mu, var = xr.mean(0), xr.var(0, unbiased=False)
    xh = (xr - mu) * (var + 1e-5).rsqrt()
    y = torch.relu(xh * (scr + 1.0) + shr)
    first = torch.autograd.grad(y, (xr, scr, shr), dyr, create_graph=True)
    second = torch.autograd.grad((first[0] * vx.cuda()).sum(), (xr, dyr, scr))
    f = lambda t: t.float().cuda().contiguous()
    yk, mk, rk = ops.bn_act_fwd(f(x), f(sc), f(sh), 1.0, True, 1e-5)
    # gate agreement matters: count mismatches
    flips = int(((yk > 0) != (y > 0)).sum())
    dxk, dsk, dtk = ops.bn_act_bwd(f(dy), yk, f(x), mk, rk, f(sc), 1.0, True)
    gx, gdy, gs = ops.bn_act_bwd2(f(dy), yk, f(x), mk, rk, f(sc), 1.0, True, f(vx), None, None)
    print(f"R={R} C={C} mean={mean} std={std}: y {rel(yk, y.detach()):.1e} mean {rel(mk[:C], mu.detach()):.1e} "
          f"rstd {rel(rk, (var + 1e-5).rsqrt().detach()):.1e} flips {flips} | dx {rel(dxk, first[0].detach()):.1e} "
          f"dscale {rel(dsk, first[1].detach()):.1e} dshift {rel(dtk, first[2].detach()):.1e} | "
          f"gx {rel(gx, second[0]):.1e} gdy {rel(gdy, second[1]):.1e} gscale {rel(gs, second[2]):.1e}")
    # same elementwise arithmetic in fp32, but with the fp64 statistics rounded to fp32: how many gates still flip?
    mu32, rs32 = mu.detach().float(), (var + 1e-5).rsqrt().detach().float()
    y_exact_stats = torch.relu((f(x) - mu32) * rs32 * (f(sc) + 1.0) + f(sh))
    print(f"     gates flipped with exact statistics rounded to fp32: {int(((y_exact_stats > 0) != (y > 0)).sum())};"
          f" |mean error| / std: ours {float((((mk[:C].double() + mk[C:].double()).cpu() - mu.detach().cpu()).abs().cuda() / var.detach().sqrt()).max()):.1e},"
          f" rounded exact {float(((mu32.double() - mu.detach()).abs() / var.detach().sqrt()).max()):.1e}")
