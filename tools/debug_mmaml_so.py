#!/usr/bin/env python
"""Which op's second-order terms are off?  The MMAML meta-step at the real shapes with each op class switchable between
the B200 Function and plain torch ops (fp32, CUDA), against the all-torch fp64 run."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
import torch
import torch.nn.functional as F
from b200np import lib
from b200np.engine import AggregateFn, LinearFn
from b200np.mmaml import BnActFn, Conv3x3S2Fn
from trainer.losses import LossFunc

P = lib.PREC_FP32_SIMT
lossf = LossFunc("mse", "shapenet_1d")
CH = [1, 32, 64, 128, 256]
N, HW = int(os.environ.get("N", 15)), int(os.environ.get("HW", 128))
STEPS, LR = int(os.environ.get("STEPS", 2)), 0.05


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(tuple(shape), generator=g, dtype=torch.float64) * scale


def net(x, ps, films, use):
    h = x                                             # NHWC
    for i in range(4):
        w, b = ps[2 * i], ps[2 * i + 1]
        gam, bet = films[i][:CH[i + 1]], films[i][CH[i + 1]:]
        if use["conv"]:
            h = Conv3x3S2Fn.apply(P, h, w, b)
        else:
            h = F.conv2d(h.permute(0, 3, 1, 2), w, b, stride=2, padding=1).permute(0, 2, 3, 1)
        if use["bn"]:
            h = BnActFn.apply(h, gam, bet, 1.0, True, 1e-5, None, None, 0.1)
        else:
            mu, var = h.mean((0, 1, 2)), h.var((0, 1, 2), unbiased=False)
            h = torch.relu((h - mu) * (var + 1e-5).rsqrt() * (1 + gam) + bet)
    n = h.shape[0]
    if use["mean"]:
        feat = AggregateFn.apply(0, h.reshape(n, -1, h.shape[-1]))
    else:
        feat = h.reshape(n, -1, h.shape[-1]).mean(1)
    if use["lin"]:
        return LinearFn.apply(lib.ACT_TANH, P, feat, None, ps[8], ps[9])
    return torch.tanh(feat @ ps[8].t() + ps[9])


def loss(pred, y, use):
    if use["loss"]:
        return lossf.calc_loss(pred, None, y)
    return torch.mean(torch.sum((y[..., :2] - pred) ** 2, dim=-1))


def meta(x, y, xv, yv, ps, films, use):
    cur = list(ps)
    for _ in range(STEPS):
        l = loss(net(x, cur, films, use), y, use)
        g = torch.autograd.grad(l, cur, create_graph=True, allow_unused=True)
        cur = [p if gi is None else p - LR * gi.clamp(-20, 20) for p, gi in zip(cur, g)]
    return loss(net(xv, cur, films, use), yv, use)


shapes = []
for i in range(4):
    shapes += [(CH[i + 1], CH[i], 3, 3), (CH[i + 1],)]
shapes += [(2, 256), (2,)]
ps64 = [rnd(*s, seed=i, scale=(2.0 / (s[1] * 9 + s[0] * 9)) ** 0.5 if len(s) == 4 else 0.1) for i, s in enumerate(shapes)]
films64 = [rnd(2 * CH[i + 1], seed=40 + i, scale=0.3) for i in range(4)]
x64, xv64 = torch.rand(N, HW, HW, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(7)), \
    torch.rand(N, HW, HW, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(8))
y64, yv64 = rnd(N, 2, seed=9), rnd(N, 2, seed=10)


torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
if os.environ.get("REAL"):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["B200NP_MMAML"] = "1"
    import test_mmaml as tm
    m_, _ = tm._build_models()
    ps64 = [p.detach().double() for _, p in m_.named_parameters()]
    if os.environ.get("REAL") == "bias":          # same weights, non-zero biases
        ps64 = [p if p.dim() > 1 else rnd(*p.shape, seed=77, scale=0.1) for p in ps64]


def run(use, dtype):
    cv = lambda t: t.to(dtype).cuda()
    ps = [cv(t).requires_grad_() for t in ps64]
    films = [cv(t).requires_grad_() for t in films64]
    o = meta(cv(x64), cv(y64), cv(xv64), cv(yv64), ps, films, use)
    o.backward()
    return float(o), [t.grad.double().cpu() for t in ps + films]


none = dict(conv=False, bn=False, mean=False, lin=False, loss=False)
o64, g64 = run(none, torch.float64)
o32, g32 = run(none, torch.float32)
rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-30))
cat = lambda gs: torch.cat([g.reshape(-1) for g in gs])
print(f"N={N} HW={HW} STEPS={STEPS}  outer {o64:.6f}; torch fp32 vs fp64: {rel(cat(g32), cat(g64)):.2e}")
for name in ("conv", "bn", "mean", "lin", "loss", "all"):
    use = dict(none) if name != "all" else {k: True for k in none}
    if name != "all":
        use[name] = True
    o, g = run(use, torch.float32)
    worst = max((rel(a, b), i) for i, (a, b) in enumerate(zip(g, g64)) if float(b.norm()) > 1e-9)
    print(f"  B200 {name:5s}: outer {o:.6f}  concatenated {rel(cat(g), cat(g64)):.2e}  worst tensor {worst[0]:.2e} (#{worst[1]})")
