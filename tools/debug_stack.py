import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200"), os.path.join(ROOT, "tests")]
import torch
import torch.nn.functional as F
from b200np import lib
from b200np.engine import AggregateFn, LinearFn
from b200np.mmaml import BnActFn, Conv3x3S2Fn
import test_second_order as t2
P = lib.PREC_FP32_SIMT
shapes = [(5, 16, 16, 3), (8, 3, 3, 3), (8,), (8,), (8,), (12, 8, 3, 3), (12,), (2, 12), (2,)]
ins64 = [t2.rnd(*s, seed=i) for i, s in enumerate(shapes)]
cot64 = t2.rnd(5, 2, seed=50)

def b200(x, w1, b1, g1, t1, w2, b2, wf, bf, upto=9):
    h = Conv3x3S2Fn.apply(P, x, w1, b1)
    if upto == 1: return h
    h = BnActFn.apply(h, g1, t1, 1.0, True, 1e-5, None, None, 0.1)
    if upto == 2: return h
    h = Conv3x3S2Fn.apply(P, h, w2, b2)
    if upto == 3: return h
    h = BnActFn.apply(h, None, None, 0.0, True, 1e-5, None, None, 0.1)
    if upto == 4: return h
    n = h.shape[0]
    feat = AggregateFn.apply(0, h.reshape(n, -1, h.shape[-1]))
    if upto == 5: return feat
    return LinearFn.apply(lib.ACT_TANH, P, feat, None, wf, bf)

def bn(z):
    mu, var = z.mean((0, 1, 2)), z.var((0, 1, 2), unbiased=False)
    return (z - mu) * (var + 1e-5).rsqrt()

def ref(x, w1, b1, g1, t1, w2, b2, wf, bf, upto=9):
    h = t2._conv_ref(x, w1, b1)
    if upto == 1: return h
    h = torch.relu(bn(h) * (1 + g1) + t1)
    if upto == 2: return h
    h = t2._conv_ref(h, w2, b2)
    if upto == 3: return h
    h = torch.relu(bn(h))
    if upto == 4: return h
    feat = h.mean((1, 2))
    if upto == 5: return feat
    return torch.tanh(feat @ wf.t() + bf)

for upto in (1, 2, 3, 4, 5, 9):
    for cg in (False, True):
        a = [t.float().cuda().requires_grad_() for t in ins64]
        o = b200(*a, upto=upto)
        ga = torch.autograd.grad(o.sum() if True else o, a[0], create_graph=cg, allow_unused=True)[0]
        r = [t.clone().cuda().requires_grad_() for t in ins64]
        gr = torch.autograd.grad(ref(*r, upto=upto).sum(), r[0], allow_unused=True)[0]
        na = None if ga is None else float(ga.norm())
        print(f"upto {upto} create_graph={cg}: |dx| ours {na} ref {None if gr is None else float(gr.norm()):.4e}")
