#!/usr/bin/env python
"""The second-order meta-step on the real modules (seeded init, hash images): B200 drop-in vs the functional oracle in
fp64 ON THE GPU, with switches: STEPS, EMB=0 (constant random FiLM inputs instead of the embedding model), LR."""
import os, sys
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200"), os.path.join(ROOT, "tests")]
os.environ["B200NP_MMAML"] = "1"
import numpy as np
import torch
import test_mmaml as tm
from b200np import engine
from oracle import mmaml_oracle, synth
from trainer.losses import LossFunc

STEPS, LR, EMB = int(os.environ.get("STEPS", 2)), float(os.environ.get("LR", 0.05)), int(os.environ.get("EMB", 1))
engine.set_precision(os.environ.get("PREC", "fp32"))
model, emb = tm._build_models()
model, emb = model.cuda(), emb.cuda()
lossf = LossFunc("mse", "shapenet_1d")
x_tr, y_tr, x_val, y_val = tm._meta_batches("cuda")
if os.environ.get("DATA") == "rand":
    gg = torch.Generator().manual_seed(11)
    x_tr = torch.rand(x_tr.shape, generator=gg).cuda()
    x_val = torch.rand(x_val.shape, generator=gg).cuda()
print("image stats: distinct values", int(torch.unique(x_tr).numel()), "mean", float(x_tr.mean()))
mse = lambda p, y: torch.mean(torch.sum((y[..., :2] - p) ** 2, dim=-1))
g = torch.Generator().manual_seed(3)
const_emb = [torch.randn(1, 2 * c, generator=g) * 0.3 for c in (32, 64, 128, 256)]


def ours():
    model.zero_grad(); emb.zero_grad()
    embeddings = emb(x_tr) if EMB else [e.cuda().requires_grad_() for e in const_emb]
    params = model.param_dict
    for _ in range(STEPS):
        loss = lossf.calc_loss(model(x_tr, params=params, embeddings=embeddings), None, y_tr)
        grads = torch.autograd.grad(loss, params.values(), create_graph=True, allow_unused=True)
        params = OrderedDict((n, p if gi is None else p - LR * gi.clamp(-20, 20)) for (n, p), gi in zip(params.items(), grads))
    outer = lossf.calc_loss(model(x_val, params=params, embeddings=embeddings), None, y_val)
    outer.backward()
    out = {f"model.{k}": p.grad.double().cpu() for k, p in model.named_parameters()}
    if EMB:
        out.update({f"emb.{k}": p.grad.double().cpu() for k, p in emb.named_parameters()})
    else:
        out.update({f"film{i}": e.grad.double().cpu() for i, e in enumerate(embeddings)})
    return float(outer), out


def truth(dtype=torch.float64, dev="cuda"):
    pm = {k: v.detach().to(dev, dtype).requires_grad_(True) for k, v in model.named_parameters()}
    pe = {k: v.detach().to(dev, dtype).requires_grad_(True) for k, v in emb.named_parameters()}
    xt, yt, xv, yv = (t.to(dev, dtype) for t in (x_tr, y_tr, x_val, y_val))
    if EMB:
        embeddings, _ = mmaml_oracle.conv_embedding(pe, xt)
    else:
        embeddings = [e.to(dev, dtype).requires_grad_() for e in const_emb]
    params = dict(pm)
    for _ in range(STEPS):
        loss = mse(mmaml_oracle.gated_conv(params, xt, embeddings), yt)
        grads = torch.autograd.grad(loss, list(params.values()), create_graph=True, allow_unused=True)
        params = {n: (p if gi is None else p - LR * gi.clamp(-20, 20)) for (n, p), gi in zip(params.items(), grads)}
    outer = mse(mmaml_oracle.gated_conv(params, xv, embeddings), yv)
    outer.backward()
    out = {f"model.{k}": p.grad.double().cpu() for k, p in pm.items()}
    if EMB:
        out.update({f"emb.{k}": p.grad.double().cpu() for k, p in pe.items()})
    else:
        out.update({f"film{i}": e.grad.double().cpu() for i, e in enumerate(embeddings)})
    return float(outer), out


torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
o1, g1 = ours()
o2, g2 = truth()
o3, g3 = truth(torch.float32, "cuda")       # the reference formulation in fp32 on this GPU (no TF32)
o4, g4 = truth(torch.float32, "cpu")
rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-30))
print(f"STEPS={STEPS} LR={LR} EMB={EMB}: outer ours {o1:.6f} truth(gpu fp64) {o2:.6f} torch gpu fp32 {o3:.6f} torch cpu fp32 {o4:.6f}")
big = max(float(v.norm()) for v in g2.values())
for k in g2:
    if float(g2[k].norm()) > 1e-9 * big:
        print(f"   {k:40s} ours {rel(g1[k], g2[k]):.2e}   torch-gpu-fp32 {rel(g3[k], g2[k]):.2e}   torch-cpu-fp32 {rel(g4[k], g2[k]):.2e}  |g| {float(g2[k].norm()):.3e}")

# ---- functional re-run of "ours" (the Functions called directly, no module plumbing), constant FiLM inputs
if not EMB:
    from b200np import lib
    from b200np.engine import AggregateFn, LinearFn
    from b200np.mmaml import BnActFn, Conv3x3S2Fn
    P = engine.PRECISION
    RUNSTATS = int(os.environ.get("RUNSTATS", 0))

    def fnet(x, ps, films):
        h = x.reshape(x.shape[0], x.shape[2], x.shape[3], 1)
        for i in range(4):
            C = ps[2 * i].shape[0]
            e = films[i].reshape(-1)
            rm = torch.zeros(C, device="cuda") if RUNSTATS else None
            rv = torch.ones(C, device="cuda") if RUNSTATS else None
            h = BnActFn.apply(Conv3x3S2Fn.apply(P, h, ps[2 * i], ps[2 * i + 1]), e[:C], e[C:], 1.0, True, 1e-5, rm, rv, 0.1)
        n = h.shape[0]
        feat = AggregateFn.apply(0, h.reshape(n, -1, h.shape[-1]))
        return LinearFn.apply(lib.ACT_TANH, P, feat, None, ps[8], ps[9])

    names = [k for k, _ in model.named_parameters()]
    ps = [p.detach().clone().requires_grad_() for _, p in model.named_parameters()]
    films = [e.cuda().requires_grad_() for e in const_emb]
    cur = list(ps)
    for _ in range(STEPS):
        l = lossf.calc_loss(fnet(x_tr, cur, films), None, y_tr)
        gr = torch.autograd.grad(l, cur, create_graph=True, allow_unused=True)
        cur = [p if gi is None else p - LR * gi.clamp(-20, 20) for p, gi in zip(cur, gr)]
    o = lossf.calc_loss(fnet(x_val, cur, films), None, y_val)
    o.backward()
    print(f"functional (RUNSTATS={RUNSTATS}): outer {float(o):.6f}")
    for k, p in zip(names, ps):
        t = g2[f"model.{k}"]
        if float(t.norm()) > 1e-9 * big:
            print(f"   {k:40s} functional vs gpu64 {rel(p.grad.double().cpu(), t):.2e}")
