#!/usr/bin/env python
"""Per layer of GatedConvModel: ReLU gates of the B200 forward vs fp64 on the SAME layer input (which op flips gates?)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200"), os.path.join(ROOT, "tests")]
os.environ["B200NP_MMAML"] = "1"
import torch
import torch.nn.functional as F
import test_mmaml as tm
from b200np import engine, ops
from b200np.mmaml import BnActFn, Conv3x3S2Fn

engine.set_precision(os.environ.get("PREC", "fp32"))
P = engine.PRECISION
model, emb = tm._build_models()
model, emb = model.cuda(), emb.cuda()
x_tr, _, _, _ = tm._meta_batches("cuda")
if os.environ.get("DATA") == "rand":
    x_tr = torch.rand(x_tr.shape, generator=torch.Generator().manual_seed(11)).cuda()
with torch.no_grad():
    embeddings = emb(x_tr)
    h = x_tr.reshape(x_tr.shape[0], x_tr.shape[2], x_tr.shape[3], 1)
    for i in range(1, 5):
        w, b = model.param_dict[f"features.layer{i}_conv.weight"], model.param_dict[f"features.layer{i}_conv.bias"]
        C = w.shape[0]
        e = embeddings[i - 1].reshape(-1)
        z = Conv3x3S2Fn.apply(P, h, w, b)
        y = BnActFn.apply(z, e[:C], e[C:], 1.0, True, 1e-5, None, None, 0.1)
        # fp64 on the same input h
        z64 = F.conv2d(h.double().permute(0, 3, 1, 2), w.double(), b.double(), stride=2, padding=1).permute(0, 2, 3, 1)
        def bn64(t):
            mu, var = t.mean((0, 1, 2)), t.var((0, 1, 2), unbiased=False)
            return (t - mu) * (var + 1e-5).rsqrt() * (1 + e[:C].double()) + e[C:].double()
        pre_a, pre_b = bn64(z64), bn64(z.double())          # exact conv + exact bn | our conv + exact bn
        f_total = int(((y > 0) != (pre_a > 0)).sum())
        f_bn_only = int(((y > 0) != (pre_b > 0)).sum())
        conv_err = float((z.double() - z64).norm() / z64.norm())
        ratio = float((z64.mean((0, 1, 2)).abs() / z64.std((0, 1, 2))).max())
        near = int((pre_a.abs() < 1e-5).sum())
        print(f"layer {i}: {y.numel():8d} gates, flipped vs fp64 {f_total} (norm+FiLM alone on our conv output: {f_bn_only}); "
              f"conv rel err {conv_err:.1e}; max |mean|/std of conv output {ratio:.1f}; |pre-activation| < 1e-5: {near}")
        h = y
