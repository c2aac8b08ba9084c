#!/usr/bin/env python
"""Golden vectors of one SECOND-ORDER MMAML meta-step on the reference's own modules (build container only)
-> tests/golden/golden_mmaml2_v1.npz.

What trainer/meta_learner_reg.py does per task (adapt :129-160, update_params :113-127 with first_order=False and
inner_loop_grad_clip=20 as train.py:97-104 builds it, step :170-186), on the UNMODIFIED networks/gated_conv_net.py and
networks/conv_embedding_model.py built like networks/MMAMLShapeNet1D.py:31-75 (seed 2578):

    embeddings = embedding_model(x_train)
    params = model.param_dict
    twice:  loss = azimuth_mse(model(x_train, params, embeddings), y_train)
            grads = autograd.grad(loss, params, create_graph=True); params = params - fast_lr * clamp(grads, +-20)
    outer = azimuth_mse(model(x_val, params, embeddings), y_val);  outer.backward()

Stored: the inner losses, the outer loss and predictions, and the outer gradient of every parameter of both nets
(fingerprints; the tensors themselves up to 1024 elements), plus the same gradients with create_graph=False so that the
tests can tell second order from first order.
"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, HERE]
from oracle import ref_shims, synth  # noqa: E402
from make_golden_mmaml import build, fingerprint  # noqa: E402

N_IMG, SEED_TRAIN, SEED_VAL, FAST_LR, CLIP, STEPS = 15, 31, 32, 0.05, 20.0, 2


def mse(pred, y):
    return torch.mean(torch.sum((y[..., :2] - pred) ** 2, dim=-1))     # trainer/losses.py:59-61


def meta_step(model, emb, x_tr, y_tr, x_val, y_val, second_order):
    embeddings = emb(x_tr)
    params = OrderedDict(model.named_parameters())
    inner = []
    for _ in range(STEPS):
        loss = mse(model(x_tr, params=params, embeddings=embeddings), y_tr)
        grads = torch.autograd.grad(loss, list(params.values()), create_graph=second_order, allow_unused=True)
        params = OrderedDict((n, p if g is None else p - FAST_LR * g.clamp(min=-CLIP, max=CLIP))
                             for (n, p), g in zip(params.items(), grads))
        inner.append(loss.item())
    pred = model(x_val, params=params, embeddings=embeddings)
    outer = mse(pred, y_val)
    outer.backward()
    return inner, outer, pred


def batches():
    cx, cy, _, _ = synth.task_batch("shapenet_1d", 1, N_IMG, 1, seed=SEED_TRAIN)
    vx, vy, _, _ = synth.task_batch("shapenet_1d", 1, N_IMG, 1, seed=SEED_VAL)
    return (torch.from_numpy(a[0]) for a in (cx, cy, vx, vy))


def main():
    assert ref_shims.reference_available()
    out = {}
    for tag, second in (("so", True), ("fo", False)):
        model, emb = build()
        x_tr, y_tr, x_val, y_val = batches()
        inner, outer, pred = meta_step(model, emb, x_tr, y_tr, x_val, y_val, second)
        out[f"{tag}/inner_losses"] = np.array(inner)
        out[f"{tag}/outer_loss"] = np.array(outer.item())
        out[f"{tag}/pred"] = pred.detach().numpy()
        for net, m in (("model", model), ("emb", emb)):
            names, fps = [], []
            for k, p in m.named_parameters():
                if p.grad is not None:
                    names.append(k)
                    fps.append(fingerprint(p.grad))
                    if p.numel() <= 1024:
                        out[f"{tag}/{net}/grad/{k}"] = p.grad.numpy().copy()
            out[f"{tag}/{net}/grad_keys"] = np.array(names)
            out[f"{tag}/{net}/grad_fp"] = np.stack(fps)
    path = os.path.join(HERE, "golden_mmaml2_v1.npz")
    np.savez_compressed(path, **out)
    so, fo = out["so/model/grad_fp"][:, 2], out["fo/model/grad_fp"][:, 2]
    print("wrote", path, os.path.getsize(path), "outer", out["so/outer_loss"], "inner", out["so/inner_losses"])
    print("model grad norms second / first order:", np.round(so, 5), np.round(fo, 5))


if __name__ == "__main__":
    main()
