#!/usr/bin/env python
"""Generate tests/golden/golden_v1.npz by running the UNMODIFIED reference (build container only).

    python tests/golden/make_golden.py

The reference (/root/reference) is imported through oracle/ref_shims.py (four import stubs, no
arithmetic touched), its own nn.Modules are constructed with the shipped seed (2578) and run on
integer-hash synthetic task batches (oracle/synth.py), and for every case we record

* a fingerprint of every state_dict entry right after construction (init / RNG-order contract),
* the prediction ``mu``, the loss from the reference's own ``LossFunc``,
* a fingerprint of every parameter gradient (+ a few small gradients in full),
* hooked intermediates: encoder features, context features, attention output, and the integer
  argmax of the adaptive max-pool and of the CNP max aggregation.

Inputs are not stored: ``oracle.synth`` regenerates them bit-exactly anywhere.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shims, synth  # noqa: E402

CASES = {
    # name: (method, task, agg_mode, img_agg, extra cfg, T, nc, nt)
    "anp_distractor": ("ANPDistractor", "distractor", "attention", "max", dict(dim_w=16), 2, 3, 4),
    "anp_distractor_nc0": ("ANPDistractor", "distractor", "attention", "max", dict(dim_w=16), 2, 0, 3),
    "cnp_distractor_max": ("CNPDistractor", "distractor", "max", "max", dict(dim_w=16), 2, 3, 4),
    "cnp_distractor_mean": ("CNPDistractor", "distractor", "mean", "max", dict(dim_w=16), 2, 3, 4),
    "cnp_distractor_baco": ("CNPDistractor", "distractor", "baco", "max", dict(dim_w=16), 2, 3, 4),
    "anp_3d": ("ANP", "shapenet_3d", "attention", "reshape", dict(), 2, 3, 2),
    "cnp_3d_max": ("CondNeuralProcess", "shapenet_3d", "max", "reshape", dict(), 2, 3, 2),
    "cnp_1d_mean": ("CNPShapeNet1D", "shapenet_1d", "mean", "",
                    dict(dim_w=64, dim_r=100, dim_z=64, n_hidden_units_r=[100, 100]), 2, 3, 3),
    "cnp_1d_max": ("CNPShapeNet1D", "shapenet_1d", "max", "",
                   dict(dim_w=64, dim_r=100, dim_z=64, n_hidden_units_r=[100, 100]), 2, 3, 3),
    "anp_1d": ("ANPShapeNet1D", "shapenet_1d", "attention", "",
               dict(dim_w=64, dim_r=64, dim_z=64, n_hidden_units_r=[100, 100]), 2, 3, 3),
}


def probe_vec(n, seed=7):
    """fixed +-1 probe vector for a scalar fingerprint <t, probe>."""
    return (synth.hash_u32((n,), seed) & np.uint32(1)).astype(np.float64) * 2.0 - 1.0


def fingerprint(t):
    a = t.detach().double().reshape(-1).numpy()
    return np.array([a.sum(), np.abs(a).sum(), np.sqrt((a * a).sum()), float(a @ probe_vec(a.size))])


def run_case(name, out):
    method, task, agg, img_agg, extra, T, nc, nt = CASES[name]
    cfg = ref_shims.make_config(method, task, T, agg, img_agg, **extra)
    model = ref_shims.reference_class(method)(cfg)
    lossf = ref_shims.reference_lossfunc()("mse", task)
    sd = model.state_dict()
    out[f"{name}/keys"] = np.array(list(sd.keys()))
    out[f"{name}/init_fp"] = np.stack([fingerprint(v) for v in sd.values()])
    out[f"{name}/shapes"] = np.array([str(tuple(v.shape)) for v in sd.values()])

    inter = {}

    def grab(key):
        def hook(_m, _i, o):
            inter.setdefault(key, []).append(o.detach().clone())
        return hook

    hooks = []
    if hasattr(model, "img_encoder"):
        hooks.append(model.img_encoder.register_forward_hook(grab("enc_feat")))

        def pool_hook(_m, i, _o):
            _, idx = torch.nn.functional.adaptive_max_pool2d(i[0], (2, 2), return_indices=True)
            inter.setdefault("pool_idx", []).append(idx.detach().clone())
        hooks.append(model.img_encoder.resnet.adaptmax.register_forward_hook(pool_hook))
        hooks.append(model.task_encoder.register_forward_hook(grab("ctx_feat")))
    if hasattr(model, "encoder_w0"):
        hooks.append(model.encoder_w0.register_forward_hook(grab("enc_feat")))
        hooks.append(model.encoder_r.register_forward_hook(grab("ctx_feat")))
    if hasattr(model, "attn"):
        hooks.append(model.attn.register_forward_hook(grab("attn_out")))

    cx, cy, tx, ty = (torch.from_numpy(a) for a in synth.task_batch(task, T, nc, nt, seed=11))
    model.train()
    mu, var, kl = model(cx, cy, tx)
    assert var is None and kl == 0
    loss = lossf.calc_loss(mu, None, ty)
    loss.backward()
    for h in hooks:
        h.remove()

    out[f"{name}/mu"] = mu.detach().numpy()
    out[f"{name}/loss"] = np.array(loss.item(), dtype=np.float64)
    names, fps = [], []
    for k, p in model.named_parameters():
        if p.grad is not None:
            names.append(k)
            fps.append(fingerprint(p.grad))
            if p.numel() <= 512:
                out[f"{name}/grad/{k}"] = p.grad.numpy().copy()
    out[f"{name}/grad_keys"] = np.array(names)
    out[f"{name}/grad_fp"] = np.stack(fps)
    for k, lst in inter.items():
        for i, t in enumerate(lst):
            out[f"{name}/inter/{k}{i}"] = t.numpy()
    if nc and agg == "max":
        out[f"{name}/inter/agg_idx"] = inter["ctx_feat"][0].max(dim=1).indices.numpy()
    # eval-mode loss for the ShapeNet1D family (degree error, losses.py:63-76)
    if task == "shapenet_1d":
        out[f"{name}/loss_test"] = np.array(lossf.calc_loss(mu.detach(), None, ty, test=True).item())
    print(f"{name}: loss={loss.item():.6f} params={len(sd)} grads={len(names)}")


def loss_cases(out):
    LossFunc = ref_shims.reference_lossfunc()
    for task in ("distractor", "shapenet_3d", "shapenet_1d"):
        _, _, _, L, O = synth.TASKS[task]
        y = torch.from_numpy(synth.labels(task, (3, 5), 101))
        mu = torch.from_numpy(synth.uniform((3, 5, O), 102, -1.0, 1.0)).requires_grad_(True)
        if task == "distractor":
            mu = (mu.detach() * 64 + 64).requires_grad_(True)
        lf = LossFunc("mse", task)
        loss = lf.calc_loss(mu, None, y)
        loss.backward()
        out[f"loss/{task}/mu"] = mu.detach().numpy()
        out[f"loss/{task}/y"] = y.numpy()
        out[f"loss/{task}/loss"] = np.array(loss.item())
        out[f"loss/{task}/dmu"] = mu.grad.numpy()
        if task == "shapenet_1d":
            out[f"loss/{task}/loss_test"] = np.array(lf.calc_loss(mu.detach(), None, y, test=True).item())


def main():
    assert ref_shims.reference_available(), "needs /root/reference (build container only)"
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out = {}
    for name in CASES:
        run_case(name, out)
    loss_cases(out)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
