import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "what-matters-for-meta-learning_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


# name: (method, task, agg_mode, img_agg, extra cfg, T, nc, nt) -- mirrors tests/golden/make_golden.py
CASES = {
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


def make_config(method, task, T, agg_mode, img_agg, device="cpu", seed=2578, **extra):
    """Namespace with the attributes configs/config.py:33-104 would set."""
    img_size, input_dim, output_dim = {
        "shapenet_3d": ([64, 64, 4], 4, 4), "shapenet_1d": ([128, 128, 1], 3, 2),
        "distractor": ([128, 128, 1], 2, 2)}[task]
    base = dict(method=method, task=task, tasks_per_batch=T, agg_mode=agg_mode, img_agg=img_agg,
                dim_w=None, dim_r=None, dim_z=None, n_hidden_units_r=None, seed=seed, device=device,
                temperature=0.07, img_size=img_size, input_dim=input_dim, output_dim=output_dim,
                loss_type="mse", beta=0, contrastive=False)
    base.update(extra)
    return types.SimpleNamespace(**base)


def build_product_model(case, device="cpu", T=None):
    import importlib
    method, task, agg, img_agg, extra, T0, nc, nt = CASES[case]
    cfg = make_config(method, task, T or T0, agg, img_agg, device=device, **extra)
    cls = getattr(importlib.import_module("networks." + method), method)
    return cls(cfg), cfg


def oracle_cfg(cfg):
    return dict(tasks_per_batch=cfg.tasks_per_batch, agg_mode=cfg.agg_mode, img_agg=cfg.img_agg,
                task=cfg.task, dim_w=cfg.dim_w, dim_r=cfg.dim_r, dim_z=cfg.dim_z)


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), allow_pickle=False)


def probe_vec(n, seed=7):
    from oracle import synth
    return (synth.hash_u32((n,), seed) & np.uint32(1)).astype(np.float64) * 2.0 - 1.0


def fingerprint(t):
    a = t.detach().double().reshape(-1).cpu().numpy()
    return np.array([a.sum(), np.abs(a).sum(), np.sqrt((a * a).sum()), float(a @ probe_vec(a.size))])


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
