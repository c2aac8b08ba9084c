#!/usr/bin/env python
"""Generate tests/golden/sampler_golden.npz by calling the UNMODIFIED reference sampler (build container only).

The reference's `ShapeNetDistractor` loads pickled data that is not shipped; its batching code does not depend on
that, so the object is created without `__init__`, given a synthetic uint8 image bank (oracle.sampler.synthetic_bank)
and its own private `__yield_random_task_batch` is called with a seeded numpy RNG.  Stored: the four returned tensors
for a train-mode batch (with task augmentation) and an eval-mode batch.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims, sampler  # noqa: E402

ref_shims.install()
sys.path.insert(0, "/root/reference")
from dataset.shapenet_distractor import ShapeNetDistractor  # noqa: E402

out = {}
bank, centers, item_indices = sampler.synthetic_bank()
for name, mode, task_aug, seed, T, shot in (("train_aug", "train", True, 5, 3, 4), ("eval", "eval", False, 6, 2, 5)):
    ds = object.__new__(ShapeNetDistractor)
    ds.mode, ds.source = mode, "train" if mode == "train" else "validation"
    ds.data_aug, ds.task_aug, ds.num_noise = False, task_aug, 16
    ds.instances_per_item = 36
    np.random.seed(seed)
    xs, xq, ys, yq = ds._ShapeNetDistractor__yield_random_task_batch(T, bank, None, item_indices, centers.copy(), shot)
    for k, v in (("xs", xs), ("xq", xq), ("ys", ys), ("yq", yq)):
        out[f"{name}/{k}"] = v.numpy()
    out[f"{name}/cfg"] = np.array([seed, T, shot, int(task_aug), int(mode == "eval")])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sampler_golden.npz"), **out)
print({k: v.shape for k, v in out.items()})
