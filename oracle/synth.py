"""Synthetic task batches from pure integer arithmetic (TEST INFRASTRUCTURE).

The reference's datasets are Git-LFS blobs that are absent, so every test and
benchmark uses synthetic batches of the shapes the datasets produce:

* images  ``uint8/255`` fp32 in [0,1]  (dataset/shapenet_distractor.py:233-234,256-257)
  Distractor / ShapeNet1D ``[T,n,1,128,128]``, ShapeNet3D ``[T,n,3,64,64]``
* labels  Distractor: pixel centres in [0,128)   (dataset/shapenet_distractor.py:253-254)
          ShapeNet1D: ``[cos a, sin a, a]``       (dataset/shapenet_1d.py:171-193)
          ShapeNet3D: unit quaternions            (dataset/shapenet_3d.py:225-227)

Everything is derived from a 32-bit integer hash, so the same arrays come out on every host,
numpy version and torch version -- golden vectors do not need to store their inputs.
"""
import numpy as np

_M32 = np.uint64(0xFFFFFFFF)


def _mix(x):
    """splitmix-style 32-bit finaliser on uint64 lanes (wraps mod 2^32 explicitly)."""
    x = x & _M32
    x = ((x ^ (x >> np.uint64(16))) * np.uint64(0x7FEB352D)) & _M32
    x = ((x ^ (x >> np.uint64(15))) * np.uint64(0x846CA68B)) & _M32
    x = x ^ (x >> np.uint64(16))
    return x & _M32


def hash_u32(shape, seed):
    """uint32 array of ``shape``; element i = mix(mix(i + seed*0x9E3779B9))."""
    n = int(np.prod(shape)) if len(shape) else 1
    idx = np.arange(n, dtype=np.uint64)
    x = _mix(idx + np.uint64((seed * 0x9E3779B9) & 0xFFFFFFFF))
    x = _mix(x ^ np.uint64(seed & 0xFFFFFFFF))
    return x.astype(np.uint32).reshape(shape)


def images(shape, seed):
    """fp32 images = uint8/255, like the datasets deliver them."""
    u8 = (hash_u32(shape, seed) >> np.uint32(24)).astype(np.float32)
    return (u8 / np.float32(255.0)).astype(np.float32)


def uniform(shape, seed, lo=0.0, hi=1.0):
    """fp32 uniform in [lo,hi) with 24 random bits (exactly representable steps)."""
    u = (hash_u32(shape, seed) >> np.uint32(8)).astype(np.float64) / float(1 << 24)
    return (lo + (hi - lo) * u).astype(np.float32)


TASKS = {
    # task: (C, H, W, label_dim, out_dim)   -- configs/config.py:87-104
    "distractor": (1, 128, 128, 2, 2),
    "shapenet_1d": (1, 128, 128, 3, 2),
    "shapenet_3d": (3, 64, 64, 4, 4),
}


def labels(task, shape_tn, seed):
    T, n = shape_tn
    if task == "distractor":
        return uniform((T, n, 2), seed, 0.0, 128.0)
    if task == "shapenet_1d":
        a = uniform((T, n), seed, 0.0, 2.0 * np.pi).astype(np.float64)
        return np.stack([np.cos(a), np.sin(a), a], axis=-1).astype(np.float32)
    if task == "shapenet_3d":
        q = uniform((T, n, 4), seed, -1.0, 1.0).astype(np.float64)
        q[..., 1] = np.abs(q[..., 1]) + 1e-3
        q = q / np.linalg.norm(q, axis=-1, keepdims=True)
        return q.astype(np.float32)
    raise ValueError(task)


def task_batch(task, T, nc, nt, seed=0):
    """(ctx_x, ctx_y, tgt_x, tgt_y) numpy fp32, shapes as dataset.get_batch returns them
    (trainer/model_trainer.py:64-70)."""
    C, H, W, _, _ = TASKS[task]
    ctx_x = images((T, nc, C, H, W), seed * 4 + 1)
    tgt_x = images((T, nt, C, H, W), seed * 4 + 2)
    ctx_y = labels(task, (T, nc), seed * 4 + 3)
    tgt_y = labels(task, (T, nt), seed * 4 + 4)
    return ctx_x, ctx_y, tgt_x, tgt_y
