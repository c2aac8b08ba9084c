"""Task sampler: the numpy oracle against the reference's own batching code (golden), and the
device-side sampler against the oracle -- bit-exact, images and labels."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sampler_golden.npz")


def _oracle_batch(seed, T, shot, task_aug, eval_mode):
    bank, centers, item_indices = sampler.synthetic_bank()
    np.random.seed(seed)
    ctx, tgt = sampler.draw_task_indices(item_indices, 36, T, shot, "eval" if eval_mode else "train")
    ys, yq = sampler.gather_labels(centers, ctx, tgt, task_aug)
    return sampler.gather_images(bank, ctx), sampler.gather_images(bank, tgt), ys, yq


@pytest.mark.parametrize("name", ["train_aug", "eval"])
def test_oracle_sampler_matches_reference_golden(name):
    g = np.load(GOLD)
    seed, T, shot, task_aug, eval_mode = (int(v) for v in g[f"{name}/cfg"])
    xs, xq, ys, yq = _oracle_batch(seed, T, shot, bool(task_aug), bool(eval_mode))
    for k, v in (("xs", xs), ("xq", xq), ("ys", ys), ("yq", yq)):
        assert v.dtype == np.float32 and np.array_equal(v, g[f"{name}/{k}"]), (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["train_aug", "eval"])
def test_device_sampler_bit_exact(name):
    from b200np.data import DeviceTaskSampler
    g = np.load(GOLD)
    seed, T, shot, task_aug, eval_mode = (int(v) for v in g[f"{name}/cfg"])
    bank, centers, item_indices = sampler.synthetic_bank()
    ds = DeviceTaskSampler(bank, centers, item_indices, 36, mode="eval" if eval_mode else "train", task_aug=bool(task_aug))
    np.random.seed(seed)
    got = ds.get_batch(T, shot)
    for k, v in zip(("xs", "xq", "ys", "yq"), got):
        assert v.is_cuda and v.dtype == torch.float32
        assert np.array_equal(v.cpu().numpy(), g[f"{name}/{k}"]), (name, k)


@pytest.mark.gpu
def test_gather_images_u8_rgb_and_full_size():
    """3-channel banks (ShapeNet3D-style 64x64x3) and the 128x128x1 bench shape against the numpy oracle."""
    from b200np import ops
    for (n, H, W, C, M) in ((40, 64, 64, 3, 33), (50, 128, 128, 1, 720)):
        rng = np.random.RandomState(3)
        bank = rng.randint(0, 256, size=(n, H, W, C), dtype=np.uint8)
        rows = rng.randint(0, n, size=(1, M))
        ref = sampler.gather_images(bank, rows)
        got = ops.gather_images_u8(torch.from_numpy(bank).cuda(), torch.from_numpy(rows).to(torch.int32).cuda())
        assert np.array_equal(got.cpu().numpy(), ref)
