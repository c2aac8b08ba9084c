"""Device-side task sampler for the distractor data.

The reference assembles every batch on the host -- a python loop with `np.where` over the whole index array per task,
uint8 -> float conversion and a channel permute (dataset/shapenet_distractor.py:209-299) -- and ships 47 MB of fp32
images to the GPU per step.  Here the uint8 image bank lives in HBM; the host only draws the row indices (the same
numpy RNG call sequence as the reference, so a seeded run picks the same tasks) and the labels, a few KB per step,
and `b200np_gather_images_u8` builds the fp32 NCHW batch on the device.
"""
import numpy as np
import torch

from . import ops


class DeviceTaskSampler:
    def __init__(self, images_u8, centers, item_indices, instances_per_item, mode="train", task_aug=False,
                 num_noise=16, device="cuda"):
        """images_u8 [n,H,W,C] uint8 (numpy or tensor), centers [n,2], item_indices [n] -- the arrays
        `ShapeNetDistractor.__extract_data` produces (shapenet_distractor.py:301-317)."""
        self.bank = torch.as_tensor(images_u8, dtype=torch.uint8).contiguous().to(device)
        self.centers = np.asarray(centers, dtype=np.float64)
        self.item_indices = np.asarray(item_indices)
        self.items = np.unique(self.item_indices)
        self.instances_per_item, self.mode = instances_per_item, mode
        self.task_aug, self.num_noise = task_aug, num_noise
        self.device = device

    def draw(self, tasks_per_batch, shot, rng=np.random):
        """Row indices of one batch: `__generateRandomTask` (shapenet_distractor.py:272-299), same RNG calls."""
        ctx, tgt = [], []
        for _ in range(tasks_per_batch):
            task_item = rng.choice(self.items)
            permutation = rng.permutation(self.instances_per_item)
            rows = np.where(self.item_indices == task_item)[0][permutation]
            train = rows[:shot]
            test = rows if self.mode == "eval" else rows[shot:]
            train = train[rng.permutation(train.shape[0])]
            test = test[rng.permutation(test.shape[0])]
            ctx.append(train)
            tgt.append(test)
        return np.array(ctx), np.array(tgt)

    def labels(self, ctx_rows, tgt_rows, rng=np.random):
        """Centre labels with the task-augmentation shift (shapenet_distractor.py:235-236,247-254)."""
        ys, yq = np.array(self.centers[ctx_rows]), np.array(self.centers[tgt_rows])
        if self.task_aug and self.mode == "train":
            noise = np.linspace(0, 16, self.num_noise + 1)[:-1]
            y_noise = rng.choice(noise, (ctx_rows.shape[0], 2))[:, None, :]
            ys = ys + y_noise
            yq = yq + y_noise
            ys %= 128
            yq %= 128
        return ys.astype(np.float32), yq.astype(np.float32)

    def get_batch(self, tasks_per_batch, shot, rng=np.random):
        """-> (xs [T,nc,C,H,W], xq [T,nt,C,H,W], ys [T,nc,2], yq [T,nt,2]) CUDA fp32 tensors."""
        ctx_rows, tgt_rows = self.draw(tasks_per_batch, shot, rng)
        ys, yq = self.labels(ctx_rows, tgt_rows, rng)
        to_dev = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).pin_memory().to(self.device, non_blocking=True)
        xs = ops.gather_images_u8(self.bank, to_dev(ctx_rows, torch.int32))
        xq = ops.gather_images_u8(self.bank, to_dev(tgt_rows, torch.int32))
        return xs, xq, to_dev(ys, torch.float32), to_dev(yq, torch.float32)
