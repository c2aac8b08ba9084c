"""CPU restatement (numpy) of the reference's task sampler for the distractor dataset -- TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py).  Follows dataset/shapenet_distractor.py:

* `draw_task_indices`  : `__generateRandomTask` (:261-299) -- the numpy RNG call sequence that picks an item, permutes
  its instances, splits them into context / target and shuffles both (`shuffle_batch`, :33-38); returns ROW INDICES
  into the image bank instead of the images themselves.
* `gather_images`      : `__yield_random_task_batch` (:233-236,256-259) + `utils/utils.py:26-30`:
  `255 - uint8`, `astype(float32) / 255.0`, channel-last -> `[T, n, C, H, W]`.
* `gather_labels`      : the centre labels with the task-augmentation shift (:247-254).

Pinned by tests/golden/sampler_golden.npz, produced by calling the reference's own private methods on a synthetic
image bank (tests/golden/make_sampler_golden.py).
"""
import numpy as np


def draw_task_indices(item_indices, instances_per_item, T, shot, mode="train", rng=np.random):
    """-> (ctx_rows [T, shot], tgt_rows [T, nt]) int64 row indices into the bank; consumes `rng` exactly like the
    reference does for source in ('train', 'validation')  (shapenet_distractor.py:272-299)."""
    ctx, tgt = [], []
    items = np.unique(item_indices)
    for _ in range(T):
        task_item = rng.choice(items)                                   # :275
        permutation = rng.permutation(instances_per_item)               # :276
        rows = np.where(item_indices == task_item)[0][permutation]      # :282
        train = rows[:shot]                                             # :285
        test = rows if mode == "eval" else rows[shot:]                  # :287-292
        train = train[rng.permutation(train.shape[0])]                  # shuffle_batch :37
        test = test[rng.permutation(test.shape[0])]
        ctx.append(train)
        tgt.append(test)
    return np.array(ctx), np.array(tgt)


def gather_images(bank_u8, rows):
    """bank_u8 [n, H, W, C] uint8, rows [T, k] -> float32 [T, k, C, H, W]  (:233-234, :256-259, utils.py:26-30)."""
    x = 255 - bank_u8[rows]                       # uint8 arithmetic, as in the reference
    x = x.astype(np.float32) / 255.0
    return np.ascontiguousarray(np.transpose(x, (0, 1, 4, 2, 3)))


def gather_labels(centers, ctx_rows, tgt_rows, task_aug=False, num_noise=16, rng=np.random):
    """-> (ys [T, nc, 2], yq [T, nt, 2]) float32  (:235-236, :247-254, :261)."""
    ys, yq = np.array(centers[ctx_rows]), np.array(centers[tgt_rows])
    if task_aug:
        noise = np.linspace(0, 16, num_noise + 1)[:-1]
        y_noise = rng.choice(noise, (ctx_rows.shape[0], 2))[:, None, :]
        ys = ys + y_noise
        yq = yq + y_noise
        ys %= 128
        yq %= 128
    return ys.astype(np.float32), yq.astype(np.float32)


def synthetic_bank(n_items=6, instances_per_item=36, H=16, W=16, C=1, seed=11):
    """Deterministic uint8 image bank + centres + item indices (integer hash, no RNG state involved)."""
    from . import synth
    n = n_items * instances_per_item
    bank = (synth.hash_u32((n, H, W, C), seed) >> np.uint32(24)).astype(np.uint8)
    centers = (synth.hash_u32((n, 2), seed + 1) % np.uint32(128)).astype(np.float64)
    item_indices = np.repeat(np.arange(n_items), instances_per_item)
    return bank, centers, item_indices
