"""Multi-GPU plumbing: one process per GPU, meta-batch sharded by task, NCCL over NVLink.

Tasks are independent, so rank r owns tasks [r*T/W, (r+1)*T/W) and the only
data-path exchanges are
  * one SUM all-reduce of the flat gradient buffer per step (outer-loop gradient), and
  * for ANP models two scalar all-reduces: MAX of the FAVOR+ key stabiliser in forward
    (networks/fast_attention.py:97 takes the max over the WHOLE meta-batch) and SUM of its gradient
    (+ tie count) in backward.
With world size 1 (or torch.distributed not initialised) every function is a no-op.
"""
import torch
import torch.distributed as td


def initialised():
    return td.is_available() and td.is_initialized()


def world_size():
    return td.get_world_size() if initialised() else 1


def rank():
    return td.get_rank() if initialised() else 0


def all_reduce_max(t):
    if world_size() > 1:
        td.all_reduce(t, op=td.ReduceOp.MAX)
    return t


def all_reduce_sum(t):
    if world_size() > 1:
        td.all_reduce(t, op=td.ReduceOp.SUM)
    return t


def shard_tasks(total_tasks, world=None, r=None):
    """Contiguous task range [lo, hi) of rank r; requires total_tasks % world == 0 because every
    rank must build its model with the same local tasks_per_batch (networks/models.py:115)."""
    world = world_size() if world is None else world
    r = rank() if r is None else r
    if total_tasks % world:
        raise ValueError(f"tasks_per_batch={total_tasks} is not divisible by world size {world}")
    per = total_tasks // world
    return r * per, (r + 1) * per


def local_config(config, world=None):
    """Copy of `config` with tasks_per_batch replaced by the per-rank task count."""
    import copy
    world = world_size() if world is None else world
    lo, hi = shard_tasks(config.tasks_per_batch, world, 0)
    c = copy.copy(config)
    c.tasks_per_batch = hi - lo
    return c


def all_reduce_grads(flat_grad):
    """SUM all-reduce of the flat gradient; the 1/world scaling is folded into FusedAdam's
    grad_scale (the loss is a mean over equal-sized task groups)."""
    return all_reduce_sum(flat_grad)
