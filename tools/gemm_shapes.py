#!/usr/bin/env python
"""Lists the dense-layer GEMMs of one ANPDistractor step (shape, operand roles, epilogue) and times each
distinct call alone -- where the ~60 small launches spend their time."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
import torch
from bench import NC, NT, TASKS_PER_GPU, make_cfg
from b200np import engine, ops
from b200np.optim import FlatParams, FusedAdam
from networks.ANPDistractor import ANPDistractor
from oracle import synth
from trainer.losses import LossFunc

engine.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
T = TASKS_PER_GPU
model = ANPDistractor(make_cfg(T, "cuda:0")).to("cuda:0")
flat = FlatParams(model); opt = FusedAdam(flat, lr=1e-4); lossf = LossFunc("mse", "distractor")
b = [torch.from_numpy(a).cuda() for a in synth.task_batch("distractor", T, NC, NT, seed=1)]


def step():
    opt.zero_grad(); mu, _, _ = model(b[0], b[1], b[2]); loss = lossf.calc_loss(mu, None, b[3]); loss.backward(); opt.step()


for _ in range(2):
    step()
calls = []
orig = ops.gemm


def logged(A, B, Cmat, M, N, K, a_rs, a_cs, b_rs, b_cs, ldc, **kw):
    g = len(A) if isinstance(A, (list, tuple)) else 1
    key = (M, N, K, g, "Ak" if a_cs == 1 else "Am", "Bk" if b_rs == 1 else "Bn", kw.get("beta", 0.0), kw.get("act", 0),
           bool(kw.get("bias") is not None), bool(kw.get("sum_groups", False)))
    calls.append((key, (A, B, Cmat, M, N, K, a_rs, a_cs, b_rs, b_cs, ldc), kw))
    return orig(A, B, Cmat, M, N, K, a_rs, a_cs, b_rs, b_cs, ldc, **kw)


ops.gemm = logged
engine.ops.gemm = logged
step()
torch.cuda.synchronize()
ops.gemm = orig
engine.ops.gemm = orig
count = collections.Counter(k for k, _, _ in calls)
seen = {}
tot = 0.0
for key, args, kw in calls:
    if key in seen:
        continue
    for _ in range(3):
        orig(*args, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        orig(*args, **kw)
    e1.record(); torch.cuda.synchronize()
    seen[key] = e0.elapsed_time(e1) / 20 * 1e3
for key, us in sorted(seen.items(), key=lambda kv: -kv[1] * count[kv[0]]):
    tot += us * count[key]
    print(f"{count[key]:3d} x {us:7.1f} us  M={key[0]:5d} N={key[1]:5d} K={key[2]:5d} groups={key[3]} {key[4]} {key[5]} beta={key[6]} act={key[7]} bias={key[8]} sum={key[9]}")
print(f"{len(calls)} GEMM calls per step, {tot / 1e3:.3f} ms when run back to back")
