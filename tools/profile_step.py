#!/usr/bin/env python
"""Per-kernel GPU time of the eager step from CUPTI (torch.profiler): warm caches, real concurrency --
the complement of the ncu launch list (cold-cache, serialised).  Prints ms per step per kernel name and,
with --shapes, the time of the library GEMMs broken down by (M, N, K, roles)."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
import torch
from torch.profiler import ProfilerActivity, profile
import importlib
from bench import MODELS, make_cfg
from b200np import engine
from b200np.optim import FlatParams, FusedAdam
from oracle import synth
from trainer.losses import LossFunc

engine.set_precision(sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "tf32x3")
MODEL = os.environ.get("PROFILE_MODEL", "ANPDistractor")      # any key of bench.MODELS
task, _, _, _, T, views = MODELS[MODEL]
NC = 15
NT = views - NC if task == "distractor" else 15
model = getattr(importlib.import_module(f"networks.{MODEL}"), MODEL)(make_cfg(T, "cuda:0", MODEL)).to("cuda:0")
flat = FlatParams(model); opt = FusedAdam(flat, lr=1e-4); lossf = LossFunc("mse", task)
b = [torch.from_numpy(a).cuda() for a in synth.task_batch(task, T, NC, NT, seed=1)]
print(f"# {MODEL} T={T} nc={NC} nt={NT}")


def step():
    opt.zero_grad(); mu, _, _ = model(b[0], b[1], b[2]); loss = lossf.calc_loss(mu, None, b[3]); loss.backward(); opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
n = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n):
        step()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        k = ev.name.replace("b200np::<unnamed>::", "").replace("(anonymous namespace)::", "").replace("void ", "")[:80]
        agg[k][0] += 1
        agg[k][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
tot = sum(v for _, v in agg.values())
print(f"kernel time {tot / n / 1e3:.3f} ms/step over {sum(c for c, _ in agg.values()) / n:.0f} launches/step")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v / n / 1e3:9.3f} ms {100 * v / tot:5.1f}% {c / n:7.1f}/step {v / c:8.1f} us  {k}")
