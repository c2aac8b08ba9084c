#!/usr/bin/env python
"""Timeline of the CAPTURED step (CUDA-graph replay) from CUPTI kernel timestamps: how much of the step's span has no
kernel running, which kernels run ALONE (nothing else resident on the GPU -- for a latency-bound kernel that is idle
silicon), and how much kernel time overlaps.  Complements tools/profile_step.py (per-kernel sums of the eager step).

    python tools/timeline_gaps.py [tf32x3|tf32|fp32] > gpurun_out/timeline.txt
"""
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from bench import NC, NT, TASKS_PER_GPU, make_cfg  # noqa: E402
from b200np import engine  # noqa: E402
from b200np.optim import FlatParams, FusedAdam, GraphedStep  # noqa: E402
from networks.ANPDistractor import ANPDistractor  # noqa: E402
from oracle import synth  # noqa: E402
from trainer.losses import LossFunc  # noqa: E402

engine.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
T = TASKS_PER_GPU
model = ANPDistractor(make_cfg(T, "cuda:0")).to("cuda:0")
opt = FusedAdam(FlatParams(model), lr=1e-4)
lossf = LossFunc("mse", "distractor")
b = [torch.from_numpy(a).cuda() for a in synth.task_batch("distractor", T, NC, NT, seed=1)]
gs = GraphedStep(model, lossf, opt, b, warmup=3)
for _ in range(3):
    gs(b)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gs(b)
    torch.cuda.synchronize()

ev = []
streams = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and "Memcpy" not in e.name and "Memset" not in e.name:
        t0 = e.time_range.start
        ev.append((t0, t0 + (e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total),
                   e.name.replace("b200np::<unnamed>::", "").replace("(anonymous namespace)::", "").replace("void ", "")[:70]))
        streams[(t0, ev[-1][2])] = getattr(e, "device_resource_id", -1)
ev.sort()
if os.environ.get("TIMELINE_TRACE"):   # chronological listing: start (us from the first kernel), duration, stream, name
    with open(os.environ["TIMELINE_TRACE"], "w") as fh:
        for s_, e_, k_ in ev:
            fh.write(f"{s_ - ev[0][0]:9.1f} {e_ - s_:8.1f} s{streams.get((s_, k_), -1)} {k_}\n")
t_begin, t_end = ev[0][0], max(e[1] for e in ev)
span = t_end - t_begin
# sweep: time with 0 / 1 / >= 2 kernels in flight; solo time per kernel name
points = sorted([(s, 1, i) for i, (s, _, _) in enumerate(ev)] + [(e, -1, i) for i, (_, e, _) in enumerate(ev)])
active, last = set(), t_begin
busy = [0.0, 0.0, 0.0]
solo = collections.defaultdict(float)
for t, d, i in points:
    dt = t - last
    if dt > 0:
        busy[min(len(active), 2)] += dt
        if len(active) == 1:
            solo[ev[next(iter(active))][2]] += dt
    last = t
    if d > 0:
        active.add(i)
    else:
        active.discard(i)
ksum = sum(e - s for s, e, _ in ev)
print(f"captured step: {len(ev)} kernels, span {span / 1e3:.3f} ms, sum of kernel durations {ksum / 1e3:.3f} ms")
print(f"  no kernel in flight   {busy[0] / 1e3:7.3f} ms ({100 * busy[0] / span:4.1f} %)")
print(f"  exactly one kernel    {busy[1] / 1e3:7.3f} ms ({100 * busy[1] / span:4.1f} %)")
print(f"  two or more kernels   {busy[2] / 1e3:7.3f} ms ({100 * busy[2] / span:4.1f} %)")
print("time spent as the ONLY kernel in flight, by kernel (latency-bound ones here are idle SMs):")
tot = collections.defaultdict(float)
cnt = collections.Counter()
for s, e, k in ev:
    tot[k] += e - s
    cnt[k] += 1
for k, v in sorted(solo.items(), key=lambda kv: -kv[1])[:24]:
    print(f"  {v / 1e3:7.3f} ms solo of {tot[k] / 1e3:7.3f} ms total ({cnt[k]:3d} launches)  {k}")
