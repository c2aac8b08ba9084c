#!/usr/bin/env python
"""How long does the host take to enqueue one step (Python + ctypes + autograd), versus the GPU time?"""
import os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
import torch
from bench import NC, NT, TASKS_PER_GPU, make_cfg
from b200np import engine
from b200np.lib import LIB
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
for _ in range(3): step()
torch.cuda.synchronize()
n = 10
l0 = LIB.b200np_launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(n): step()
t_enq = time.perf_counter() - t0
e1.record(); torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host enqueue {1e3*t_enq/n:.2f} ms/step, GPU (events) {e0.elapsed_time(e1)/n:.2f} ms/step, wall {1e3*t_all/n:.2f} ms/step, "
      f"launches/step {(LIB.b200np_launch_count()-l0)/n:.0f}")
