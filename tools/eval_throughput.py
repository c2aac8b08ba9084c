#!/usr/bin/env python
"""Evaluation-path throughput: no-grad ANPDistractor forward + loss at the evaluator's shapes
(nc context views, all 36 views as targets, evaluator/model_evaluator.py:102-109), eager vs per-shape CUDA graph."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
import torch
from bench import TASKS_PER_GPU, make_cfg
from b200np import engine
from b200np.optim import GraphedEval
from networks.ANPDistractor import ANPDistractor
from oracle import synth
from trainer.losses import LossFunc

engine.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32x3")
T = TASKS_PER_GPU
model = ANPDistractor(make_cfg(T, "cuda:0")).to("cuda:0").eval()
lossf = LossFunc("mse", "distractor")
ge = GraphedEval(model, lossf)
for nc in (1, 5, 15, 25):
    b = [torch.from_numpy(a).cuda() for a in synth.task_batch("distractor", T, nc, 36, seed=nc)]
    def eager():
        with torch.no_grad():
            mu, _, _ = model(b[0], b[1], b[2], test=True)
            return lossf.calc_loss(mu, None, b[3], test=True)
    res = {}
    for name, fn in (("eager", eager), ("graph", lambda: ge(*b)[1])):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(10):
            fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = (e0.elapsed_time(e1) / 10, (time.perf_counter() - t0) * 100)
    print(f"nc={nc:2d} nt=36 T={T}: eager {res['eager'][0]:.2f} ms GPU / {res['eager'][1]:.2f} ms wall, "
          f"graph {res['graph'][0]:.2f} ms -> {T / res['graph'][0] * 1e3:.0f} tasks/s", flush=True)
