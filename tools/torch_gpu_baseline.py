#!/usr/bin/env python
"""Context number (not a bench arm): the reference's own PyTorch ops (cuDNN / cuBLAS / ATen, via the
oracle restatement, which calls the same torch functions as the reference modules) on the B200, for
the bench workload (ANPDistractor, T=20, nc=15, nt=21).  Reported in DESIGN.md as the on-box bar."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200")]
from bench import NC, NT, TASKS_PER_GPU, make_cfg  # noqa: E402
from networks.ANPDistractor import ANPDistractor  # noqa: E402
from oracle import np_oracle, synth  # noqa: E402


def run(tf32, steps=20, warmup=5):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    T = TASKS_PER_GPU
    model = ANPDistractor(make_cfg(T, "cpu"))
    ocfg = dict(tasks_per_batch=T, agg_mode="attention", img_agg="max", task="distractor")
    tr = np_oracle.OracleTrainer("ANPDistractor", ocfg, model.state_dict(), lr=1e-4)
    tr.sd = {k: v.detach().cuda().requires_grad_(v.requires_grad) for k, v in tr.sd.items()}
    tr.params = np_oracle.params_with_grad(tr.sd)
    tr.opt = torch.optim.Adam(list(tr.params.values()), lr=1e-4)
    batch = [torch.from_numpy(a).cuda() for a in synth.task_batch("distractor", T, NC, NT, seed=1)]

    def step():
        tr.opt.zero_grad(set_to_none=True)
        mu, loss = tr.forward_loss(*batch)
        loss.backward()
        tr.opt.step()
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"impl": "torch-ops-on-gpu", "tf32": tf32, "ms_per_step": ms, "tasks_per_s": T / ms * 1e3,
            "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}


if __name__ == "__main__":
    for tf32 in (True, False):
        print(json.dumps(run(tf32)), flush=True)
