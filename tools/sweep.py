#!/usr/bin/env python
"""Throughput sweep of the meta-train step on ONE GPU (BASELINE.json configs 2, 4 and 5):

* ANPDistractor over the context-set size nc (nt = 36 - nc, the train-mode split of the 36 views,
  dataset/shapenet_distractor.py:287-294) and the meta-batch size T,
* the other model families at their shipped shapes (CNPShapeNet1D mean aggregation T=10, ANP / ShapeNet3D T=20, ...),
* the three precision modes on the headline shape.

One step = zero_grad + forward + loss + backward + Adam, replayed as a CUDA graph (b200np.optim.GraphedStep), timed with
CUDA events after warm-up.  Inputs are uniform random images generated on the device (throughput only; parity lives in
tests/).  Prints one JSON object per line:  python tools/sweep.py [quick|full] > gpurun_out/sweep.jsonl
"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "what-matters-for-meta-learning_b200"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402

from b200np import engine  # noqa: E402
from b200np.optim import FlatParams, FusedAdam, GraphedStep  # noqa: E402
from conftest import CASES, make_config  # noqa: E402  (the same config namespaces the parity tests build)
from oracle.synth import TASKS  # noqa: E402  (shape table only)
from trainer.losses import LossFunc  # noqa: E402

DEV = "cuda:0"


def device_batch(task, T, nc, nt, seed):
    C, H, W, label_dim, _ = TASKS[task]
    g = torch.Generator(device=DEV).manual_seed(seed)
    r = lambda *s: torch.rand(*s, device=DEV, generator=g)
    if task == "distractor":
        lab = lambda n: r(T, n, 2) * 128.0
    elif task == "shapenet_1d":
        def lab(n):
            a = r(T, n) * 6.2831853
            return torch.stack([a.cos(), a.sin(), a], -1)
    else:
        def lab(n):
            q = r(T, n, 4) * 2 - 1
            return q / q.norm(dim=-1, keepdim=True)
    return [r(T, nc, C, H, W), lab(nc), r(T, nt, C, H, W), lab(nt)]


def measure(case, T, nc, nt, prec, target_ms=400.0):
    method, task, agg, img_agg, extra, _, _, _ = CASES[case]
    engine.set_precision(prec)
    cfg = make_config(method, task, T, agg, img_agg, device=DEV, **extra)
    model = getattr(importlib.import_module("networks." + method), method)(cfg).to(DEV)
    opt = FusedAdam(FlatParams(model), lr=1e-4)
    lossf = LossFunc("mse", task)
    batches = [device_batch(task, T, nc, nt, s) for s in range(2)]
    torch.cuda.reset_peak_memory_stats()
    gs = GraphedStep(model, lossf, opt, batches[0], warmup=2)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    e0, e1 = ev(), ev()
    gs(batches[1]); gs(batches[0])
    torch.cuda.synchronize()
    e0.record(); gs(batches[1]); e1.record()
    torch.cuda.synchronize()
    steps = int(max(3, min(30, target_ms / max(e0.elapsed_time(e1), 1e-3))))
    e0.record()
    for i in range(steps):
        gs(batches[i & 1])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    loss = float(gs.loss)
    out = dict(model=method, task=task, agg=agg, precision=prec, tasks=T, nc=nc, nt=nt, steps=steps,
               ms_per_step=round(ms, 4), tasks_per_s=round(T / ms * 1e3, 1),
               peak_mem_GB=round(torch.cuda.max_memory_allocated() / 1e9, 2), loss_finite=bool(loss == loss))
    del gs, model, opt, batches
    torch.cuda.empty_cache()
    return out


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "quick"
    jobs = []
    # (4)/(2): the families at their shipped shapes (cfg/train/*.yaml: tasks_per_batch, max_ctx_num / views)
    jobs += [("cnp_1d_mean", 10, 15, 15, "tf32x3"), ("cnp_1d_max", 10, 15, 15, "tf32x3"), ("anp_1d", 10, 15, 15, "tf32x3"),
             ("anp_3d", 20, 15, 15, "tf32x3"), ("cnp_3d_max", 20, 15, 15, "tf32x3"),
             ("cnp_distractor_max", 20, 15, 21, "tf32x3"), ("cnp_distractor_baco", 20, 15, 21, "tf32x3")]
    # precision modes on the headline shape
    jobs += [("anp_distractor", 20, 15, 21, p) for p in ("tf32x3", "tf32", "fp32")]
    # (5): context-set size x meta-batch
    Ts = (8, 20, 64, 128, 256, 512) if mode == "full" else (8, 64)
    ncs = (1, 5, 10, 15, 20, 25) if mode == "full" else (1, 25)
    for T in Ts:
        for nc in ncs:
            if (T, nc) != (20, 15):
                jobs.append(("anp_distractor", T, nc, 36 - nc, "tf32x3"))
    for case, T, nc, nt, prec in jobs:
        try:
            print(json.dumps(measure(case, T, nc, nt, prec)), flush=True)
        except Exception as e:  # keep sweeping; the line records the failure
            print(json.dumps(dict(case=case, tasks=T, nc=nc, nt=nt, precision=prec, error=repr(e)[:300])), flush=True)
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
