#!/usr/bin/env python
"""Benchmark: ANPDistractor meta-train tasks/sec (BASELINE.json's metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--precision tf32x3|tf32|fp32]

One step = zero_grad + forward + loss + backward + (gradient all-reduce) + Adam on a synthetic
Distractor-shaped meta-batch: 20 tasks per GPU (cfg/train/ANP_Distractor.yaml:10), 15 context
and 21 target images of 128x128x1 per task (weak scaling: per-GPU work is fixed).

Prints ONE JSON line (rank 0).  `value` is timed with inputs resident in HBM; `e2e` times the same
step through the public API with pinned HOST inputs copied in and the loss read back every step.
`--impl reference` times the CPU restatement of the reference (oracle/np_oracle.py, kind "port":
the Python reference itself cannot travel to the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "what-matters-for-meta-learning_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

TASKS_PER_GPU, NC, NT = 20, 15, 21
METRIC, UNIT = "ANPDistractor meta-train tasks/sec", "tasks/s"


def make_cfg(T, device):
    return types.SimpleNamespace(device=device, img_size=[128, 128, 1], task="distractor", tasks_per_batch=T,
                                 input_dim=2, output_dim=2, agg_mode="attention", img_agg="max", dim_w=16,
                                 seed=2578)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"], src="measured")
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, src="fallback")


# --------------------------------------------------------------------------------------------
# CPU reference arm (oracle port)
# --------------------------------------------------------------------------------------------
def cpu_port_step_time(sample_tasks, steps, warmup):
    from oracle import np_oracle, synth
    from networks.ANPDistractor import ANPDistractor
    torch.set_num_threads(os.cpu_count() or 1)
    model = ANPDistractor(make_cfg(sample_tasks, "cpu"))  # parameter holder only (seeded init)
    ocfg = dict(tasks_per_batch=sample_tasks, agg_mode="attention", img_agg="max", task="distractor")
    tr = np_oracle.OracleTrainer("ANPDistractor", ocfg, model.state_dict(), lr=1e-4)
    batch = [torch.from_numpy(a) for a in synth.task_batch("distractor", sample_tasks, NC, NT, seed=1)]
    for _ in range(warmup):
        tr.step(*batch)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        tr.step(*batch)
        times.append(time.perf_counter() - t0)
    return times


def run_reference_arm(args, rank):
    if rank != 0:
        return
    sample = 2
    times = cpu_port_step_time(sample, args.steps, args.warmup)
    total = sum(times)
    value = sample * len(times) / total
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ANPDistractor meta-train step, nc={NC} nt={NT} 128x128x1, {TASKS_PER_GPU} tasks/GPU",
                   "note": "CPU arm: each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} tasks/step (nc={NC}, nt={NT}), full fwd+loss+bwd+Adam, oracle/np_oracle.py"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nme, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def run_b200_arm(args, rank, world, local_rank):
    import torch.distributed as td
    from b200np import engine, ops
    from b200np.lib import LIB
    from b200np.optim import FlatParams, FusedAdam, GraphedStep
    from networks.ANPDistractor import ANPDistractor
    from oracle import synth  # input generator only (integer hash); no oracle compute here
    from trainer.losses import LossFunc

    if "B200NP_WGRAD_WAVES" in os.environ:   # diagnostic entry point: pixel chunks per weight-gradient launch
        LIB.b200np_debug_set_wgrad_waves(int(os.environ["B200NP_WGRAD_WAVES"]))
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    engine.set_precision(args.precision)
    T = TASKS_PER_GPU
    model = ANPDistractor(make_cfg(T, str(dev))).to(dev)
    flat = FlatParams(model)
    opt = FusedAdam(flat, lr=1e-4)
    lossf = LossFunc("mse", "distractor")

    # 4 distinct resident batches per rank (189 MB of inputs > 126 MB L2; a step also streams ~2 GB of
    # activations, far beyond L2), rotated between iterations
    nbatch = 4
    host = [[torch.from_numpy(a).pin_memory() for a in synth.task_batch("distractor", T, NC, NT, seed=100 * rank + i)]
            for i in range(nbatch)]
    resident = [[t.to(dev) for t in b] for b in host]

    def step(batch):
        cx, cy, tx, ty = batch
        opt.zero_grad()
        mu, _, _ = model(cx, cy, tx)
        loss = lossf.calc_loss(mu, None, ty)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            td.all_reduce(ms, op=td.ReduceOp.MAX)
        return float(ms)

    launches_per_step = None
    if args.graph:
        # the whole step (every kernel of fwd + loss + bwd + all-reduce + Adam) captured once, replayed per step
        l0 = LIB.b200np_launch_count()
        gstep = GraphedStep(model, lossf, opt, resident[0], warmup=max(args.warmup, 3))
        launches_per_step = (LIB.b200np_launch_count() - l0) // (max(args.warmup, 3) + 1)
        run = gstep
    else:
        run = step
    for i in range(args.warmup):
        run(resident[i % nbatch])
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = LIB.b200np_launch_count()
    ms = timed(lambda i: run(resident[i % nbatch]), args.steps)
    launches = launches_per_step * args.steps if args.graph else LIB.b200np_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None

    # end to end: pinned host inputs -> device every step, loss read back every step
    def e2e_step(i):
        if args.graph:
            # this step's inputs were put on the wire (pinned host -> device, copy stream) while the previous step
            # computed; this call starts the copy of the next step's inputs and reads this step's loss back
            return float(gstep(host[i % nbatch], next_batch=host[(i + 1) % nbatch]))
        b = [t.to(dev, non_blocking=True) for t in host[i % nbatch]]
        return float(step(b).detach())
    e2e_steps = max(3, min(args.steps, 10))
    h2d = sum(t.numel() * 4 for t in host[0])
    if args.profile:
        ms_e2e = float("nan")
    else:
        e2e_step(0)
        ms_e2e = timed(e2e_step, e2e_steps)

    def finish():
        """Common exit: a captured graph holds NCCL kernels, so drop it before the communicator; every rank
        leaves through the same barrier; multi-rank processes then exit without interpreter teardown (an
        orderly NCCL shutdown after graph capture was observed to stall for minutes)."""
        nonlocal run
        if args.graph:
            run = None
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        finish()
        return
    roof = dominant_kernel_roofline(args, dev) if not args.profile else None
    cpu = None
    if world == 1 and not args.no_cpu_baseline and not args.profile:
        sample = 2
        times = cpu_port_step_time(sample, 3, 1)
        cpu = {"value": sample * len(times) / sum(times), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"3 steps of {sample} tasks (nc={NC}, nt={NT}), fwd+loss+bwd+Adam, oracle/np_oracle.py"}
    tasks = T * world
    line = {
        "metric": METRIC, "value": tasks * args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": {"tf32x3": "f32 (3xTF32 split on tcgen05, fp32 accumulate)", "tf32": "tf32", "fp32": "f32"}[args.precision],
        "data": "synthetic",
        "config": {"workload": f"ANPDistractor meta-train step (fwd+loss+bwd+allreduce+Adam), {T} tasks/GPU, "
                               f"nc={NC} nt={NT}, 128x128x1 images, global tasks={tasks}",
                   "precision": args.precision, "cuda_graph": bool(args.graph),
                   "l2": "inputs rotate over 4 resident batches (189 MB > 126 MB L2); ~2 GB of activations streamed per step"},
        "e2e": {"value": tasks * e2e_steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "steps": e2e_steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    finish()


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel at exactly the size
# timed below, from `ncu --set full` of `bench.py --roofline-only` (profiles/ncu_tapwgrad_r1.txt,
# profiles/ncu_tapconv_halo_r1.txt).
NCU_TRAFFIC_BYTES = {"tapwgrad": 921.6e6, "tapconv_halo": 862.4e6}


def _time_kernel(run, reps=10):
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / reps


def dominant_kernel_roofline(args, dev):
    """Times the dominant kernel of the step alone, with CUDA events on the launching stream (torch's current
    stream, which is the one the C ABI is given).  By CUPTI kernel time (profiles/profile_step_r1.txt) the
    dominant kernel is the tcgen05 weight gradient `tapwgrad_umma_kernel` (28% of the step); it is timed on its
    largest instance, the fused `conv2 3x3 + 1x1 s2 skip` weight gradient of layer1 at the step's own size
    (all 1140 images of a 20-task ANP step: contraction over M = 1140*32*32 pixels, 64 x 640 outputs).
    Algorithmic FLOPs = 2*M*64*640 per launch (75.5 + 8.4 MFLOP/image).
    Peak = measured dense bf16 / 2 (tf32 issues at half the bf16 rate); the fp32-grade 3xTF32 split spends
    3 MMAs per algorithmic MAC, so its ceiling on this scale is 1/3.  The forward kernel of the same layer
    (`tapconv_halo_kernel`, second by time) is reported beside it."""
    from b200np import ops
    from b200np.lib import PREC_FP32_SIMT, PREC_TF32, PREC_TF32X3
    prec = {"tf32x3": PREC_TF32X3, "tf32": PREC_TF32, "fp32": PREC_FP32_SIMT}[args.precision]
    N = TASKS_PER_GPU * (NC + NT) + TASKS_PER_GPU * NT
    g = torch.Generator(device="cpu").manual_seed(0)
    x = torch.rand(N, 64, 64, 64, generator=g).to(dev)          # block input  (NHWC)
    h = torch.rand(N, 32, 32, 64, generator=g).to(dev)          # conv1 output (NHWC)
    dy = torch.randn(N, 32, 32, 64, generator=g).to(dev)        # gradient of the block output
    w2 = (torch.randn(64, 64, 3, 3, generator=g) * 0.04).to(dev)
    ws = (torch.randn(64, 64, 1, 1, generator=g) * 0.1).to(dev)
    b = torch.zeros(64, device=dev)
    wf2 = ops.pack_conv_weight(w2)
    wfs = ops.pack_conv_weight(ws)
    flops = 2.0 * N * 32 * 32 * 64 * 640
    # both kernels touch h, every second pixel of every second row of x (the stride-2 skip) and one more
    # 32x32x64 tensor (dY or the output): 3 x 299 MB
    alg_bytes = (h.numel() + x.numel() // 4 + N * 32 * 32 * 64) * 4
    pk = peaks()
    peak = pk["bf16"] / 2.0

    def entry(name, key, sec):
        ach = flops / sec / 1e12
        return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "peak_source": f"{pk['src']} bf16 burst {pk['bf16']} TFLOP/s / 2 (tf32 rate)",
                "traffic": NCU_TRAFFIC_BYTES[key], "launch_ms": sec * 1e3,
                "hbm_view": {"algorithmic_GB": alg_bytes / 1e9, "achieved_GBps": alg_bytes / sec / 1e9,
                             "peak_GBps": pk["hbm"], "frac": alg_bytes / sec / 1e9 / pk["hbm"]}}

    if prec == PREC_FP32_SIMT:
        t_w = _time_kernel(lambda: ops.conv_wgrad(h, dy, 3, 1, prec))
    else:
        t_w = _time_kernel(lambda: ops.conv_wgrad(h, dy, 3, 1, prec, skip=(x, 2)))
    t_f = _time_kernel(lambda: ops.conv_fwd(h, wf2, b, 1, 1, prec, skip=(x, wfs, b, 2)))
    out = entry("tapwgrad_umma (layer1 conv2 3x3 + 1x1 skip weight gradient, pixel contraction on tcgen05)",
                "tapwgrad", t_w)
    out["second_kernel"] = entry("tapconv_halo (layer1 conv2 3x3 + 1x1 skip + bias + ReLU forward, implicit GEMM "
                                 "on tcgen05)", "tapconv_halo", t_f)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("B200NP_PRECISION", "tf32x3"),
                    choices=["tf32x3", "tf32", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch every kernel from Python each step instead of replaying the captured CUDA graph")
    ap.add_argument("--roofline-only", action="store_true",
                    help="run only the dominant-kernel timing (the command captured by `ncu --set full`)")
    ap.add_argument("--profile", action="store_true",
                    help="only the resident-input timed loop (for ncu launch lists); skips e2e / roofline / CPU legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if args.roofline_only:
        print(json.dumps(dominant_kernel_roofline(args, torch.device("cuda:0"))))
        return
    if args.profile:
        args.graph = False  # ncu needs the individual launches
    else:
        args.warmup = max(args.warmup, 3)
    run_b200_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
