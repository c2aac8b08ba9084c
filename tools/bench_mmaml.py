#!/usr/bin/env python
"""One MMAML meta-iteration (cfg/train/MMAML_ShapeNet1D_DA+TA.yaml: 10 tasks, 15 context / 15 query images of 128x128,
5 inner updates with create_graph=True, outer backward, two Adam steps) driven by the reference's own MetaLearner:
the B200 drop-in nets against the reference's nets through torch / cuDNN on the same GPU.

    python tools/bench_mmaml.py [steps] > gpurun_out/bench_mmaml.json
"""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "what-matters-for-meta-learning_b200")
sys.path[:0] = [ROOT, PKG]
os.environ["B200NP_MMAML"] = "1"
import numpy as np
import torch

from oracle import ref_shims, synth

ref_shims.install()
import b200_run  # noqa: E402
b200_run.setup_path(ref_shims.REFERENCE_ROOT)
from b200np import engine  # noqa: E402
from networks import _refload  # noqa: E402
from networks.conv_embedding_model import ConvEmbeddingModel  # noqa: E402
from networks.gated_conv_net import GatedConvModel  # noqa: E402
from trainer.losses import LossFunc  # noqa: E402

T, NC, NQ, UPDATES, STEPS = 10, 15, 15, 5, int(sys.argv[1]) if len(sys.argv) > 1 else 5
spec = importlib.util.spec_from_file_location("_ref_meta_learner_reg",
                                              os.path.join(ref_shims.REFERENCE_ROOT, "trainer", "meta_learner_reg.py"))
mlr = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mlr)


def build(gcls, ecls):
    torch.manual_seed(2578)
    model = gcls(input_channels=1, output_size=2, use_max_pool=False, num_channels=32, img_side_len=128,
                 condition_type='affine', condition_order='low2high', verbose=False)
    emb = ecls(input_size=np.prod((1, 128, 128)), output_size=2, embedding_dims=[64, 128, 256, 512], hidden_size=128,
               num_layers=2, convolutional=True, num_conv=4, num_channels=32, rnn_aggregation=False,
               embedding_pooling='avg', batch_norm=True, avgpool_after_conv=True, linear_before_rnn=False,
               num_sample_embedding=0, img_size=(1, 128, 128), verbose=False)
    return model.cuda(), emb.cuda()


class TorchLoss:
    def calc_loss(self, pred, var, y, test=False):
        return torch.mean(torch.sum((y[..., :2] - pred) ** 2, dim=-1))


def run(model, emb, lossf, tag):
    opts = [torch.optim.Adam(model.parameters(), lr=5e-4), torch.optim.Adam(emb.parameters(), lr=5e-4)]
    ml = mlr.MetaLearner(model, emb, opts, fast_lr=0.002, loss_func=lossf, first_order=False, num_updates=UPDATES,
                         inner_loop_grad_clip=20.0, collect_accuracies=False, device="cuda", embedding_grad_clip=2.0,
                         model_grad_clip=2.0)
    cx, cy, qx, qy = (torch.from_numpy(a).cuda() for a in synth.task_batch("shapenet_1d", T, NC, NQ, seed=1))
    losses = []

    def it():
        _, adapted, embeddings = ml.adapt(cx, cy)
        losses.append(float(ml.step(adapted, embeddings, qx, qy, is_training=True, test=False)["loss"]))
    for _ in range(2):
        it()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        it()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    return {"impl": tag, "ms_per_meta_iteration": ms, "tasks_per_s": T / ms * 1e3, "loss_first": losses[0], "loss_last": losses[-1],
            "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}


out = {"workload": f"MMAML meta-iteration: {T} tasks x ({UPDATES} inner updates, create_graph=True) + outer step, "
                   f"{NC} context / {NQ} query images 128x128x1", "steps": STEPS}
ref_g = _refload.reference_module("gated_conv_net").GatedConvModel
ref_e = _refload.reference_module("conv_embedding_model").ConvEmbeddingModel
for prec in ("tf32x3", "tf32"):
    engine.set_precision(prec)
    torch.cuda.reset_peak_memory_stats()
    out[f"b200_{prec}"] = run(*build(GatedConvModel, ConvEmbeddingModel), LossFunc("mse", "shapenet_1d"), f"b200 {prec}")
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.cuda.reset_peak_memory_stats()
    out["reference_cudnn_" + ("tf32_default" if tf32 else "fp32")] = run(*build(ref_g, ref_e), TorchLoss(), "reference torch/cuDNN")
print(json.dumps(out))

if os.environ.get("PROFILE"):
    from torch.profiler import ProfilerActivity, profile
    import collections
    for tag, (g_, e_, lf) in (("b200", (GatedConvModel, ConvEmbeddingModel, LossFunc("mse", "shapenet_1d"))),
                              ("reference", (ref_g, ref_e, TorchLoss()))):
        engine.set_precision("tf32x3")
        model, emb = build(g_, e_)
        opts = [torch.optim.Adam(model.parameters(), lr=5e-4), torch.optim.Adam(emb.parameters(), lr=5e-4)]
        ml = mlr.MetaLearner(model, emb, opts, fast_lr=0.002, loss_func=lf, first_order=False, num_updates=UPDATES,
                             inner_loop_grad_clip=20.0, collect_accuracies=False, device="cuda", embedding_grad_clip=2.0,
                             model_grad_clip=2.0)
        cx, cy, qx, qy = (torch.from_numpy(a).cuda() for a in synth.task_batch("shapenet_1d", T, NC, NQ, seed=1))
        for _ in range(2):
            _, ad, em = ml.adapt(cx, cy); ml.step(ad, em, qx, qy, is_training=True, test=False)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            _, ad, em = ml.adapt(cx, cy); ml.step(ad, em, qx, qy, is_training=True, test=False)
            torch.cuda.synchronize()
        tot, cnt = collections.defaultdict(float), collections.Counter()
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                tot[e.name[:70]] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
                cnt[e.name[:70]] += 1
        print(f"# {tag}: {sum(cnt.values())} launches, {sum(tot.values()) / 1e3:.1f} ms of kernel time per meta-iteration", file=sys.stderr)
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:14]:
            print(f"#   {v / 1e3:8.2f} ms {cnt[k]:6d}x  {k}", file=sys.stderr)
