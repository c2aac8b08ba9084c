#!/usr/bin/env python
"""Golden vectors of the reference's MMAML conv nets (build container only) -> tests/golden/golden_mmaml_v1.npz.

The UNMODIFIED networks/gated_conv_net.py and networks/conv_embedding_model.py are imported through oracle/ref_shims.py,
built exactly like networks/MMAMLShapeNet1D.py:31-75 does (seed 2578, GatedConvModel then ConvEmbeddingModel), and run
on one task of 15 integer-hash images: embeddings = embedding_model(x); logits = model(x, embeddings=embeddings);
azimuth loss (trainer/losses.py:59-61); first-order gradients of every parameter of both nets.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims, synth  # noqa: E402

N_IMG, SEED = 15, 31


def probe_vec(n, seed=7):
    return (synth.hash_u32((n,), seed) & np.uint32(1)).astype(np.float64) * 2.0 - 1.0


def fingerprint(t):
    a = t.detach().double().reshape(-1).numpy()
    return np.array([a.sum(), np.abs(a).sum(), np.sqrt((a * a).sum()), float(a @ probe_vec(a.size))])


def build():
    ref_shims.install()
    import importlib
    gcn = importlib.import_module("networks.gated_conv_net")
    cem = importlib.import_module("networks.conv_embedding_model")
    torch.manual_seed(2578)
    model = gcn.GatedConvModel(input_channels=1, output_size=2, use_max_pool=False, num_channels=32, img_side_len=128,
                               condition_type='affine', condition_order='low2high', verbose=False)
    emb = cem.ConvEmbeddingModel(input_size=np.prod((1, 128, 128)), output_size=2, embedding_dims=[64, 128, 256, 512],
                                 hidden_size=128, num_layers=2, convolutional=True, num_conv=4, num_channels=32,
                                 rnn_aggregation=False, embedding_pooling='avg', batch_norm=True, avgpool_after_conv=True,
                                 linear_before_rnn=False, num_sample_embedding=0, img_size=(1, 128, 128), verbose=False)
    return model, emb


def main():
    assert ref_shims.reference_available()
    model, emb = build()
    out = {}
    for tag, m in (("model", model), ("emb", emb)):
        sd = m.state_dict()
        out[f"{tag}/keys"] = np.array(list(sd.keys()))
        out[f"{tag}/init_fp"] = np.stack([fingerprint(v.float()) for v in sd.values()])
    cx, cy, tx, ty = synth.task_batch("shapenet_1d", 1, N_IMG, 1, seed=SEED)
    x = torch.from_numpy(cx[0])
    y = torch.from_numpy(cy[0])
    embeddings = emb(x)
    logits = model(x, embeddings=embeddings)
    loss = torch.mean(torch.sum((y[..., :2] - logits) ** 2, dim=-1))
    loss.backward()
    out["logits"] = logits.detach().numpy()
    out["loss"] = np.array(loss.item())
    for j, e in enumerate(embeddings):
        out[f"embedding{j}"] = e.detach().numpy()
    for tag, m in (("model", model), ("emb", emb)):
        names, fps = [], []
        for k, p in m.named_parameters():
            if p.grad is not None:
                names.append(k)
                fps.append(fingerprint(p.grad))
                if p.numel() <= 1024:
                    out[f"{tag}/grad/{k}"] = p.grad.numpy().copy()
        out[f"{tag}/grad_keys"] = np.array(names)
        out[f"{tag}/grad_fp"] = np.stack(fps)
        bufs = dict(m.named_buffers())
        pre = "features.layer1_bn." if tag == "model" else "conv.bn1."
        out[f"{tag}/running_mean1"] = bufs[pre + "running_mean"].numpy().copy()
        out[f"{tag}/running_var1"] = bufs[pre + "running_var"].numpy().copy()
    model.zero_grad()
    out["logits_noemb"] = model(x).detach().numpy()          # plain use of the net, no FiLM
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_mmaml_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "loss", loss.item())


if __name__ == "__main__":
    main()
