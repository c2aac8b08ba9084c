#!/usr/bin/env python
"""Golden vectors of the reference's Bayes-by-backprop layers (build container only) -> tests/golden/golden_bbb_v1.npz.

The UNMODIFIED networks/bbb/BBBConv.py / BBBLinear.py (through oracle/ref_shims.py): an MR-encoder-shaped stack
(networks/CNPMR.py:29-52: BBBConv2d 1->32 and 32->48, 3x3 stride 2 padding 1, ReLU; flatten; BBBLinear) built under
torch.manual_seed(2578), run in training mode under torch.manual_seed(77) -- the reference draws its noise with the
CPU generator (BBBConv.py:86), so the draw is reproducible -- on 4 integer-hash images; loss = sum(out^2) + 1e-3 * kl.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shims, synth  # noqa: E402


def build(bbb):
    torch.manual_seed(2578)
    c1 = bbb.BBBConv2d(1, 32, 3, stride=2, padding=1, bias=True)
    c2 = bbb.BBBConv2d(32, 48, 3, stride=2, padding=1, bias=True)
    fc = bbb.BBBLinear(48 * 8 * 8, 16, bias=True)
    return c1, c2, fc


def run(c1, c2, fc, x):
    torch.manual_seed(77)
    h = torch.relu(c1(x))
    h = torch.relu(c2(h))
    out = fc(h.reshape(h.size(0), -1))
    kl = c1.kl_loss() + c2.kl_loss() + fc.kl_loss()
    loss = out.pow(2).sum() + 1e-3 * kl
    return out, kl, loss


def main():
    assert ref_shims.reference_available()
    ref_shims.install()
    import importlib
    bbb = importlib.import_module("networks.bbb")
    c1, c2, fc = build(bbb)
    for m in (c1, c2, fc):
        m.train()
    x = torch.from_numpy(synth.images((4, 1, 32, 32), 5))
    out, kl, loss = run(c1, c2, fc, x)
    loss.backward()
    res = {"out": out.detach().numpy(), "kl": np.array(kl.item()), "loss": np.array(loss.item())}
    for tag, m in (("c1", c1), ("c2", c2), ("fc", fc)):
        for k, p in m.named_parameters():
            res[f"{tag}/init/{k}"] = p.detach().numpy().copy()
            res[f"{tag}/grad/{k}"] = p.grad.numpy().copy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_bbb_v1.npz")
    np.savez_compressed(path, **res)
    print("wrote", path, os.path.getsize(path), "kl", kl.item(), "loss", loss.item())


if __name__ == "__main__":
    main()
