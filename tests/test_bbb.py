"""Bayes-by-backprop layers of the "MR" variants (networks/bbb/BBBConv.py:37-105, BBBLinear.py).

The reference draws the weight noise with the CPU generator, so parity is exact under the same seed: the drop-in layers
(B200NP_BBB=1) must reproduce the reference's seeded initialisation bit for bit (CPU test) and, on the GPU, its outputs,
its KL term and every parameter gradient for an MR-encoder-shaped stack (golden vectors:
tests/golden/make_golden_bbb.py)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_l2
from oracle import synth


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_bbb_v1.npz"), allow_pickle=False)


def _layers():
    os.environ["B200NP_BBB"] = "1"
    for k in [k for k in sys.modules if k.startswith("networks.bbb")]:
        del sys.modules[k]
    import importlib
    bbb = importlib.import_module("networks.bbb")
    assert bbb.BBBConv2d.__module__ == "networks.bbb.BBBConv"
    torch.manual_seed(2578)
    c1 = bbb.BBBConv2d(1, 32, 3, stride=2, padding=1, bias=True)
    c2 = bbb.BBBConv2d(32, 48, 3, stride=2, padding=1, bias=True)
    fc = bbb.BBBLinear(48 * 8 * 8, 16, bias=True)
    return c1, c2, fc


def test_seeded_init_matches_reference(golden):
    for tag, m in zip(("c1", "c2", "fc"), _layers()):
        for k, p in m.named_parameters():
            np.testing.assert_array_equal(p.detach().numpy(), golden[f"{tag}/init/{k}"])


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["tf32x3", "fp32"])
def test_bbb_stack_matches_reference(prec, golden):
    from b200np import engine
    engine.set_precision(prec)
    c1, c2, fc = _layers()
    for m in (c1, c2, fc):
        m.to("cuda")
        m.device = "cuda"
        m.train()
    x = torch.from_numpy(synth.images((4, 1, 32, 32), 5)).cuda()
    torch.manual_seed(77)                     # the noise comes from the CPU generator, like in the reference
    h = torch.relu(c1(x))
    h = torch.relu(c2(h))
    out = fc(h.reshape(h.size(0), -1))
    kl = c1.kl_loss() + c2.kl_loss() + fc.kl_loss()
    loss = out.pow(2).sum() + 1e-3 * kl
    loss.backward()
    torch.cuda.synchronize()
    assert rel_l2(out.detach().cpu().numpy(), golden["out"]) < 1e-4
    assert abs(float(kl) - float(golden["kl"])) < 1e-5 * abs(float(golden["kl"]))
    assert abs(float(loss) - float(golden["loss"])) < 1e-4 * abs(float(golden["loss"]))
    for tag, m in (("c1", c1), ("c2", c2), ("fc", fc)):
        for k, p in m.named_parameters():
            assert rel_l2(p.grad.cpu().numpy(), golden[f"{tag}/grad/{k}"]) < 1e-3, (tag, k)
    # evaluation mode without sampling uses the posterior means
    c1.eval()
    y = c1(x, sample=False)
    ref = torch.nn.functional.conv2d(x.double(), c1.W_mu.double(), c1.bias_mu.double(), stride=2, padding=1)
    assert rel_l2(y.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 1e-5
