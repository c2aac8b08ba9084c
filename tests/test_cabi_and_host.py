"""CPU-side checks: the C-ABI library loads and exports every symbol include/b200np.h declares,
the ctypes table covers the header, the product path refuses to run without CUDA, and the
task-sharding / gradient all-reduce logic works with world_size 2 over gloo."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "b200np.h")).read()
    return sorted(set(re.findall(r"\b(b200np_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from b200np import lib
    dll = ctypes.CDLL(lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(dll, s), f"{s} declared in include/b200np.h but not exported"
    assert sorted(lib.SIGNATURES) == syms, set(lib.SIGNATURES) ^ set(syms)
    assert lib.LIB.b200np_version() >= 100
    assert lib.LIB.b200np_strerror(-1).decode().startswith("bad argument")


def test_no_oracle_import_in_product_path():
    pkg = os.path.join(ROOT, "what-matters-for-meta-learning_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("# oracle", ""), f"{f} references the oracle"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda():
    from conftest import build_product_model
    from oracle import synth
    model, _ = build_product_model("cnp_distractor_max")
    cx, cy, tx, _ = (torch.from_numpy(a) for a in synth.task_batch("distractor", 2, 3, 4, seed=1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(cx, cy, tx)


def test_shard_tasks():
    from b200np import dist
    assert dist.shard_tasks(160, 8, 3) == (60, 80)
    assert dist.shard_tasks(20, 1, 0) == (0, 20)
    with pytest.raises(ValueError):
        dist.shard_tasks(20, 8, 0)
    import types
    c = dist.local_config(types.SimpleNamespace(tasks_per_batch=160), world=8)
    assert c.tasks_per_batch == 20


_WORKER = r"""
import os, sys, torch, torch.distributed as td
sys.path[:0] = [sys.argv[1], os.path.join(sys.argv[1], "what-matters-for-meta-learning_b200")]
from b200np import dist
td.init_process_group("gloo")
r, w = td.get_rank(), td.get_world_size()
assert dist.world_size() == 2 and dist.rank() == r
# sharded mean-loss gradient == (1/W) * sum of per-rank gradients (SURVEY.md section 8e)
torch.manual_seed(0)
x = torch.randn(8, 5); wgt = torch.randn(5, requires_grad=True)
full = ((x @ wgt) ** 2).mean(); full.backward(); g_full = wgt.grad.clone()
lo, hi = dist.shard_tasks(8)
wl = wgt.detach().clone().requires_grad_()
((x[lo:hi] @ wl) ** 2).mean().backward()
g = wl.grad.clone(); dist.all_reduce_grads(g); g /= w
assert torch.allclose(g, g_full, atol=1e-6), (g, g_full)
m = torch.tensor([float(r + 1)]); dist.all_reduce_max(m); assert float(m) == 2.0
td.destroy_process_group()
print("ok", r)
"""


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29631", str(script), ROOT]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
