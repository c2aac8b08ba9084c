#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_parity_at_size.py tests/test_mmaml.py -m gpu -x -q ) > gpurun_out/pytest_s2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_s2.log
tail -6 gpurun_out/pytest_s2.log | cut -c1-300
python - <<'PY' 2>&1 | tee gpurun_out/wg_s2.log
import os, sys, torch
sys.path[:0] = ['.', 'what-matters-for-meta-learning_b200']
from b200np import ops
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
g = torch.Generator().manual_seed(0)
for N in (720, 420):
    x = torch.rand(N, 64, 64, 64, generator=g).cuda(); dh = torch.randn(N, 32, 32, 64, generator=g).cuda()
    print("s2 wgrad layer1 conv1, N =", N, "ms", t(lambda: ops.conv_wgrad(x, dh, 3, 2, 1)))
PY
B200NP_WGRAD_TMA=0 python - <<'PY' 2>&1 | tee -a gpurun_out/wg_s2.log
import os, sys, torch
sys.path[:0] = ['.', 'what-matters-for-meta-learning_b200']
from b200np import ops
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
g = torch.Generator().manual_seed(0)
for N in (720, 420):
    x = torch.rand(N, 64, 64, 64, generator=g).cuda(); dh = torch.randn(N, 32, 32, 64, generator=g).cuda()
    print("(gather kernel) s2 wgrad layer1 conv1, N =", N, "ms", t(lambda: ops.conv_wgrad(x, dh, 3, 2, 1)))
PY
python bench.py --no-cpu-baseline --no-dropin 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', d['ms_per_step'], 'value', d['value'])" | tee -a gpurun_out/wg_s2.log
