#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_parity_at_size.py tests/test_gpu_models.py -m gpu -x -q ) > gpurun_out/pytest_trunc.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_trunc.log
grep -E "passed|failed|parity@size|^E  " gpurun_out/pytest_trunc.log | head -20
python bench.py --roofline-only 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('wgrad ms', d['launch_ms'], 'fwd ms', d['second_kernel']['launch_ms'])"
python bench.py --no-cpu-baseline --no-dropin 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', d['ms_per_step'], 'value', d['value'])"
python tools/diag_precision.py 2>&1 | tail -30
