#!/bin/bash
# TMA-fed halo kernel: correctness (op tests + at-size parity) and A/B timing against the register-staged kernel
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_parity_at_size.py tests/test_gpu_models.py -m gpu -x -q ) > gpurun_out/pytest_tma.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_tma.log
tail -15 gpurun_out/pytest_tma.log
for v in 1 0; do
  echo "== B200NP_HALO_TMA=$v roofline-only"; B200NP_HALO_TMA=$v timeout 300 python bench.py --roofline-only 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('wgrad ms', d['launch_ms'], 'fwd ms', d['second_kernel']['launch_ms'])"
  echo "== B200NP_HALO_TMA=$v step"; B200NP_HALO_TMA=$v timeout 600 python bench.py --no-cpu-baseline --no-dropin 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', d['ms_per_step'], 'value', d['value'], 'launches', d['gpu_launches'])"
done 2>&1 | tee gpurun_out/tma_ab.log
