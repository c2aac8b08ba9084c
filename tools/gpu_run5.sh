#!/bin/bash
# TMA-fed halo weight gradient: correctness and A/B timing
mkdir -p gpurun_out
( B200NP_WGRAD_HALO=1 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_parity_at_size.py -m gpu -x -q -k "conv_block or many_tiles" ) > gpurun_out/pytest_wg.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_wg.log
tail -12 gpurun_out/pytest_wg.log
for cfg in "0 1" "1 0" "1 1"; do
  set -- $cfg
  echo "== WGRAD_HALO=$1 WGRAD_TMA=$2 roofline-only"; B200NP_WGRAD_HALO=$1 B200NP_WGRAD_TMA=$2 timeout 300 python bench.py --roofline-only 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('wgrad ms', d['launch_ms'], 'fwd ms', d['second_kernel']['launch_ms'])"
done 2>&1 | tee gpurun_out/wg_ab.log
echo "== WGRAD_HALO=1 WGRAD_TMA=1 step"; B200NP_WGRAD_HALO=1 timeout 600 python bench.py --no-cpu-baseline --no-dropin 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', d['ms_per_step'], 'value', d['value'])" | tee -a gpurun_out/wg_ab.log
