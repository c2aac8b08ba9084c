#!/bin/bash
mkdir -p gpurun_out
( B200NP_HALO_CG2=1 timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_parity_at_size.py -m gpu -x -q -k "conv_block or many_tiles" ) > gpurun_out/pytest_cg2.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_cg2.log
tail -15 gpurun_out/pytest_cg2.log | cut -c1-300
for v in 1 0; do
B200NP_HALO_CG2=$v timeout 300 python bench.py --roofline-only 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('CG2=$v fwd ms', d['launch_ms'], 'wgrad ms', d['second_kernel']['launch_ms'])"
done
B200NP_HALO_CG2=1 timeout 600 python bench.py --no-cpu-baseline --no-dropin 2>/dev/null | grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('CG2=1 ms/step', d['ms_per_step'], 'value', d['value'])"
