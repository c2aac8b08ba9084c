#!/bin/bash
for v in 1 0; do
B200NP_HALO_CG2=$v timeout 300 python bench.py --roofline-only 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('CG2=$v fwd ms', d['launch_ms'], 'wgrad ms', d['second_kernel']['launch_ms'])"
done
B200NP_HALO_CG2=1 timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "conv_block" 2>&1 | tail -2
