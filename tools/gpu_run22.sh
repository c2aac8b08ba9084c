#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -m gpu -q -k "loss" 2>&1 | tail -5
python tools/hbm_kernels.py 2>&1 | tail -30
python tools/run_dominant_kernel.py tf32x3 stem 420 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stem_ -c 4 -o gpurun_out/stem_r2 -f python tools/run_dominant_kernel.py tf32x3 stem 420 > gpurun_out/stem_r2_ncu.log 2>&1
tail -3 gpurun_out/stem_r2_ncu.log
