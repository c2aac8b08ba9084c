#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -m gpu -q -s -k "favor_cluster_split" 2>&1 | grep -E "split|passed|failed" | head -16
python tools/bench_mmaml.py 5 2>gpurun_out/bench_mmaml.err | tail -1 > gpurun_out/bench_mmaml.json
tail -3 gpurun_out/bench_mmaml.err
python -c "
import json; d=json.load(open('gpurun_out/bench_mmaml.json'))
for k,v in d.items():
    if isinstance(v, dict): print(k, round(v['ms_per_meta_iteration'],2), 'ms', round(v['tasks_per_s'],1), 'tasks/s', 'loss', round(v['loss_first'],5), round(v['loss_last'],5), 'mem', round(v['peak_mem_GB'],2))
"
