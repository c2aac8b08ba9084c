#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_second_order.py tests/test_mmaml.py -m gpu -q 2>&1 | tail -4
python tools/bench_mmaml.py 5 2>gpurun_out/bench_mmaml.err | tail -1 > gpurun_out/bench_mmaml.json
tail -2 gpurun_out/bench_mmaml.err
python -c "
import json; d=json.load(open('gpurun_out/bench_mmaml.json'))
for k,v in d.items():
    if isinstance(v, dict): print(k, round(v['ms_per_meta_iteration'],2), 'ms', round(v['tasks_per_s'],1), 'tasks/s', 'loss', round(v['loss_first'],5), round(v['loss_last'],5), 'mem', round(v['peak_mem_GB'],2))
"
