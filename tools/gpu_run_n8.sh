#!/bin/bash
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 "$@" 2>gpurun_out/n8.err | grep '^{' | tail -1; }
run --steps 20 --warmup 5 > gpurun_out/bench_r2_n8.json
run --steps 20 --warmup 5 --scaling strong --tasks 64 > gpurun_out/bench_r2_n8_strong64.json
run --steps 10 --warmup 3 --scaling strong --tasks 512 > gpurun_out/bench_r2_n8_strong512.json
for f in gpurun_out/bench_r2_n8*.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', d['n_gpus'], d['scaling'], d['config']['global_tasks'], 'ms', round(d['ms_per_step'],3), 'tasks/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"; done
tail -3 gpurun_out/n8.err
