#!/bin/bash
# 2-GPU re-check after the last single-GPU changes: sharded-gradient parity, N=1 and N=2 weak-scaling lines on the SAME box
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpu_parity_r2.txt
( python -m pytest tests/test_gpu_multi.py -m gpu -q -s ) >> gpurun_out/multi_gpu_parity_r2.txt 2>&1
tail -4 gpurun_out/multi_gpu_parity_r2.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-dropin 2>/dev/null | grep '^{' | tail -1 > gpurun_out/bench_r2_n1_samebox.json
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 "$@" 2>/dev/null | grep '^{' | tail -1; }
run --steps 20 --warmup 5 > gpurun_out/bench_r2_n2.json
for f in gpurun_out/bench_r2_n1_samebox.json gpurun_out/bench_r2_n2.json; do python -c "
import json,sys; d=json.load(open('$f')); print('$f', d['n_gpus'], d['scaling'], 'ms', round(d['ms_per_step'],3), 'tasks/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"; done
