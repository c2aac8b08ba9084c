#!/bin/bash
mkdir -p gpurun_out
python bench.py --no-cpu-baseline --no-dropin 2>/dev/null | grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 ms/step', d['ms_per_step'], 'value', d['value'], 'roof fwd', d['roofline']['launch_ms'], 'wgrad', d['roofline']['second_kernel']['launch_ms'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 5 2>/dev/null | grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2 ms/step', d['ms_per_step'], 'value', d['value'])"
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 20 --warmup 5 --no-graph 2>/dev/null | grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2 eager ms/step', d['ms_per_step'])"
