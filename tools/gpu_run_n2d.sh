#!/bin/bash
python -m pytest tests/test_gpu_multi.py -m gpu -q -s 2>&1 | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --no-dropin 2>/dev/null | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=2 ms', round(d['ms_per_step'],3), 'tasks/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-dropin 2>/dev/null | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=1 ms', round(d['ms_per_step'],3), 'tasks/s', round(d['value'],1))"
