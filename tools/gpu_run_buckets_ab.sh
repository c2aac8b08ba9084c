#!/bin/bash
mkdir -p gpurun_out
run() { env "$1" python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline --no-dropin 2>/dev/null | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'ms', round(d['ms_per_step'],3), 'tasks/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))"; }
run B200NP_BUCKETS=0
run B200NP_BUCKETS=1
