#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv or wgrad or stem or favor" 2>&1 | tail -3
run() { label=$1; shift
  r=$(env "$@" python bench.py --no-cpu-baseline --no-dropin --steps 40 --warmup 5 2>/dev/null | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])")
  echo "$label: $r" | tee -a gpurun_out/ab_run26.txt
}
: > gpurun_out/ab_run26.txt
run "default (split4, reducer 32 lanes)" X=1
run "split8" B200NP_FAVOR_SPLIT=8
run "default again" X=1
python tools/profile_step.py 2>/dev/null | grep -E "reduce_partials|favor_attn_fwd|kernel time"
