#!/bin/bash
python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py tests/test_parity_at_size.py -m gpu -q -k "favor or anp or ANP or benchmark_size" 2>&1 | tail -3
for i in 1 2; do python bench.py --no-cpu-baseline --no-dropin --steps 40 --warmup 5 2>/dev/null | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms', d['ms_per_step'], d['value'])"; done
python tools/profile_step.py 2>/dev/null | grep -E "favor_attn|kernel time"
