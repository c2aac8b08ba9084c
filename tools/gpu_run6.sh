#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=8 -rs ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|^ERROR|mmaml/|parity@size" gpurun_out/pytest_gpu.log | head -30
( time python bench.py --no-cpu-baseline ) > gpurun_out/bench_n1.log 2>&1
tail -1 gpurun_out/bench_n1.log | head -c 300
python - <<'PY'
import json
for l in open('gpurun_out/bench_n1.log'):
    if l.startswith('{'):
        d=json.loads(l); print('\nms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'dropin', d['e2e_dropin']['fixed_shot'], 'roof', d['roofline']['launch_ms'], d['roofline']['frac'], 'fwd', d['roofline']['second_kernel']['launch_ms'], 'launches', d['gpu_launches'])
PY
