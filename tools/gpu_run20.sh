#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=8 -rs ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|^ERROR|mmaml/|parity@size" gpurun_out/pytest_gpu.log | head -30
( time python bench.py ) > gpurun_out/bench_n1.log 2>&1
grep '^{' gpurun_out/bench_n1.log | tail -1 > gpurun_out/bench_r2_n1.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n1.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'dropin', d['e2e_dropin']['fixed_shot']['value'], d['e2e_dropin']['shot_uniform_1_15']['value'])
print('roof', d['roofline']['launch_ms'], d['roofline']['frac'], 'second', d['roofline']['second_kernel']['launch_ms'], d['roofline']['second_kernel']['frac'])
print('cpu', d['cpu_baseline'], 'gpu_ref', d['gpu_reference'])
PY
( time python bench.py --impl reference --steps 3 --warmup 1 ) 2>&1 | tail -4 | cut -c1-600
python tools/diag_precision.py > gpurun_out/precision_r2.txt 2>&1
python tools/sweep.py full > gpurun_out/sweep_r2.jsonl 2> gpurun_out/sweep_r2.err; tail -3 gpurun_out/sweep_r2.jsonl | cut -c1-200
