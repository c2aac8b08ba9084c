#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py ) > gpurun_out/bench_n1.log 2>&1
grep '^{' gpurun_out/bench_n1.log | tail -1 > gpurun_out/bench_r2_n1.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n1.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'dropin', d['e2e_dropin']['fixed_shot']['value'], d['e2e_dropin']['shot_uniform_1_15']['value'])
print('roof', d['roofline']['launch_ms'], d['roofline']['frac'], 'second', d['roofline']['second_kernel']['launch_ms'], d['roofline']['second_kernel']['frac'], 'launches', d.get('gpu_launches'), 'clocks', d.get('clocks'))
PY
python tools/profile_step.py 2>/dev/null | grep -v "Warn\|_warn_once" > gpurun_out/profile_step_r2_final.txt; head -3 gpurun_out/profile_step_r2_final.txt
TIMELINE_TRACE=gpurun_out/trace_r2_final.txt python tools/timeline_gaps.py 2>/dev/null > gpurun_out/timeline_r2_final.txt; head -4 gpurun_out/timeline_r2_final.txt
