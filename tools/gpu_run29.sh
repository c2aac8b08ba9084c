#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( time python bench.py ) > gpurun_out/bench_n1.log 2>&1
grep '^{' gpurun_out/bench_n1.log | tail -1 > gpurun_out/bench_r2_n1.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n1.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'dropin', d['e2e_dropin']['fixed_shot']['value'], d['e2e_dropin']['shot_uniform_1_15']['value'])
print('roof', d['roofline']['launch_ms'], d['roofline']['frac'], 'second', d['roofline']['second_kernel']['launch_ms'], d['roofline']['second_kernel']['frac'])
print('cpu', d['cpu_baseline']['value'], 'gpu_ref', json.dumps(d['gpu_reference'])[:300])
print('hbm', json.dumps({k: round(v['frac_of_hbm_peak'], 3) for k, v in d['hbm_kernels'].items()}))
PY
( time python bench.py --impl reference --steps 3 --warmup 1 ) 2>&1 | tail -4 | cut -c1-400
TIMELINE_TRACE=gpurun_out/trace_r2_final.txt python tools/timeline_gaps.py > gpurun_out/timeline_r2_final.txt 2>&1
head -5 gpurun_out/timeline_r2_final.txt
python tools/profile_step.py > gpurun_out/profile_step_r2_final.txt 2>&1
PROFILE_MODEL=CNPShapeNet1D python tools/profile_step.py > gpurun_out/profile_step_cnp1d_r2_final.txt 2>&1
python tools/sweep.py full > gpurun_out/sweep_r2_final.jsonl 2> gpurun_out/sweep_r2_final.err; tail -2 gpurun_out/sweep_r2_final.jsonl | cut -c1-200
