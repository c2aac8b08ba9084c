#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py ) > gpurun_out/bench_n1.log 2>&1
grep '^{' gpurun_out/bench_n1.log | tail -1 > gpurun_out/bench_r2_n1.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2_n1.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'dropin', d['e2e_dropin']['fixed_shot']['value'], d['e2e_dropin']['shot_uniform_1_15']['value'], 'cpu', d['cpu_baseline']['value'])
PY
python tools/sweep.py full > gpurun_out/sweep_r2_final.jsonl 2> gpurun_out/sweep_r2_final.err; python - <<'PY'
import json
for l in open('gpurun_out/sweep_r2_final.jsonl'):
    d=json.loads(l)
    if d.get('model')!='ANPDistractor' or (d['tasks']==20 and d['nc']==15):
        print(d['model'], d.get('agg'), d.get('precision'), d['tasks'], round(d['ms_per_step'],3), round(d['tasks_per_s'],1))
PY
