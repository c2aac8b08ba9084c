#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
grep -E "passed|failed|^FAILED" gpurun_out/pytest_gpu.log | tail -3
python tools/sweep.py full > gpurun_out/sweep_r2_final.jsonl 2> gpurun_out/sweep_r2_final.err; python - <<'PY'
import json
for l in open('gpurun_out/sweep_r2_final.jsonl'):
    d=json.loads(l)
    if d.get('model')!='ANPDistractor' or (d['tasks']==20 and d['nc']==15):
        print(d['model'], d.get('agg'), d.get('precision'), d['tasks'], round(d['ms_per_step'],3), round(d['tasks_per_s'],1))
PY
PROFILE_MODEL=CNPShapeNet1D python tools/profile_step.py 2>/dev/null | grep -v "Warn\|_warn_once" > gpurun_out/profile_step_cnp1d_r2_final.txt
