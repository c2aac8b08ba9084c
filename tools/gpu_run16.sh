#!/bin/bash
for r in "" 0 1; do
B200NP_WGRAD_ONLY_ROLE=$r python bench.py --roofline-only 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('only_role=[$r] wgrad ms', d['second_kernel']['launch_ms'], 'fwd ms', d['launch_ms'])"
done
python bench.py --no-cpu-baseline --no-dropin 2>/dev/null | grep '^{' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 ms/step', d['ms_per_step'], 'value', d['value'])"
