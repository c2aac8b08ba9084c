#!/bin/bash
python -m pytest tests/test_gpu_ops.py -m gpu -q -k "implicit_conv" 2>&1 | tail -6
python -m pytest tests/test_gpu_models.py tests/test_parity_at_size.py -m gpu -q -k "1d or 1D" 2>&1 | tail -4
for v in 1 0; do B200NP_W0_IMPLICIT=$v python bench.py --model CNPShapeNet1D --no-cpu-baseline --no-dropin --steps 40 --warmup 5 2>/dev/null | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('W0_IMPLICIT=$v', d['config']['workload'][:60], 'ms', round(d['ms_per_step'],3), 'tasks/s', round(d['value'],1))"; done
PROFILE_MODEL=CNPShapeNet1D python tools/profile_step.py 2>/dev/null | grep -v Warn | head -12
