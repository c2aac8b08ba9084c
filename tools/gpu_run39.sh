#!/bin/bash
python -m pytest tests/test_gpu_models.py tests/test_parity_at_size.py tests/test_dropin.py -m gpu -q -k "1d or 1D" 2>&1 | tail -3
for i in 1 2; do python bench.py --model CNPShapeNet1D --no-cpu-baseline --no-dropin --steps 40 --warmup 5 2>/dev/null | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('CNPShapeNet1D ms', round(d['ms_per_step'],3), round(d['value'],1))"; done
PROFILE_MODEL=CNPShapeNet1D python tools/profile_step.py 2>/dev/null | grep -v "Warn\|_warn_once" | head -14
