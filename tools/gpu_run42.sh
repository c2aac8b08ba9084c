#!/bin/bash
python -m pytest tests/test_gpu_models.py tests/test_parity_at_size.py -m gpu -q -k "3d or 3D or anp_full or cnp_np" 2>&1 | tail -3
for v in 1 0; do B200NP_STEM_GEMM=$v python bench.py --model ANP --no-cpu-baseline --no-dropin --steps 40 --warmup 5 2>/dev/null | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('STEM_GEMM=$v ANP ms', round(d['ms_per_step'],3), round(d['value'],1))"; done
PROFILE_MODEL=ANP python tools/profile_step.py 2>/dev/null | grep -v "Warn\|_warn_once" | head -9
