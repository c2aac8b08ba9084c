#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_models.py tests/test_parity_at_size.py tests/test_dropin.py -m gpu -x -q -k "1d or 1D" ) > gpurun_out/pytest_1d.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_1d.log
tail -5 gpurun_out/pytest_1d.log | cut -c1-300
for m in CNPShapeNet1D ANPShapeNet1D ANP; do
python bench.py --model $m --no-cpu-baseline --no-dropin 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$m', 'ms/step', d['ms_per_step'], 'value', d['value'], 'launches/step', d['gpu_launches']/d['steps'])"
done | tee gpurun_out/models_r2.log
