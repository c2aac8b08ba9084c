#!/bin/bash
# round-2 GPU validation run #2: full GPU test suite (no -x: see every failure), then the bench line
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --durations=25 -rs ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( time python bench.py --no-cpu-baseline ) > gpurun_out/bench_n1.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_n1.log
tail -40 gpurun_out/pytest_gpu.log
