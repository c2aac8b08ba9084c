#!/bin/bash
# round-2 GPU validation run #1: tests, bench line, sanitizer
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
( time python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
( time python bench.py ) > gpurun_out/bench_n1.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_n1.log
( time timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke ) > gpurun_out/sanitizer_memcheck.log 2>&1
( time timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke ) > gpurun_out/sanitizer_racecheck.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench_n1.log | cut -c1-1500; tail -5 gpurun_out/sanitizer_memcheck.log; tail -5 gpurun_out/sanitizer_racecheck.log
