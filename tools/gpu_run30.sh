#!/bin/bash
# compute-sanitizer on the kernels added at the end of round 2: FAVOR+ cluster / DSMEM forward (smoke), vectorised reducers
# (smoke), second-order batch-norm kernels and fp64 statistics (tests/test_second_order.py subset)
mkdir -p gpurun_out
( time timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python __graft_entry__.py smoke ) > gpurun_out/sanitizer_memcheck.log 2>&1
( time timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python __graft_entry__.py smoke ) > gpurun_out/sanitizer_racecheck.log 2>&1
( time timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python __graft_entry__.py smoke ) > gpurun_out/sanitizer_synccheck.log 2>&1
( time timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_second_order.py -m gpu -q -x -k "maml_step or bn_act or conv_function" ) > gpurun_out/sanitizer_memcheck_second_order.log 2>&1
( time timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_second_order.py -m gpu -q -x -k "maml_step or bn_act" ) > gpurun_out/sanitizer_racecheck_second_order.log 2>&1
for f in memcheck racecheck synccheck memcheck_second_order racecheck_second_order; do echo == $f; grep -E "SUMMARY|passed|failed|smoke:" gpurun_out/sanitizer_$f.log | tail -3; done
python -m pytest tests/test_gpu_ops.py -m gpu -q -k "favor_cluster_split" 2>&1 | tail -2
