#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -m gpu -q -x -k "favor or anp or ANP" 2>&1 | tail -4
run() { # label, env...
  label=$1; shift
  r=$(env "$@" python bench.py --no-cpu-baseline --no-dropin --steps 40 --warmup 5 2>/dev/null | grep '^{' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])")
  echo "$label: $r" | tee -a gpurun_out/ab_run24.txt
}
: > gpurun_out/ab_run24.txt
run "early split4 (default)" X=1
run "early split1" B200NP_FAVOR_SPLIT=1
run "early split2" B200NP_FAVOR_SPLIT=2
run "late reserve0" B200NP_FORK=late
run "late reserve8" B200NP_FORK=late B200NP_RESERVE_SMS=8
run "late reserve16" B200NP_FORK=late B200NP_RESERVE_SMS=16
run "late reserve24" B200NP_FORK=late B200NP_RESERVE_SMS=24
run "late reserve32" B200NP_FORK=late B200NP_RESERVE_SMS=32
run "late reserve48" B200NP_FORK=late B200NP_RESERVE_SMS=48
run "early stemwg4" B200NP_STEM_WGRAD_CTAS=4
run "early stemwg2" B200NP_STEM_WGRAD_CTAS=2
run "early stemwg16" B200NP_STEM_WGRAD_CTAS=16
B200NP_FORK=late B200NP_RESERVE_SMS=16 TIMELINE_TRACE=gpurun_out/trace_late16.txt python tools/timeline_gaps.py > gpurun_out/timeline_late16.txt 2>&1
head -5 gpurun_out/timeline_late16.txt
