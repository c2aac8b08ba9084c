#!/bin/bash
mkdir -p gpurun_out
TIMELINE_TRACE=gpurun_out/trace_early.txt python tools/timeline_gaps.py > gpurun_out/timeline_early.txt 2>&1
B200NP_FORK=late TIMELINE_TRACE=gpurun_out/trace_late.txt python tools/timeline_gaps.py > gpurun_out/timeline_late.txt 2>&1
head -5 gpurun_out/timeline_early.txt gpurun_out/timeline_late.txt
