#!/bin/bash
# ncu launch list of the bench's profile command on the final code (serialised / cold cache: compare shares)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 700 --csv --log-file gpurun_out/launches_r2.csv python bench.py --profile --steps 2 --warmup 3 > gpurun_out/launches_r2.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r2.csv gpurun_out/launches_r2_summary.txt
head -14 gpurun_out/launches_r2_summary.txt
