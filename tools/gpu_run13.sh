#!/bin/bash
mkdir -p gpurun_out
python tools/profile_step.py > gpurun_out/profile_step_r2.txt 2>&1
PROFILE_MODEL=CNPShapeNet1D python tools/profile_step.py > gpurun_out/profile_step_cnp1d_r2.txt 2>&1
python tools/timeline_gaps.py > gpurun_out/timeline_r2.txt 2>&1
head -14 gpurun_out/profile_step_r2.txt; head -14 gpurun_out/profile_step_cnp1d_r2.txt; tail -12 gpurun_out/timeline_r2.txt
ncu --set full --clock-control none --import-source on -k regex:tapwgrad_halo_tma_kernel -s 3 -c 1 -f -o gpurun_out/ncu_wgrad_r2 python bench.py --roofline-only > gpurun_out/ncu_wgrad_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tapconv_halo_tma_kernel -s 3 -c 1 -f -o gpurun_out/ncu_halo_r2 python bench.py --roofline-only > gpurun_out/ncu_halo_r2.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"adam_kernel|ctx_agg_fwd_vec4|loss_multi|amp2_flat_fwd" --csv --log-file gpurun_out/ncu_hbm_kernels_r2.csv python tools/hbm_kernels.py > gpurun_out/hbm_kernels_r2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 700 --csv --log-file gpurun_out/launches_r2.csv python bench.py --profile --steps 2 --warmup 3 > gpurun_out/launches_r2.log 2>&1
ls -la gpurun_out | tail -12
