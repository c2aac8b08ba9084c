#!/bin/bash
# ncu --set full of the implicit-convolution GEMMs (encoder_w0, CNPShapeNet1D): weight-gradient form (B virtual) and forward form (A virtual)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_umma_kernel.*int.2" -s 2 -c 1 -f -o gpurun_out/ncu_convgemm_wgrad_r2 python bench.py --model CNPShapeNet1D --profile --steps 2 --warmup 1 > gpurun_out/ncu_convgemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_umma_kernel.*int.1>" -s 2 -c 1 -f -o gpurun_out/ncu_convgemm_fwd_r2 python bench.py --model CNPShapeNet1D --profile --steps 2 --warmup 1 >> gpurun_out/ncu_convgemm.log 2>&1
ls -la gpurun_out/ncu_convgemm*; tail -3 gpurun_out/ncu_convgemm.log
