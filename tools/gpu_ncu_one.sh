#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16_kernel -s ${SKIP:-110} -c 1 -f -o gpurun_out/r01_conv_one python tools/layer_profile.py run sscd 256 > gpurun_out/ncu_one.log 2>&1
tail -2 gpurun_out/ncu_one.log
ls -la gpurun_out/*.ncu-rep
