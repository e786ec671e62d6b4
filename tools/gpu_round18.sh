#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 4 -c 1 -o gpurun_out/gemm_l1c3 python tools/layer_profile.py run sscd 256 > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/ncu_gemm.log
