#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section ComputeWorkloadAnalysis --section LaunchStats --clock-control none -k regex:gemm_bf16_kernel -s 108 -c 54 -f -o /tmp/r01_sscd_convs python tools/layer_profile.py run sscd 256 > gpurun_out/ncu_sscd.log 2>&1
tail -2 gpurun_out/ncu_sscd.log
ncu -i /tmp/r01_sscd_convs.ncu-rep --page raw --csv > gpurun_out/r01_sscd_convs_raw.csv 2> gpurun_out/ncu_export.log
ls -la gpurun_out/r01_sscd_convs_raw.csv
