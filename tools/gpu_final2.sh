#!/bin/bash
bash tools/gpu_final.sh
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sim_topk_kernel -s 2 -c 1 -f -o gpurun_out/r01_sim_topk_k10 python tools/gpu_case.py 10000 100000 512 10 > gpurun_out/ncu_sim_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_sim_launches_k10.csv python tools/gpu_case.py 10000 100000 512 10 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_sim_launches_k1.csv python tools/gpu_case.py 10000 100000 512 1 > /dev/null 2>&1
