#!/bin/bash
# ncu evidence for profiles/ (1 GPU).  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sim_topk_kernel -s 2 -c 1 -f -o gpurun_out/r01_sim_topk_k10 python tools/gpu_case.py 10000 100000 512 10 > gpurun_out/ncu_sim_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_sim_launches_k10.csv python tools/gpu_case.py 10000 100000 512 10 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_sim_launches_k1.csv python tools/gpu_case.py 10000 100000 512 1 > /dev/null 2>&1
bash tools/gpu_bench_launches.sh
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_sscd.csv python tools/layer_profile.py run sscd 256 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_vit.csv python tools/layer_profile.py run vit 256 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_inception.csv python tools/layer_profile.py run inception 128 > /dev/null 2>&1
timeout 300 python tools/net_bench.py sscd 256 fast | tail -1
timeout 300 python tools/net_bench.py vit 256 fast | tail -1
timeout 300 python tools/net_bench.py inception 128 fast | tail -1
