#!/bin/bash
# ncu evidence for profiles/ (1 GPU).  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sim_topk_kernel -s 2 -c 1 -f -o gpurun_out/r01_sim_topk_k10 python tools/gpu_case.py 10000 100000 512 10 > gpurun_out/ncu_sim_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_sim_launches_k10.csv python tools/gpu_case.py 10000 100000 512 10 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_sim_launches_k1.csv python tools/gpu_case.py 10000 100000 512 1 > /dev/null 2>&1
# only this library's kernels (ncu matches the bare function name); the weight calibration of bench.py runs torch kernels first
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:gemm_bf16_kernel|sim_topk_kernel|maxpool3_bf16_kernel|pool_kernel|stem_s2d_u8_kernel|im2col_u8_kernel|reduce_hw_kernel|l2_normalize_kernel|to_bf16_rows_kernel|rescore_select_kernel|col_sum_kernel|col_mean_finish_kernel|centre_decision_kernel|gather_rows_kernel|exact_' -c 4000 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 1 --warmup 3 --queries 512 --gallery 4096 --batch 256 --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_sscd.csv python tools/layer_profile.py run sscd 256 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_vit.csv python tools/layer_profile.py run vit 256 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_inception.csv python tools/layer_profile.py run inception 128 > /dev/null 2>&1
timeout 300 python tools/net_bench.py sscd 256 fast | tail -1
timeout 300 python tools/net_bench.py vit 256 fast | tail -1
timeout 300 python tools/net_bench.py inception 128 fast | tail -1
