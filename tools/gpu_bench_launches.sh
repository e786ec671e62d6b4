#!/bin/bash
# launch list of this library's kernels in a (small) bench.py step; ncu matches the bare function name
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:gemm_bf16_kernel|conv3x3_halo_kernel|sim_topk_kernel|maxpool3_bf16_kernel|pool_kernel|stem_s2d_u8_kernel|im2col_u8_kernel|reduce_hw_kernel|l2_normalize_kernel|to_bf16_rows_kernel|rescore_select_kernel|col_sum_kernel|col_mean_finish_kernel|centre_decision_kernel|gather_rows_kernel|exact_' -c 4000 --csv --log-file gpurun_out/r01_bench_launches.csv python bench.py --steps 1 --warmup 3 --queries 512 --gallery 4096 --batch 256 --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
tail -c 200 gpurun_out/bench_under_ncu.log; wc -l gpurun_out/r01_bench_launches.csv
