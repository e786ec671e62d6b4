#!/bin/bash
# follow-up captures: the conv3+conv1 pair kernel, bench launch list, sim launch list after the conversion-kernel fix
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 $NCU -k regex:expand_reduce_kernel -s 3 -c 1 -o gpurun_out/r02_expand_reduce_pair python tools/layer_profile.py run sscd 256 > /dev/null 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:attention_|centre_decision_kernel|col_mean_finish_kernel|col_sum_kernel|conv3x3_halo_kernel|conv_exact_kernel|exact_scan_kernel|exact_select_kernel|expand_reduce_kernel|gather_rows_kernel|gemm_bf16_kernel|im2col_u8_kernel|l2_normalize_kernel|layernorm|maxpool3_bf16_kernel|pool_kernel|reduce_hw_kernel|rescore_select_kernel|sim_topk_kernel|stem_conv_kernel|stem_rows_kernel|stem_s2d_u8_kernel|to_bf16_rows_kernel|topk_merge_kernel' -c 4000 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 1 --warmup 3 --queries 512 --gallery 4096 --no-e2e --parity-steps 0 > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/r02_bench_launches.csv
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_sim_launches_k10.csv python tools/gpu_case.py 10000 100000 512 10 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r02_sim_launches_k10.csv
timeout 120 python tools/gpu_case.py 10000 100000 512 10 | cut -c1-200
timeout 120 python tools/gpu_case.py 50000 125000 512 10 | cut -c1-200
timeout 120 python tools/gpu_case.py 100000 100000 512 2 | cut -c1-200
