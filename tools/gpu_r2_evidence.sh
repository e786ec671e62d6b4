mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:attention_|centre_decision_kernel|col_mean_finish_kernel|col_sum_kernel|conv3x3_halo_kernel|conv_exact_kernel|exact_scan_kernel|exact_select_kernel|expand_reduce_kernel|gather_rows_kernel|gemm_bf16_kernel|im2col_u8_kernel|l2_normalize_kernel|layernorm|maxpool3_bf16_kernel|pool_kernel|reduce_hw_kernel|rescore_select|sim_topk_kernel|stem_conv_kernel|stem_rows_kernel|stem_s2d_u8_kernel|to_bf16_rows_kernel|topk_merge_kernel' -c 6000 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 1 --warmup 3 --queries 768 --gallery 6144 --no-e2e --parity-steps 0 > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/r02_bench_launches.csv
python tools/launch_summary.py gpurun_out/r02_bench_launches.csv > gpurun_out/r02_bench_launch_summary.txt; cat gpurun_out/r02_bench_launch_summary.txt
DCR_B200_TUNING=1 timeout 300 ncu --set full --clock-control none --import-source on -f -k regex:gemm_bf16_kernel -s 8 -c 1 -o gpurun_out/r02_gemm_pair python tools/conv_ab.py DCR_GEMM_CG2 1 1 256 14 14 1024 256 1 1 0 > /dev/null 2>&1
ls -la gpurun_out/r02_gemm_pair*
