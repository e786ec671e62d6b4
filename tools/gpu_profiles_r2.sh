#!/bin/bash
# Round-2 ncu evidence for profiles/ (1 GPU).  Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
rm -f gpurun_out/r02_*.ncu-rep
NCU="ncu --set full --clock-control none --import-source on -f"
# the kernels new in round 2, one launch each (batch 256 SSCD forward)
timeout 300 $NCU -k regex:expand_reduce_kernel -s 6 -c 1 -o gpurun_out/r02_expand_reduce python tools/layer_profile.py run sscd 256 > /dev/null 2>&1
timeout 300 $NCU -k regex:stem_conv_kernel -s 2 -c 1 -o gpurun_out/r02_stem_conv python tools/layer_profile.py run sscd 256 > /dev/null 2>&1
timeout 300 $NCU -k regex:conv3x3_halo_kernel -s 6 -c 1 -o gpurun_out/r02_halo_l1 python tools/layer_profile.py run sscd 256 > /dev/null 2>&1
# the graded kernel again (conversion / re-score kernels changed around it)
timeout 600 $NCU -k regex:sim_topk_kernel -s 2 -c 1 -o gpurun_out/r02_sim_topk_k10 python tools/gpu_case.py 10000 100000 512 10 > gpurun_out/ncu_sim_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_sim_launches_k10.csv python tools/gpu_case.py 10000 100000 512 10 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_sim_launches_k1.csv python tools/gpu_case.py 10000 100000 512 1 > /dev/null 2>&1
# launch list of a (small) bench.py step: same code path as the default run
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:attention_|centre_decision_kernel|col_mean_finish_kernel|col_sum_kernel|conv3x3_halo_kernel|conv_exact_kernel|exact_scan_kernel|exact_select_kernel|expand_reduce_kernel|gather_rows_kernel|gemm_bf16_kernel|im2col_u8_kernel|l2_normalize_kernel|layernorm|maxpool3_bf16_kernel|pool_kernel|reduce_hw_kernel|rescore_select_kernel|sim_topk_kernel|stem_conv_kernel|stem_rows_kernel|stem_s2d_u8_kernel|to_bf16_rows_kernel|topk_merge_kernel' -c 4000 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 1 --warmup 3 --queries 512 --gallery 4096 --no-e2e --parity-steps 0 > gpurun_out/bench_under_ncu.log 2>&1
tail -c 200 gpurun_out/bench_under_ncu.log; wc -l gpurun_out/r02_bench_launches.csv
# per-layer tables
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_vit.csv python tools/layer_profile.py run vit 256 > /dev/null 2>&1
python tools/layer_profile.py report gpurun_out/lp_vit.csv vit 256 > gpurun_out/r02_layers_vit.txt 2>&1; tail -1 gpurun_out/r02_layers_vit.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_inception.csv python tools/layer_profile.py run inception 128 > /dev/null 2>&1
python tools/layer_profile.py report gpurun_out/lp_inception.csv inception 128 > gpurun_out/r02_layers_inception.txt 2>&1; tail -1 gpurun_out/r02_layers_inception.txt
ls -la gpurun_out/*.ncu-rep
