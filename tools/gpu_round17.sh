#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_nets_gpu.py tests/test_fid_gpu.py -q -x -s 2>&1 | grep -E "max_abs_err|inception|passed|failed|Error|assert" | head -20
for A in "sscd 256 fast" "vit 256 fast"; do timeout 300 python tools/net_bench.py $A 2>&1 | tail -1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_sscd.csv python tools/layer_profile.py run sscd 256 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_vit.csv python tools/layer_profile.py run vit 256 > /dev/null 2>&1
