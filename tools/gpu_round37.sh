#!/bin/bash
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_nets_gpu.py -x -q -m gpu 2>&1 | tail -6
timeout 300 python tools/net_bench.py sscd 256 fast 10
DCR_GEMM_NO_ARES=1 timeout 300 python tools/net_bench.py sscd 256 fast 10
timeout 300 python tools/net_bench.py vit 256 fast 10
timeout 300 python tools/net_bench.py inception 128 fast 10
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_sscd.csv python tools/layer_profile.py run sscd 256 > /dev/null 2>&1
