#!/bin/bash
timeout 900 python -m pytest tests/test_conv_gemm_gpu.py tests/test_sim_topk_gpu.py tests/test_nets_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/net_bench.py sscd 256 fast 10
DCR_CONV_NO_HALO=1 timeout 300 python tools/net_bench.py sscd 256 fast 10
timeout 300 python tools/net_bench.py vit 256 fast 10
timeout 300 python tools/net_bench.py inception 128 fast 10
S='s/CASE.*ms_per_call/ms_per_call/; s/stats=.*kernel_ms/kernel_ms/'
timeout 300 python tools/gpu_case.py 10000 100000 512 10 | sed -e "$S"
timeout 300 python tools/gpu_case.py 10000 100000 512 1 | sed -e "$S"
DCR_SIM_DEBUG_EPILOGUE=3 timeout 300 python tools/gpu_case.py 10000 100000 512 1 | sed -e "$S"
