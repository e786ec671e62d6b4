#!/bin/bash
timeout 900 python -m pytest tests/test_sim_topk_gpu.py tests/test_nets_gpu.py -x -q -m gpu 2>&1 | tail -5
S='s/first_call.*ms_per_call/ms_per_call/; s/stats=.*kernel_ms/kernel_ms/'
timeout 300 python tools/gpu_case.py 10000 100000 1024 10 | sed -e "$S"
timeout 300 python tools/gpu_case.py 10000 100000 768 1 | sed -e "$S"
