#!/bin/bash
timeout 900 python -m pytest tests/test_sim_topk_gpu.py tests/test_embedding_search.py -x -q -m gpu 2>&1 | tail -4
S='s/CASE.*bad_rows/bad_rows/; s/max_score_err.*ms_per_call/ms_per_call/; s/stats=.*kernel_ms/kernel_ms/'
run() { echo "== k=$*"; K=$1; shift; env "$@" timeout 300 python tools/gpu_case.py 10000 100000 512 $K | sed -e "$S"; }
run 1 A=0
run 1 DCR_SIM_SHARE_THR=0
run 10 A=0
run 10 DCR_SIM_SHARE_THR=0
run 10 DCR_SIM_KP0=10
timeout 300 python tools/gpu_case.py 50000 125000 512 10 | sed -e "$S"
timeout 300 python tools/gpu_case.py 100000 100000 512 2 | sed -e "$S"
