#!/bin/bash
S='s/CASE.*bad_rows/bad_rows/; s/max_score_err.*ms_per_call/ms_per_call/; s/stats=.*kernel_ms/kernel_ms/'
run() { echo "== k=$*"; K=$1; shift; env "$@" timeout 300 python tools/gpu_case.py 10000 100000 512 $K | sed -e "$S"; }
run 10 DCR_SIM_CHUNK_MB=25
run 10 DCR_SIM_CHUNK_MB=60
run 10 DCR_SIM_CHUNK_MB=110
run 10 DCR_SIM_STAGES=4
run 1 DCR_SIM_CHUNK_MB=110
