#!/bin/bash
S='s/first_call.*ms_per_call/ms_per_call/; s/stats=.*kernel_ms/kernel_ms/'
for C in "300 1000 64 1" "256 4096 512 10" "1000 20000 128 10" "777 33333 100 2"; do
  timeout 120 python tools/gpu_case.py $C | sed -e "$S" || echo "CASE $C FAILED rc=$?"
done
run() { echo "== $*"; K=$1; shift; env "$@" timeout 300 python tools/gpu_case.py 10000 100000 512 $K | sed -e "$S"; }
run 1 A=0
run 1 DCR_SIM_SETS=1
run 10 A=0
run 10 DCR_SIM_KP0=10
run 10 DCR_SIM_KP0=14
run 10 DCR_SIM_SETS=1
run 10 DCR_SIM_DEBUG_EPILOGUE=1
timeout 300 python tools/gpu_case.py 50000 125000 512 10 | sed -e "$S"
timeout 900 python -m pytest tests/test_sim_topk_gpu.py -x -q -m gpu 2>&1 | tail -5
