#!/bin/bash
S='s/CASE.*ms_per_call/ms_per_call/; s/stats=.*kernel_ms/kernel_ms/'
run() { echo "== k=$*"; K=$1; shift; env "$@" timeout 300 python tools/gpu_case.py 10000 100000 512 $K | sed -e "$S"; }
run 1 A=0
run 1 DCR_SIM_SETS=2
run 1 DCR_SIM_STAGES=4
run 1 DCR_SIM_KP0=8
run 10 A=0
run 10 DCR_SIM_STAGES=4
run 10 DCR_SIM_STAGES=5 DCR_SIM_CAP=32
run 10 DCR_SIM_KP0=10
run 10 DCR_SIM_DEBUG_EPILOGUE=1
run 10 DCR_SIM_DEBUG_EPILOGUE=2
