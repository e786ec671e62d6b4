#!/bin/bash
S='s/CASE.*ms_per_call/ms_per_call/; s/stats=.*kernel_ms/kernel_ms/'
run() { echo "== k=$*"; K=$1; shift; env "$@" timeout 300 python tools/gpu_case.py 10000 100000 512 $K | sed -e "$S"; }
for K in 1 10; do
for SETS in 1 2; do
for M in 0 1 2 3; do
  run $K DCR_SIM_SETS=$SETS DCR_SIM_DEBUG_EPILOGUE=$M
done; done; done
