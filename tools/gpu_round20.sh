#!/bin/bash
# TMEM-resident-query kernel: correctness on small cases first (under timeout), then timing vs the smem kernel
S='s/first_call.*ms_per_call/ms_per_call/; s/stats=.*kernel_ms/kernel_ms/'
for C in "300 1000 64 1" "256 4096 512 10" "1000 20000 128 10" "777 33333 100 2"; do
  timeout 120 python tools/gpu_case.py $C | sed -e "$S" || echo "CASE $C FAILED rc=$?"
done
for K in 1 10; do
  echo "== ta k=$K"
  timeout 300 python tools/gpu_case.py 10000 100000 512 $K | sed -e "$S" || echo FAILED
  echo "== smem k=$K"
  DCR_SIM_KERNEL=smem timeout 300 python tools/gpu_case.py 10000 100000 512 $K | sed -e "$S" || echo FAILED
done
echo "== ta k=10 floor (no epilogue loads)"
DCR_SIM_DEBUG_EPILOGUE=2 timeout 300 python tools/gpu_case.py 10000 100000 512 10 | sed -e "$S"
echo "== ta k=10 loads only"
DCR_SIM_DEBUG_EPILOGUE=1 timeout 300 python tools/gpu_case.py 10000 100000 512 10 | sed -e "$S"
timeout 900 python -m pytest tests/test_sim_topk_gpu.py -x -q -m gpu 2>&1 | tail -5
