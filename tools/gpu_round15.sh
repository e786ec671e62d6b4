#!/bin/bash
timeout 600 python -m pytest tests/test_sim_topk_gpu.py -q -x 2>&1 | tail -3
for C in "10000 100000 512 1" "10000 100000 512 10" "50000 125000 512 1"; do
  timeout 300 python tools/gpu_case.py $C 2>&1 | sed -e 's/first_call.*ms_per_call/ms_per_call/'
done
