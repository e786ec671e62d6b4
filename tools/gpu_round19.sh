#!/bin/bash
for KP in 16 12 10; do
  echo "== kp0=$KP k=10"
  DCR_SIM_KP0=$KP timeout 300 python tools/gpu_case.py 10000 100000 512 10 | sed -e 's/first_call.*ms_per_call/ms_per_call/' -e 's/stats=.*kernel_ms/kernel_ms/'
done
for KP in 4 2; do
  echo "== kp0=$KP k=1"
  DCR_SIM_KP0=$KP timeout 300 python tools/gpu_case.py 10000 100000 512 1 | sed -e 's/first_call.*ms_per_call/ms_per_call/' -e 's/stats=.*kernel_ms/kernel_ms/'
done
