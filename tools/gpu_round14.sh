#!/bin/bash
for M in 0 1 2; do
  echo "== debug epilogue mode $M"
  DCR_SIM_DEBUG_EPILOGUE=$M timeout 300 python tools/gpu_case.py 10000 100000 512 1 2>&1 | sed -e 's/first_call.*ms_per_call/ms_per_call/' -e 's/stats=.*kernel_ms/kernel_ms/'
done
