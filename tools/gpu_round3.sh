#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for C in "10000 100000 512 1" "10000 100000 512 10" "10000 100000 384 1" "50000 125000 512 1"; do
  set -- $C
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_$1_$2_$3_$4.csv python tools/gpu_case.py $C > /dev/null 2>&1
  python tools/gpu_case.py $C
done
