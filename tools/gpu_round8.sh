#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r8.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_r8.log
tail -12 gpurun_out/pytest_r8.log
for C in "10000 100000 512 1" "10000 100000 512 10" "10000 100000 384 10" "50000 125000 512 10"; do
  timeout 300 python tools/gpu_case.py $C
done
for C in "10000 100000 512 1" "10000 100000 512 10"; do
  set -- $C
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_$1_$2_$3_$4.csv python tools/gpu_case.py $C > /dev/null 2>&1
done
timeout 300 python tools/net_bench.py sscd 256 fast 2>&1 | tail -1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_sscd128.csv python tools/net_bench.py sscd 128 fast 1 > /dev/null 2>&1
