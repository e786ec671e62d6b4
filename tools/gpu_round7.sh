#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py tests/test_nets_gpu.py -q -x -s > gpurun_out/pytest_r7.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_r7.log
grep -E "max_abs_err|passed|failed|Error|error|exit=|assert" gpurun_out/pytest_r7.log | head -30
for A in "sscd 256 fast" "vit 256 fast"; do
  timeout 300 python tools/net_bench.py $A 2>&1 | tail -2
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_sscd128.csv python tools/net_bench.py sscd 128 fast 1 > /dev/null 2>&1
