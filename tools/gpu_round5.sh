#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nets_gpu.py -q -x -s > gpurun_out/pytest_nets.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_nets.log
grep -E "max_abs_err|passed|failed|Error|error|exit=" gpurun_out/pytest_nets.log | head -40
for A in "sscd 128 fast" "sscd 256 fast" "vit 128 fast" "vit 256 fast" "sscd 32 parity"; do
  timeout 300 python tools/net_bench.py $A 2>&1 | tail -2
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_sscd128.csv python tools/net_bench.py sscd 128 fast 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_vit128.csv python tools/net_bench.py vit 128 fast 1 > /dev/null 2>&1
