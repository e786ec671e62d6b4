#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sim_topk_gpu.py tests/test_fid_gpu.py -q -x > gpurun_out/pytest_r12.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_r12.log
tail -6 gpurun_out/pytest_r12.log
for MB in 1000 48 32 24 16; do
  echo "== chunk MB $MB"
  for C in "10000 100000 512 1" "10000 100000 512 10"; do
    DCR_SIM_CHUNK_MB=$MB timeout 300 python tools/gpu_case.py $C | sed -e 's/stats=.*kernel_ms/kernel_ms/'
  done
done
DCR_SIM_CHUNK_MB=32 timeout 300 python tools/gpu_case.py 50000 125000 512 1 | sed -e 's/stats=.*kernel_ms/kernel_ms/'
timeout 600 python -m pytest tests/test_nets_gpu.py -q -x -s 2>&1 | grep -E "max_abs_err|passed|failed|Error|assert" | head
timeout 300 python tools/net_bench.py vit 256 fast 2>&1 | tail -1
DCR_ATTN_FP32=1 timeout 300 python tools/net_bench.py vit 256 fast 2>&1 | tail -1
