#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for K in 1 10; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_k$K.csv python tools/gpu_case.py 10000 100000 512 $K > gpurun_out/ncu_case_k$K.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sim_topk_kernel -c 1 -o gpurun_out/sim_topk_r1 python tools/gpu_case.py 10000 100000 512 10 > gpurun_out/ncu_full.log 2>&1
python tools/gpu_case.py 10000 100000 512 10
python tools/gpu_case.py 10000 100000 384 10
python tools/gpu_case.py 50000 125000 512 1
