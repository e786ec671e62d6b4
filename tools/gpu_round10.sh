#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r10.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_r10.log
tail -8 gpurun_out/pytest_r10.log
for C in "10000 100000 512 1" "10000 100000 512 10" "50000 125000 512 1"; do
  timeout 300 python tools/gpu_case.py $C
done
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full exit=$?"; cat gpurun_out/bench_full.json | python -c "import json,sys; d=json.load(sys.stdin); print({k:d[k] for k in ['value','ms_per_step','images_embedded_per_s']}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['launch'], d['e2e'])"
