#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sim_topk_kernel -c 1 -o gpurun_out/sim_topk_r1b python tools/gpu_case.py 10000 100000 512 1 > gpurun_out/ncu_full_b.log 2>&1
timeout 900 python bench.py --no-e2e > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; cat gpurun_out/bench_q.json | python -c "import json,sys; d=json.load(sys.stdin); print({k:d[k] for k in ['value','ms_per_step','images_embedded_per_s']}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['launch'])"
