#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_final.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit=$?"; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2>&1; tail -c 600 gpurun_out/bench_ref.json
