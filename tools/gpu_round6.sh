#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py --queries 1000 --gallery 10000 --steps 2 --warmup 1 --cpu-embed-sample 32 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "small exit=$?"; tail -c 1500 gpurun_out/bench_small.json; tail -5 gpurun_out/bench_small.err
timeout 1200 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full exit=$?"; cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json | tail -c 1200
nproc; free -g | head -2
