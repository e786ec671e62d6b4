#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -s > gpurun_out/pytest_r9.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_r9.log
grep -E "max_abs_err|inception|passed|failed|Error|error|exit=|assert" gpurun_out/pytest_r9.log | head -30
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "full exit=$?"; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
