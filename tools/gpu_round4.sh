#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -q -x --timeout 120 > gpurun_out/pytest_conv.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_conv.log
tail -40 gpurun_out/pytest_conv.log
