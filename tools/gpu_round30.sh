#!/bin/bash
timeout 600 python -m pytest tests/test_conv_gemm_gpu.py -x -q -m gpu -k halo 2>&1 | tail -15
