#!/bin/bash
timeout 900 python -m pytest tests/test_nets_gpu.py tests/test_fid_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 300 python tools/net_bench.py sscd 256 fast 10
DCR_POOL_GENERIC=1 timeout 300 python tools/net_bench.py sscd 256 fast 10
timeout 300 python tools/net_bench.py inception 128 fast 10
timeout 300 python tools/net_bench.py inception 50 exact 2
timeout 300 python tools/net_bench.py sscd 64 exact 2
