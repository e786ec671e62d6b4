#!/bin/bash
S='s/first_call.*ms_per_call/ms_per_call/; s/stats=.*kernel_ms/kernel_ms/'
# background self-similarity at full gallery size (diff_retrieval.py:403,418-419): 100k x 100k, k=2
timeout 300 python tools/gpu_case.py 100000 100000 512 2 | sed -e "$S"
timeout 300 python tools/gpu_case.py 10000 100000 512 10 | sed -e "$S"
timeout 300 python tools/gpu_case.py 10000 100000 384 10 | sed -e "$S"
# bench launch list (this library's kernels only)
bash -c "$(grep -n 'bench.py' tools/gpu_profiles.sh | head -1 | cut -d: -f2-)"
tail -c 300 gpurun_out/bench_under_ncu.log
