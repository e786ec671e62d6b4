#!/bin/bash
for ST in 3 4 5; do
  echo "== stages $ST"
  DCR_SIM_STAGES=$ST timeout 300 python tools/gpu_case.py 10000 100000 512 1 | sed -e 's/first_call.*ms_per_call/ms_per_call/'
  DCR_SIM_STAGES=$ST timeout 300 python tools/gpu_case.py 10000 100000 512 10 | sed -e 's/first_call.*ms_per_call/ms_per_call/'
done
DCR_SIM_CHUNK_MB=24 timeout 300 python tools/gpu_case.py 10000 100000 512 1 | sed -e 's/first_call.*ms_per_call/ms_per_call/'
timeout 300 python tools/gpu_case.py 10000 100000 384 1 | sed -e 's/first_call.*ms_per_call/ms_per_call/'
