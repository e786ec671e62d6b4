#!/bin/bash
S='s/CASE.*ms_per_call/ms_per_call/; s/stats=.*kernel_ms/kernel_ms/'
run() { echo "== $*"; env "$@" timeout 300 python tools/gpu_case.py 10000 100000 512 10 | sed -e "$S"; }
run DCR_SIM_DEBUG_EPILOGUE=3
run DCR_SIM_DEBUG_EPILOGUE=3 DCR_SIM_KERNEL=smem
run DCR_SIM_DEBUG_EPILOGUE=2 DCR_SIM_STAGES=4
run DCR_SIM_DEBUG_EPILOGUE=2 DCR_SIM_STAGES=8
run DCR_SIM_DEBUG_EPILOGUE=2 DCR_SIM_STAGES=16 DCR_SIM_CAP=32
run DCR_SIM_DEBUG_EPILOGUE=2 DCR_SIM_KERNEL=smem
