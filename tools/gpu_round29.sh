#!/bin/bash
# load-only timing modes of the GEMM kernel: per-layer device times of one SSCD forward (results are garbage)
for D in 1 2; do
  DCR_GEMM_DEBUG=$D timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_sscd_dbg$D.csv python tools/layer_profile.py run sscd 256 > /dev/null 2>&1
done
ls -la gpurun_out/lp_sscd_dbg*.csv
