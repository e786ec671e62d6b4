#!/bin/bash
for D in 1 2 3; do
  DCR_HALO_DEBUG=$D timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lp_sscd_h$D.csv python tools/layer_profile.py run sscd 256 > /dev/null 2>&1
done
ls gpurun_out/lp_sscd_h*.csv
